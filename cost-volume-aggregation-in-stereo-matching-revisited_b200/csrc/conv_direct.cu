// Generic direct convolution on CUDA cores, fp32 accumulate, fused BN scale/shift + activation +
// residual epilogue.
//
// This is the "every shape" member of the conv family: 3x3x3 stride 1 / stride 2, transposed
// 3x3x3 stride 2 (as 8 output-parity classes), 1x1x1, and 3x3 2-D (depth 1).  The tcgen05
// implicit-GEMM kernels (conv_tc.cu) take over the heavy layers; this kernel is also the
// on-device cross-check for them.  Replaces nn.Conv3d / nn.ConvTranspose3d + nn.BatchNorm3d +
// nn.ReLU + residual adds of /root/reference/models/submodule.py:121-124,
// models/gwcnet_dca_g.py:141-148,224-225, models/augment/cva.py:16-29,39-56.
//
// Formulation: a "tile space" of Dt x Ht x Wt points; point t reads input voxels
// t*in_stride + tap_off[k] and writes output voxel t*out_stride + out_off.
//   conv s1   : in_stride 1, out_stride 1, 27 taps with offsets k-1
//   conv s2   : in_stride 2, out_stride 1, 27 taps with offsets k-1
//   deconv    : per output parity class p: in_stride 1, out_stride 2, out_off p, taps {k: (p+1-k) even},
//               offset (p+1-k)/2  in {0,1}          (o = 2i - 1 + k)
// CTA = 256 threads, tile 1 x 4 x 32 points x NT output channels; Cin is consumed in chunks of CK
// through shared memory (input halo tile + the chunk's weights for all taps).
#include <cstring>

#include "dca_common.cuh"

namespace dca {

constexpr int CD_TH = 4, CD_TW = 32, CD_THREADS = 256;
constexpr int CD_MAX_TAPS = 27;

struct ConvDirectParams {
  const __nv_bfloat16* x; size_t x_plane; int planes_in;
  const float* w;                 // [taps][Cin][CoutPad] fp32
  const float* scale; const float* shift;   // [CoutPad] or null
  const __nv_bfloat16* res_pre; const __nv_bfloat16* res_post; size_t res_plane; int planes_res;
  void* y; size_t y_plane; int planes_out; int out_kind;   // 0 bf16 planes, 1 fp32 channels-last
  int B, Cin, Cout, CoutPad;
  int Di, Hi, Wi, Do, Ho, Wo;
  int Dt, Ht, Wt;
  int in_stride, out_stride;
  int out_off[3];
  int ntaps;
  signed char tap_off[CD_MAX_TAPS][3];
  signed char tap_w[CD_MAX_TAPS];
  int off_min[3];   // min tap offset per axis
  int ext[3];       // input tile extent per axis (d,h,w)
  int act;
};

template <int NT, int CK>
__global__ void __launch_bounds__(CD_THREADS)
conv_direct_kernel(const ConvDirectParams p) {
  constexpr int PITCH = CK + 4;            // floats per staged voxel (pad kills LDS.128 bank conflicts)
  constexpr int NCG = NT / 4;              // cout groups of 4
  constexpr int NVG = CD_THREADS / NCG;    // voxel groups
  constexpr int VPT = (CD_TH * CD_TW) / NVG;
  extern __shared__ __align__(16) float smem[];
  const int tile_vox = p.ext[0] * p.ext[1] * p.ext[2];
  float* xs = smem;                                  // [tile_vox][PITCH]
  float* ws = smem + (size_t)tile_vox * PITCH;       // [ntaps][CK][NT]

  const int wtiles = (p.Wt + CD_TW - 1) / CD_TW;
  const int tw0 = (blockIdx.x % wtiles) * CD_TW;
  const int th0 = (blockIdx.x / wtiles) * CD_TH;
  const int td = blockIdx.y % p.Dt;
  const int cot = blockIdx.y / p.Dt;                 // cout tile
  const int b = blockIdx.z;
  const int co0 = cot * NT;
  const int tid = threadIdx.x;
  const int cg = tid % NCG, vg = tid / NCG;

  // input-tile origin in input coordinates
  const int id0 = td * p.in_stride + p.off_min[0];
  const int ih0 = th0 * p.in_stride + p.off_min[1];
  const int iw0 = tw0 * p.in_stride + p.off_min[2];

  float acc[VPT][4];
#pragma unroll
  for (int k = 0; k < VPT; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;

  // per-thread base smem voxel index (without tap offset) for each of its VPT points
  int vbase[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    int v = vg + NVG * k;
    int hh = v / CD_TW, ww = v % CD_TW;
    vbase[k] = (hh * p.in_stride) * p.ext[2] + ww * p.in_stride;
  }

  for (int c0 = 0; c0 < p.Cin; c0 += CK) {
    __syncthreads();
    // ---- stage the input halo tile: CK channels of every voxel, zero outside the tensor ----
    for (int it = tid; it < tile_vox * (CK / 8); it += CD_THREADS) {
      int vox = it / (CK / 8), ch8 = it % (CK / 8);
      int dx = vox % p.ext[2];
      int r = vox / p.ext[2];
      int dy = r % p.ext[1], dz = r / p.ext[1];
      int iz = id0 + dz, iy = ih0 + dy, ix = iw0 + dx;
      float f[8];
      if (iz >= 0 && iz < p.Di && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) {
        size_t off = ((((size_t)b * p.Di + iz) * p.Hi + iy) * p.Wi + ix) * p.Cin + c0 + ch8 * 8;
        load8_rt(p.x, p.x_plane, p.planes_in, off, f);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(xs + (size_t)vox * PITCH + ch8 * 8);
      dst[0] = make_float4(f[0], f[1], f[2], f[3]);
      dst[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    // ---- stage the weights of this Cin chunk: [tap][CK][NT] ----
    for (int it = tid; it < p.ntaps * CK * (NT / 4); it += CD_THREADS) {
      int q = it % (NT / 4);
      int r = it / (NT / 4);
      int ci = r % CK, tp = r / CK;
      const float* src = p.w + ((size_t)p.tap_w[tp] * p.Cin + c0 + ci) * p.CoutPad + co0 + q * 4;
      reinterpret_cast<float4*>(ws)[it] = *reinterpret_cast<const float4*>(src);
    }
    __syncthreads();

    for (int tp = 0; tp < p.ntaps; ++tp) {
      const int toff = ((p.tap_off[tp][0] - p.off_min[0]) * p.ext[1] + (p.tap_off[tp][1] - p.off_min[1])) * p.ext[2] +
                       (p.tap_off[tp][2] - p.off_min[2]);
      const float* wt = ws + (size_t)tp * CK * NT + cg * 4;
#pragma unroll
      for (int c4 = 0; c4 < CK; c4 += 4) {
        float4 w0 = *reinterpret_cast<const float4*>(wt + (c4 + 0) * NT);
        float4 w1 = *reinterpret_cast<const float4*>(wt + (c4 + 1) * NT);
        float4 w2 = *reinterpret_cast<const float4*>(wt + (c4 + 2) * NT);
        float4 w3 = *reinterpret_cast<const float4*>(wt + (c4 + 3) * NT);
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
          float4 xv = *reinterpret_cast<const float4*>(xs + (size_t)(vbase[k] + toff) * PITCH + c4);
          acc[k][0] = fmaf(xv.x, w0.x, acc[k][0]); acc[k][1] = fmaf(xv.x, w0.y, acc[k][1]);
          acc[k][2] = fmaf(xv.x, w0.z, acc[k][2]); acc[k][3] = fmaf(xv.x, w0.w, acc[k][3]);
          acc[k][0] = fmaf(xv.y, w1.x, acc[k][0]); acc[k][1] = fmaf(xv.y, w1.y, acc[k][1]);
          acc[k][2] = fmaf(xv.y, w1.z, acc[k][2]); acc[k][3] = fmaf(xv.y, w1.w, acc[k][3]);
          acc[k][0] = fmaf(xv.z, w2.x, acc[k][0]); acc[k][1] = fmaf(xv.z, w2.y, acc[k][1]);
          acc[k][2] = fmaf(xv.z, w2.z, acc[k][2]); acc[k][3] = fmaf(xv.z, w2.w, acc[k][3]);
          acc[k][0] = fmaf(xv.w, w3.x, acc[k][0]); acc[k][1] = fmaf(xv.w, w3.y, acc[k][1]);
          acc[k][2] = fmaf(xv.w, w3.z, acc[k][2]); acc[k][3] = fmaf(xv.w, w3.w, acc[k][3]);
        }
      }
    }
  }

  // ---- epilogue: BN scale/shift, residual(s), activation, store ----
  const int co = co0 + cg * 4;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = p.scale ? __ldg(p.scale + co + j) : 1.f;
    sh[j] = p.shift ? __ldg(p.shift + co + j) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    int v = vg + NVG * k;
    int th = th0 + v / CD_TW, tw = tw0 + v % CD_TW;
    if (th >= p.Ht || tw >= p.Wt) continue;
    int oz = td * p.out_stride + p.out_off[0], oy = th * p.out_stride + p.out_off[1],
        ox = tw * p.out_stride + p.out_off[2];
    if (oz >= p.Do || oy >= p.Ho || ox >= p.Wo) continue;
    size_t vox = (((size_t)b * p.Do + oz) * p.Ho + oy) * p.Wo + ox;
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = acc[k][j] * sc[j] + sh[j];
    if (p.out_kind == 0) {
      if (co >= p.Cout) continue;           // Cout is a multiple of 8 for plane outputs
      size_t off = vox * p.Cout + co;
      if (p.res_pre) {
        uint2 hv = *reinterpret_cast<const uint2*>(p.res_pre + off);
        { const float2 a_ = h16x2_to_f2(hv.x), b_ = h16x2_to_f2(hv.y); r[0] += a_.x; r[1] += a_.y; r[2] += b_.x; r[3] += b_.y; }
        if (p.planes_res == 2) {
          uint2 lv = *reinterpret_cast<const uint2*>(p.res_pre + p.res_plane + off);
          { const float2 a_ = h16x2_to_f2(lv.x), b_ = h16x2_to_f2(lv.y); r[0] += a_.x; r[1] += a_.y; r[2] += b_.x; r[3] += b_.y; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = apply_act(r[j], p.act);
      if (p.res_post) {
        uint2 hv = *reinterpret_cast<const uint2*>(p.res_post + off);
        { const float2 a_ = h16x2_to_f2(hv.x), b_ = h16x2_to_f2(hv.y); r[0] += a_.x; r[1] += a_.y; r[2] += b_.x; r[3] += b_.y; }
        if (p.planes_res == 2) {
          uint2 lv = *reinterpret_cast<const uint2*>(p.res_post + p.res_plane + off);
          { const float2 a_ = h16x2_to_f2(lv.x), b_ = h16x2_to_f2(lv.y); r[0] += a_.x; r[1] += a_.y; r[2] += b_.x; r[3] += b_.y; }
        }
      }
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = split_bf16(r[j], l[j]);
      __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y);
      *reinterpret_cast<uint2*>(yb + off) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
      if (p.planes_out == 2)
        *reinterpret_cast<uint2*>(yb + p.y_plane + off) = make_uint2(l[0] | (l[1] << 16), l[2] | (l[3] << 16));
    } else {
      float* yf = reinterpret_cast<float*>(p.y);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co + j < p.Cout) yf[vox * p.Cout + co + j] = apply_act(r[j], p.act);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3x3 stride-1 conv to ONE output channel, fp32 logits out [B,D,H,W]
// (cva.classify.2 cva.py:53 and classif3.2 gwcnet_dca_g.py:168).
// Marching formulation: the CTA owns an 8h x 32w tile and walks a chunk of depth planes.  For every
// input plane each thread computes, for ONE (halo) voxel, the 27 dot products <x[v,:], w[tap,:]>
// (weights are kernel parameters, i.e. constant-bank FFMA operands: no load per FMA), parks them in
// shared memory, and the output threads gather the 9 in-plane neighbours per kd into a 3-deep ring of
// running sums.  Every activation is read once (plus the 1-voxel halo).
// ---------------------------------------------------------------------------------------------
constexpr int C1_TH = 8, C1_TW = 32, C1_EH = C1_TH + 2, C1_EW = C1_TW + 2;
constexpr int C1_THREADS = 352;           // >= C1_EH * C1_EW = 340
struct Cout1Weights { float w[27 * 32]; };

__global__ void __launch_bounds__(C1_THREADS)
conv3d_cout1_kernel(const __nv_bfloat16* __restrict__ x, size_t x_plane, int planes, const Cout1Weights wp,
                    float* __restrict__ y, int B, int D, int H, int W, int DC) {
  constexpr int CIN = 32;
  __shared__ float P[C1_EH * C1_EW * 27];
  const int wtiles = (W + C1_TW - 1) / C1_TW;
  const int w0 = (blockIdx.x % wtiles) * C1_TW, h0 = (blockIdx.x / wtiles) * C1_TH;
  const int d0 = blockIdx.y * DC, b = blockIdx.z, tid = threadIdx.x;
  const int d_end = min(d0 + DC, D);
  const bool producer = tid < C1_EH * C1_EW;
  const int ih = h0 - 1 + tid / C1_EW, iw = w0 - 1 + tid % C1_EW;
  const bool in_hw = producer && ih >= 0 && ih < H && iw >= 0 && iw < W;
  const bool consumer = tid < C1_TH * C1_TW;
  const int oh = tid / C1_TW, ow = tid % C1_TW;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;       // running sums of output planes d_in-1, d_in, d_in+1
  for (int d_in = d0 - 1; d_in <= d_end; ++d_in) {
    const bool plane_ok = d_in >= 0 && d_in < D;
    if (producer) {
      float acc[27];
#pragma unroll
      for (int t = 0; t < 27; ++t) acc[t] = 0.f;
      if (plane_ok && in_hw) {
        const size_t off = ((((size_t)b * D + d_in) * H + ih) * W + iw) * CIN;
#pragma unroll
        for (int c8 = 0; c8 < CIN; c8 += 8) {
          float f[8];
          load8_rt(x, x_plane, planes, off + c8, f);
#pragma unroll
          for (int t = 0; t < 27; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t] = fmaf(f[j], wp.w[t * CIN + c8 + j], acc[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 27; ++t) P[tid * 27 + t] = acc[t];
    }
    __syncthreads();
    if (consumer) {
      float s[3] = {0.f, 0.f, 0.f};
      if (plane_ok) {
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              s[kd] += P[((oh + kh) * C1_EW + ow + kw) * 27 + (kd * 3 + kh) * 3 + kw];
      }
      // input plane d_in feeds output d_in+1 through kd=0, d_in through kd=1, d_in-1 through kd=2
      r0 += s[2];
      const int od = d_in - 1;
      if (od >= d0 && od < d_end && h0 + oh < H && w0 + ow < W)
        y[(((size_t)b * D + od) * H + h0 + oh) * W + w0 + ow] = r0;
      r0 = r1 + s[1];
      r1 = s[0];
      (void)r2;
    }
    __syncthreads();
  }
}

static void fill_taps(ConvDirectParams& p, int mode, const int parity[3]) {
  // mode 0: k3 s1, 1: k3 s2, 2: deconv parity class, 3: 1x1x1, 4: 2-D 3x3
  p.ntaps = 0;
  int mn[3] = {127, 127, 127}, mx[3] = {-127, -127, -127};
  auto add = [&](int oz, int oy, int ox, int widx) {
    int t = p.ntaps++;
    p.tap_off[t][0] = (signed char)oz; p.tap_off[t][1] = (signed char)oy; p.tap_off[t][2] = (signed char)ox;
    p.tap_w[t] = (signed char)widx;
    int o[3] = {oz, oy, ox};
    for (int a = 0; a < 3; ++a) { if (o[a] < mn[a]) mn[a] = o[a]; if (o[a] > mx[a]) mx[a] = o[a]; }
  };
  if (mode == 0 || mode == 1) {
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw)
      add(kd - 1, kh - 1, kw - 1, (kd * 3 + kh) * 3 + kw);
  } else if (mode == 2) {
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
      int k[3] = {kd, kh, kw}, off[3]; bool ok = true;
      for (int a = 0; a < 3; ++a) { int n = parity[a] + 1 - k[a]; if (n & 1) ok = false; off[a] = n / 2; }
      if (ok) add(off[0], off[1], off[2], (kd * 3 + kh) * 3 + kw);
    }
  } else if (mode == 3) {
    add(0, 0, 0, 0);
  } else {
    for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) add(0, kh - 1, kw - 1, kh * 3 + kw);
  }
  for (int a = 0; a < 3; ++a) p.off_min[a] = mn[a];
  p.ext[0] = (mx[0] - mn[0]) + 1;
  p.ext[1] = (CD_TH - 1) * p.in_stride + (mx[1] - mn[1]) + 1;
  p.ext[2] = (CD_TW - 1) * p.in_stride + (mx[2] - mn[2]) + 1;
}

template <int NT, int CK>
static int launch_direct(const ConvDirectParams& p, cudaStream_t st) {
  size_t smem = ((size_t)p.ext[0] * p.ext[1] * p.ext[2] * (CK + 4) + (size_t)p.ntaps * CK * NT) * sizeof(float);
  if (smem > 227 * 1024) return DCA_ERR_UNSUPPORTED;
  cudaFuncSetAttribute(conv_direct_kernel<NT, CK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int wt = (p.Wt + CD_TW - 1) / CD_TW, ht = (p.Ht + CD_TH - 1) / CD_TH;
  dim3 grid(wt * ht, p.Dt * (p.CoutPad / NT), p.B);
  conv_direct_kernel<NT, CK><<<grid, CD_THREADS, smem, st>>>(p);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

}  // namespace dca

using namespace dca;

// mode: 0 k3 s1 p1 | 1 k3 s2 p1 | 2 transposed k3 s2 p1 op1 | 3 1x1x1 | 4 2-D 3x3 s1 p1 (Di == 1)
extern "C" int dca_conv3d_direct(int mode, const void* x, int planes_in, const float* w_packed, const float* scale,
                                 const float* shift, const void* res_pre, const void* res_post, int planes_res,
                                 void* y, int planes_out, int out_kind, int act, int B, int Cin, int Cout,
                                 int CoutPad, int Di, int Hi, int Wi, int Do, int Ho, int Wo, void* stream) {
  if (!x || !w_packed || !y || B <= 0 || Cin <= 0 || Cout <= 0) return DCA_ERR_ARG;
  if (Cin % 8 != 0 || CoutPad % 32 != 0 || CoutPad < Cout || mode < 0 || mode > 4) return DCA_ERR_ARG;
  if (out_kind == 0 && Cout % 8 != 0) return DCA_ERR_ARG;
  if (planes_in < 1 || planes_in > 2 || planes_out < 1 || planes_out > 2) return DCA_ERR_ARG;
  if ((mode == 0 || mode == 3 || mode == 4) && (Do != Di || Ho != Hi || Wo != Wi)) return DCA_ERR_ARG;
  if (mode == 1 && (Do != (Di + 1) / 2 || Ho != (Hi + 1) / 2 || Wo != (Wi + 1) / 2)) return DCA_ERR_ARG;
  if (mode == 2 && (Do != 2 * Di || Ho != 2 * Hi || Wo != 2 * Wi)) return DCA_ERR_ARG;
  if (mode == 4 && Di != 1) return DCA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  ConvDirectParams p;
  p.x = (const __nv_bfloat16*)x; p.x_plane = (size_t)B * Di * Hi * Wi * Cin; p.planes_in = planes_in;
  p.w = w_packed; p.scale = scale; p.shift = shift;
  p.res_pre = (const __nv_bfloat16*)res_pre; p.res_post = (const __nv_bfloat16*)res_post;
  p.res_plane = (size_t)B * Do * Ho * Wo * Cout; p.planes_res = planes_res;
  p.y = y; p.y_plane = (size_t)B * Do * Ho * Wo * Cout; p.planes_out = planes_out; p.out_kind = out_kind;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.CoutPad = CoutPad;
  p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  p.act = act;
  const int nt64 = (CoutPad % 64 == 0);
  const int ck16 = (Cin % 16 == 0);
  auto run = [&](const ConvDirectParams& q) -> int {
    if (mode == 1) return nt64 ? launch_direct<64, 8>(q, st) : launch_direct<32, 8>(q, st);
    if (nt64) return launch_direct<64, 8>(q, st);
    return ck16 ? launch_direct<32, 16>(q, st) : launch_direct<32, 8>(q, st);
  };
  if (mode == 2) {
    p.in_stride = 1; p.out_stride = 2;
    p.Dt = Di; p.Ht = Hi; p.Wt = Wi;
    for (int pc = 0; pc < 8; ++pc) {
      int parity[3] = {(pc >> 2) & 1, (pc >> 1) & 1, pc & 1};
      p.out_off[0] = parity[0]; p.out_off[1] = parity[1]; p.out_off[2] = parity[2];
      fill_taps(p, 2, parity);
      int rc = run(p);
      if (rc != DCA_OK) return rc;
    }
    return DCA_OK;
  }
  p.in_stride = (mode == 1) ? 2 : 1;
  p.out_stride = 1;
  p.out_off[0] = p.out_off[1] = p.out_off[2] = 0;
  p.Dt = Do; p.Ht = Ho; p.Wt = Wo;
  int zero[3] = {0, 0, 0};
  fill_taps(p, mode, zero);
  return run(p);
}

extern "C" int dca_conv3d_cout1(const void* x, int planes_in, const float* w_host, float* y, int B, int Cin, int D,
                                int H, int W, void* stream) {
  // w_host: HOST pointer [27][Cin] fp32; copied into the launch parameters (constant bank operands).
  if (!x || !w_host || !y || B <= 0 || D <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  if (Cin != 32) return DCA_ERR_UNSUPPORTED;
  Cout1Weights wp;
  memcpy(wp.w, w_host, sizeof(wp.w));
  const int DC = D <= 16 ? D : 16;
  dim3 grid(((W + C1_TW - 1) / C1_TW) * ((H + C1_TH - 1) / C1_TH), (D + DC - 1) / DC, B);
  conv3d_cout1_kernel<<<grid, C1_THREADS, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (size_t)B * D * H * W * Cin, planes_in, wp, y, B, D, H, W, DC);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}
