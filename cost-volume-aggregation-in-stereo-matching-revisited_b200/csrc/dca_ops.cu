// DCA-module kernels and the bandwidth-bound tail of the hot path (SURVEY 8a rows a4-pool, a6, a7,
// a8, a11, a12-upsample) plus layout / weight-packing utilities.
//
//   avgpool3d            nn.AvgPool3d(3, 2, 1)                       models/augment/cva.py:39
//   class_stats          softmax/argmax/per-class sums               models/augment/semantic_level.py:98-119
//   disp_attention       key scaling + q/k/v projections + 4-head attention over disparity + out
//                        projection (+ the aug half of cva.fuse)     semantic_level.py:126,
//                                                                    SelfAttention_bn.py:62-98, cva.py:55,69
//   upsample_fuse        trilinear x2 + cat + 1x1x1 fuse conv + BN    cva.py:64,69
//   softmax_regress      F.softmax(dim=1) + disparity_regression      gwcnet_dca_g.py:238-239, submodule.py:127-131
//   convex_upsample      PropgationNet_4x.forward tail                gwcnet_dca_g.py:118-124
#include "dca_common.cuh"

namespace dca {

// ------------------------------------------------------------------------------------------------
// AvgPool3d k3 s2 p1, count_include_pad=True (always /27).  thread = (output voxel, 8-channel chunk)
// ------------------------------------------------------------------------------------------------
template <int PLANES>
__global__ void __launch_bounds__(256)
avgpool3d_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int C, int Di, int Hi,
                 int Wi, int Do, int Ho, int Wo) {
  const int c8n = C / 8;
  const size_t total = (size_t)B * Do * Ho * Wo * c8n;
  const size_t xin_plane = (size_t)B * Di * Hi * Wi * C, yout_plane = (size_t)B * Do * Ho * Wo * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % c8n);
    size_t r = i / c8n;
    int ow = (int)(r % Wo); r /= Wo;
    int oh = (int)(r % Ho); r /= Ho;
    int od = (int)(r % Do);
    int b = (int)(r / Do);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int dz = -1; dz <= 1; ++dz) {
      int iz = 2 * od + dz;
      if (iz < 0 || iz >= Di) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        int iy = 2 * oh + dy;
        if (iy < 0 || iy >= Hi) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          int ix = 2 * ow + dx;
          if (ix < 0 || ix >= Wi) continue;
          float f[8];
          load8<PLANES>(x, xin_plane, ((((size_t)b * Di + iz) * Hi + iy) * Wi + ix) * C + c8 * 8, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= (1.0f / 27.0f);
    store8<PLANES>(y, yout_plane, ((((size_t)b * Do + od) * Ho + oh) * Wo + ow) * C + c8 * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// class statistics: P = softmax_d(logits); k_p = first argmax_d P; e_p = exp(P[k_p]); S[b,k] += e_p
// logits fp32 [B,D,H,W].  thread = pixel (coalesced along w for every d).
// ------------------------------------------------------------------------------------------------
constexpr int CS_MAXD = 256;
__global__ void __launch_bounds__(256)
class_stats_kernel(const float* __restrict__ logits, int* __restrict__ cls, float* __restrict__ e_out,
                   float* __restrict__ S, int D, int HW) {
  __shared__ float s_sum[CS_MAXD];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s_sum[i] = 0.f;
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < HW) {
    const float* lp = logits + (size_t)b * D * HW + p;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, lp[(size_t)d * HW]);
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(lp[(size_t)d * HW] - m);
    float best = -1.f; int k = 0;
    for (int d = 0; d < D; ++d) {
      float pd = expf(lp[(size_t)d * HW] - m) / s;   // same formula torch's softmax uses
      if (pd > best) { best = pd; k = d; }             // strict > keeps the FIRST maximum
    }
    float e = expf(best);
    cls[(size_t)b * HW + p] = k;
    e_out[(size_t)b * HW + p] = e;
    atomicAdd(&s_sum[k], e);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    if (s_sum[i] != 0.f) atomicAdd(&S[(size_t)b * D + i], s_sum[i]);
}

// ------------------------------------------------------------------------------------------------
// Fused DCA attention.  One warp per pixel, lane = channel (C == 32 == 4 heads x 8).
//   key[d] = x[d] * (1 + [d == k_p] * e_p / S[b,k_p])
//   q = Pq1(Pq0(x)), k = Pk1(Pk0(key)), v = Pv(key);  P = 1x1x1 conv + BN + LeakyReLU(0.1)
//   ctx[dq, head] = softmax_dk(q[dq,head] . k[dk,head] / sqrt(8)) v[dk, head]
//   out = Po(ctx);  optional  t = Wa . out  (the aug half of cva.fuse, linear, commutes with the
//   trilinear upsampling that follows)
// Weights: 7 matrices stored transposed [ci][co] fp32, then 6 x (scale[32], shift[32]).
// ------------------------------------------------------------------------------------------------
constexpr int AT_C = 32;
__device__ __forceinline__ float fast_exp2(float x) {   // ex2.approx: rel. error 2^-22, -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr int AT_NMAT = 7;
constexpr int AT_WFLOATS = AT_NMAT * AT_C * AT_C + 6 * 2 * AT_C;

__device__ __forceinline__ void warp_project(const float* __restrict__ in, float* __restrict__ out,
                                             const float* __restrict__ Wt, const float* __restrict__ ss, int Dp,
                                             int lane, bool affine_act) {
  const float sc = affine_act ? ss[lane] : 1.f, sh = affine_act ? ss[AT_C + lane] : 0.f;
  for (int d0 = 0; d0 < Dp; d0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < AT_C; c4 += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(Wt + (c4 * AT_C + lane * 4));   // [ci/4][co][4]
      const float w0 = wv.x, w1 = wv.y, w2 = wv.z, w3 = wv.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(in + (d0 + j) * AT_C + c4);
        acc[j] = fmaf(xv.x, w0, acc[j]); acc[j] = fmaf(xv.y, w1, acc[j]);
        acc[j] = fmaf(xv.z, w2, acc[j]); acc[j] = fmaf(xv.w, w3, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[j] * sc + sh;
      if (affine_act) v = v > 0.f ? v : 0.1f * v;
      out[(d0 + j) * AT_C + lane] = v;
    }
  }
  __syncwarp();
}

template <int PLANES>
__global__ void __launch_bounds__(256)
disp_attention_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ cls, const float* __restrict__ e,
                      const float* __restrict__ S, const float* __restrict__ wts, int has_wa,
                      __nv_bfloat16* __restrict__ y, int B, int D, int H, int Wd, int pad) {
  const int HW = H * Wd;
  extern __shared__ __align__(16) float smem[];
  float* W = smem;                                   // AT_WFLOATS
  const int Dp = (D + 7) & ~7;
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = smem + ((AT_WFLOATS + 3) & ~3) + (size_t)warp * 4 * Dp * AT_C;
  float* bA = buf, *bB = buf + Dp * AT_C, *bC = buf + 2 * Dp * AT_C, *bD = buf + 3 * Dp * AT_C;
  // matrices arrive transposed [ci][co]; keep them as [ci/4][co][4] so a lane fetches 4 input channels per LDS.128
  for (int i = threadIdx.x; i < AT_NMAT * AT_C * AT_C; i += blockDim.x) {
    const int m = i >> 10, r = i & 1023, ci = r >> 5, co = r & 31;
    W[(m << 10) + ((ci >> 2) * AT_C + co) * 4 + (ci & 3)] = __ldg(wts + i);
  }
  for (int i = AT_NMAT * AT_C * AT_C + threadIdx.x; i < AT_WFLOATS; i += blockDim.x) W[i] = __ldg(wts + i);
  // rows [D, Dp) are never written by the loads below: clear them once so no NaN garbage circulates
  for (int i = lane; i < 4 * Dp * AT_C; i += 32) buf[i] = 0.f;
  __syncthreads();
  const float* Wq0 = W, *Wq1 = W + 1024, *Wk0 = W + 2048, *Wk1 = W + 3072, *Wv = W + 4096, *Wo = W + 5120,
             *Wa = W + 6144;
  const float* ss = W + AT_NMAT * 1024;              // 6 x (scale, shift): q0 q1 k0 k1 v o
  const size_t plane = (size_t)B * D * HW * AT_C;
  const float qscale = 0.35355339059327373f * 1.4426950408889634f;   // 8^-0.5 * log2(e): scores in the log2 domain

  for (int pix = blockIdx.x * warps + warp; pix < B * HW; pix += gridDim.x * warps) {
    const int b = pix / HW, p = pix % HW;
    const int kp = cls[pix];
    const float wp = e[pix] / S[(size_t)b * D + kp];
    // ---- load x[d][0..31] -> bA, key -> bB.  lane = (row r = lane/4, chunk q = lane%4) ----
    for (int d0 = 0; d0 < D; d0 += 8) {
      const int d = d0 + (lane >> 2), q = lane & 3;
      if (d < D) {
        float f[8];
        load8<PLANES>(x, plane, (((size_t)b * D + d) * HW + p) * AT_C + q * 8, f);
        const float ks = (d == kp) ? 1.f + wp : 1.f;
        float4* a = reinterpret_cast<float4*>(bA + d * AT_C + q * 8);
        float4* k = reinterpret_cast<float4*>(bB + d * AT_C + q * 8);
        a[0] = make_float4(f[0], f[1], f[2], f[3]); a[1] = make_float4(f[4], f[5], f[6], f[7]);
        k[0] = make_float4(f[0] * ks, f[1] * ks, f[2] * ks, f[3] * ks);
        k[1] = make_float4(f[4] * ks, f[5] * ks, f[6] * ks, f[7] * ks);
      }
    }
    __syncwarp();
    warp_project(bA, bC, Wq0, ss + 0 * 64, Dp, lane, true);
    warp_project(bC, bD, Wq1, ss + 1 * 64, Dp, lane, true);    // q  in bD
    warp_project(bB, bA, Wk0, ss + 2 * 64, Dp, lane, true);
    warp_project(bA, bC, Wk1, ss + 3 * 64, Dp, lane, true);    // k  in bC
    warp_project(bB, bA, Wv, ss + 4 * 64, Dp, lane, true);     // v  in bA
    // ---- attention: item = (dq, head), 4 heads x D queries ----
    for (int item = lane; item < 4 * D; item += 32) {
      const int dq = item >> 2, hd = item & 3;
      float4 q0 = *reinterpret_cast<const float4*>(bD + dq * AT_C + hd * 8);
      float4 q1 = *reinterpret_cast<const float4*>(bD + dq * AT_C + hd * 8 + 4);
      q0.x *= qscale; q0.y *= qscale; q0.z *= qscale; q0.w *= qscale;
      q1.x *= qscale; q1.y *= qscale; q1.z *= qscale; q1.w *= qscale;
      float m = -INFINITY, l = 0.f, c[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = 0.f;
      for (int dk = 0; dk < D; ++dk) {
        const float4 k0 = *reinterpret_cast<const float4*>(bC + dk * AT_C + hd * 8);
        const float4 k1 = *reinterpret_cast<const float4*>(bC + dk * AT_C + hd * 8 + 4);
        float s = q0.x * k0.x + q0.y * k0.y + q0.z * k0.z + q0.w * k0.w + q1.x * k1.x + q1.y * k1.y + q1.z * k1.z +
                  q1.w * k1.w;
        const float mn = fmaxf(m, s);
        const float corr = fast_exp2(m - mn), pe = fast_exp2(s - mn);
        const float4 v0 = *reinterpret_cast<const float4*>(bA + dk * AT_C + hd * 8);
        const float4 v1 = *reinterpret_cast<const float4*>(bA + dk * AT_C + hd * 8 + 4);
        l = l * corr + pe;
        c[0] = c[0] * corr + pe * v0.x; c[1] = c[1] * corr + pe * v0.y;
        c[2] = c[2] * corr + pe * v0.z; c[3] = c[3] * corr + pe * v0.w;
        c[4] = c[4] * corr + pe * v1.x; c[5] = c[5] * corr + pe * v1.y;
        c[6] = c[6] * corr + pe * v1.z; c[7] = c[7] * corr + pe * v1.w;
        m = mn;
      }
      const float il = 1.f / l;
      float4* o = reinterpret_cast<float4*>(bB + dq * AT_C + hd * 8);
      o[0] = make_float4(c[0] * il, c[1] * il, c[2] * il, c[3] * il);
      o[1] = make_float4(c[4] * il, c[5] * il, c[6] * il, c[7] * il);
    }
    __syncwarp();
    warp_project(bB, bC, Wo, ss + 5 * 64, Dp, lane, true);     // aug_down in bC
    const float* res = bC;
    if (has_wa) { warp_project(bC, bD, Wa, nullptr, Dp, lane, false); res = bD; }
    // ---- store [d][32] rows: lane = (row, chunk).  pad = 1 writes into a tensor with a replicated 1-voxel border
    //      ([D+2][H+2][W+2], what the trilinear-x2 "up2" GEMM consumes) ----
    const int ph = p / Wd, pw = p - ph * Wd;
    const int Dp2 = D + 2 * pad, Hp2 = H + 2 * pad, Wp2 = Wd + 2 * pad;
    const size_t oplane = (size_t)B * Dp2 * Hp2 * Wp2 * AT_C;
    for (int d0 = 0; d0 < D; d0 += 8) {
      const int d = d0 + (lane >> 2), q = lane & 3;
      if (d < D) {
        float f[8];
        const float4 a = *reinterpret_cast<const float4*>(res + d * AT_C + q * 8);
        const float4 c2 = *reinterpret_cast<const float4*>(res + d * AT_C + q * 8 + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c2.x; f[5] = c2.y; f[6] = c2.z; f[7] = c2.w;
        if (!pad) {
          store8<PLANES>(y, oplane, (((size_t)b * D + d) * HW + p) * AT_C + q * 8, f);
        } else {
          // every coordinate x in [0,n) maps to x+1, plus the replicated border copies 0 (x==0) and n+1 (x==n-1)
          for (int iz = 0; iz < 3; ++iz) {
            const int zz = iz == 0 ? d + 1 : (iz == 1 ? (d == 0 ? 0 : -1) : (d == D - 1 ? D + 1 : -1));
            if (zz < 0) continue;
            for (int iy = 0; iy < 3; ++iy) {
              const int yy = iy == 0 ? ph + 1 : (iy == 1 ? (ph == 0 ? 0 : -1) : (ph == H - 1 ? H + 1 : -1));
              if (yy < 0) continue;
              for (int ix = 0; ix < 3; ++ix) {
                const int xx = ix == 0 ? pw + 1 : (ix == 1 ? (pw == 0 ? 0 : -1) : (pw == Wd - 1 ? Wd + 1 : -1));
                if (xx < 0) continue;
                store8<PLANES>(y, oplane, ((((size_t)b * Dp2 + zz) * Hp2 + yy) * Wp2 + xx) * AT_C + q * 8, f);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// trilinear x2 (align_corners=False) of t [B,Dl,Hl,Wl,32] + Wc . cost [B,2Dl,2Hl,2Wl,32], then BN.
//   out[v] = scale * (up(t)[v] + Wc cost[v]) + shift
// A warp owns one 2x2x2 output block whose 8 voxels share the same 8 low-res corners
// (block index i in [-1, n-1] per axis -> outputs {2i+1, 2i+2} clipped to [0, 2n)).
// lane = (voxel v of the block, channel octet q): every global access is a 16-byte vector, the four
// lanes of a voxel cover its 64-byte channel row.  The 32x32 matvec reads the block's cost vectors
// (padded rows) and Wc from shared memory with conflict-free LDS.128.
// ------------------------------------------------------------------------------------------------
constexpr int UF_WARPS = 8;

struct UfBlock {            // raw prefetched data of one 2x2x2 block for this lane
  uint4 c_hi, c_lo;         // cost octet of the lane's voxel
  uint4 k_hi, k_lo;         // corner (lane>>2), octet (lane&3) of t
  size_t ooff;
  bool valid;
};

template <int PLANES>
__device__ __forceinline__ void uf_fetch(UfBlock& u, long long blk, const __nv_bfloat16* __restrict__ t,
                                         const __nv_bfloat16* __restrict__ cost, int Dl, int Hl, int Wl, int nbd,
                                         int nbh, int nbw, int lane, size_t lplane, size_t hplane) {
  const int D = 2 * Dl, H = 2 * Hl, W = 2 * Wl;
  long long r = blk;
  const int bw = (int)(r % nbw) - 1; r /= nbw;
  const int bh = (int)(r % nbh) - 1; r /= nbh;
  const int bd = (int)(r % nbd) - 1;
  const int b = (int)(r / nbd);
  const int v = lane >> 2, q = lane & 3;
  const int a = v >> 2, bb = (v >> 1) & 1, c = v & 1;
  const int oz = 2 * bd + 1 + a, oy = 2 * bh + 1 + bb, ox = 2 * bw + 1 + c;
  u.valid = oz >= 0 && oz < D && oy >= 0 && oy < H && ox >= 0 && ox < W;
  u.ooff = ((((size_t)b * D + (u.valid ? oz : 0)) * H + (u.valid ? oy : 0)) * W + (u.valid ? ox : 0)) * AT_C + q * 8;
  u.c_hi = u.c_lo = make_uint4(0, 0, 0, 0);
  if (u.valid) {
    u.c_hi = *reinterpret_cast<const uint4*>(cost + u.ooff);
    if (PLANES == 2) u.c_lo = *reinterpret_cast<const uint4*>(cost + hplane + u.ooff);
  }
  // corner (a,bb,c) of the block, clamped
  const int iz = min(max(bd + a, 0), Dl - 1), iy = min(max(bh + bb, 0), Hl - 1), ix = min(max(bw + c, 0), Wl - 1);
  const size_t koff = ((((size_t)b * Dl + iz) * Hl + iy) * Wl + ix) * AT_C + q * 8;
  u.k_hi = *reinterpret_cast<const uint4*>(t + koff);
  u.k_lo = make_uint4(0, 0, 0, 0);
  if (PLANES == 2) u.k_lo = *reinterpret_cast<const uint4*>(t + lplane + koff);
}

template <int PLANES>
__global__ void __launch_bounds__(UF_WARPS * 32, 2)
upsample_fuse_kernel(const __nv_bfloat16* __restrict__ t, const __nv_bfloat16* __restrict__ cost,
                     const float* __restrict__ WcT, const float* __restrict__ scale, const float* __restrict__ shift,
                     __nv_bfloat16* __restrict__ y, int B, int Dl, int Hl, int Wl) {
  __shared__ __align__(16) float ws[AT_C * AT_C];            // Wc transposed [ci][co]
  __shared__ __align__(16) float cs[UF_WARPS][8][AT_C + 4];  // cost vectors  [warp][voxel][channel] (padded rows)
  __shared__ __align__(16) float kn[UF_WARPS][8][AT_C + 4];  // corner vectors [warp][corner][channel]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = lane >> 2, q = lane & 3;
  for (int i = threadIdx.x; i < AT_C * AT_C; i += blockDim.x) ws[i] = __ldg(WcT + i);
  __shared__ __align__(16) float s_sc[AT_C], s_sh[AT_C];
  if (threadIdx.x < AT_C) { s_sc[threadIdx.x] = __ldg(scale + threadIdx.x); s_sh[threadIdx.x] = __ldg(shift + threadIdx.x); }
  __syncthreads();
  const size_t lplane = (size_t)B * Dl * Hl * Wl * AT_C, hplane = lplane * 8;
  const int nbd = Dl + 1, nbh = Hl + 1, nbw = Wl + 1;
  const long long nblocks = (long long)B * nbd * nbh * nbw;
  const int a = v >> 2, bb = (v >> 1) & 1, c = v & 1;
  // output 2i+1 (first of the pair): .75*in[i] + .25*in[i+1];  output 2i+2: .25*in[i] + .75*in[i+1]
  const float wz0 = a ? 0.25f : 0.75f, wy0 = bb ? 0.25f : 0.75f, wx0 = c ? 0.25f : 0.75f;
  const long long stride = (long long)gridDim.x * UF_WARPS;
  long long blk = (long long)blockIdx.x * UF_WARPS + warp;
  UfBlock cur;
  if (blk < nblocks) uf_fetch<PLANES>(cur, blk, t, cost, Dl, Hl, Wl, nbd, nbh, nbw, lane, lplane, hplane);
  for (; blk < nblocks; blk += stride) {
    UfBlock nxt;
    const bool has_next = blk + stride < nblocks;
    if (has_next) uf_fetch<PLANES>(nxt, blk + stride, t, cost, Dl, Hl, Wl, nbd, nbh, nbw, lane, lplane, hplane);
    // ---- stage this block: cost octet of voxel v, corner octet of corner v ----
    {
      float f[8], g[8];
      unpack8(cur.c_hi, f);
      if (PLANES == 2) { unpack8(cur.c_lo, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += g[j]; }
      float4* crow = reinterpret_cast<float4*>(&cs[warp][v][q * 8]);
      crow[0] = make_float4(f[0], f[1], f[2], f[3]);
      crow[1] = make_float4(f[4], f[5], f[6], f[7]);
      unpack8(cur.k_hi, f);
      if (PLANES == 2) { unpack8(cur.k_lo, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += g[j]; }
      float4* krow = reinterpret_cast<float4*>(&kn[warp][v][q * 8]);
      krow[0] = make_float4(f[0], f[1], f[2], f[3]);
      krow[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncwarp();
    // ---- trilinear blend of the 8 corners for this lane's voxel / octet ----
    float up[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) up[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float wgt = ((k >> 2) ? 1.f - wz0 : wz0) * (((k >> 1) & 1) ? 1.f - wy0 : wy0) * ((k & 1) ? 1.f - wx0 : wx0);
      const float4 k0 = *reinterpret_cast<const float4*>(&kn[warp][k][q * 8]);
      const float4 k1 = *reinterpret_cast<const float4*>(&kn[warp][k][q * 8 + 4]);
      up[0] = fmaf(wgt, k0.x, up[0]); up[1] = fmaf(wgt, k0.y, up[1]); up[2] = fmaf(wgt, k0.z, up[2]);
      up[3] = fmaf(wgt, k0.w, up[3]); up[4] = fmaf(wgt, k1.x, up[4]); up[5] = fmaf(wgt, k1.y, up[5]);
      up[6] = fmaf(wgt, k1.z, up[6]); up[7] = fmaf(wgt, k1.w, up[7]);
    }
    // ---- 32x32 matvec: this lane's 8 output channels of its voxel ----
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < AT_C; c4 += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(&cs[warp][v][c4]);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[(c4 + u) * AT_C + q * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[(c4 + u) * AT_C + q * 8 + 4]);
        acc[0] = fmaf(xs[u], w0.x, acc[0]); acc[1] = fmaf(xs[u], w0.y, acc[1]);
        acc[2] = fmaf(xs[u], w0.z, acc[2]); acc[3] = fmaf(xs[u], w0.w, acc[3]);
        acc[4] = fmaf(xs[u], w1.x, acc[4]); acc[5] = fmaf(xs[u], w1.y, acc[5]);
        acc[6] = fmaf(xs[u], w1.z, acc[6]); acc[7] = fmaf(xs[u], w1.w, acc[7]);
      }
    }
    if (cur.valid) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (up[j] + acc[j]) * s_sc[q * 8 + j] + s_sh[q * 8 + j];
      store8<PLANES>(y, hplane, cur.ooff, o);
    }
    __syncwarp();
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------
// softmax over disparity fused with soft-argmin regression; the probability volume never exists.
// logits fp32 [B,D,H,W] -> pred [B,H,W].  thread = pixel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_regress_kernel(const float* __restrict__ logits, float* __restrict__ pred, int D, int HW) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* lp = logits + (size_t)b * D * HW + p;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, lp[(size_t)d * HW]);
  float s = 0.f, acc = 0.f;
  for (int d = 0; d < D; ++d) {
    const float ev = expf(lp[(size_t)d * HW] - m);
    s += ev;
    acc = fmaf(ev, (float)d, acc);
  }
  pred[(size_t)b * HW + p] = acc / s;
}

// ------------------------------------------------------------------------------------------------
// convex 4x upsampling.  mask fp32 channels-last [B,H,W,144] (c = n*16 + i*4 + j), disp [B,H,W],
// out [B,1,4H,4W].  thread = (pixel, sub-row i): softmax over the 9 neighbours for its 4 columns.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
convex_upsample_kernel(const float* __restrict__ mask, const float* __restrict__ disp, float* __restrict__ out, int B,
                       int H, int W) {
  const size_t total = (size_t)B * H * W * 4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx & 3);
    size_t pix = idx >> 2;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((size_t)W * H));
    const float* mp = mask + pix * 144 + i * 4;
    float4 mv[9];
    float dv[9];
#pragma unroll
    for (int n = 0; n < 9; ++n) {
      mv[n] = *reinterpret_cast<const float4*>(mp + n * 16);
      const int hh = h + n / 3 - 1, ww = w + n % 3 - 1;
      dv[n] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? 4.f * __ldg(disp + ((size_t)b * H + hh) * W + ww) : 0.f;
    }
    float4 mx = mv[0];
#pragma unroll
    for (int n = 1; n < 9; ++n) {
      mx.x = fmaxf(mx.x, mv[n].x); mx.y = fmaxf(mx.y, mv[n].y); mx.z = fmaxf(mx.z, mv[n].z); mx.w = fmaxf(mx.w, mv[n].w);
    }
    float4 s = make_float4(0, 0, 0, 0), a = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int n = 0; n < 9; ++n) {
      const float ex = expf(mv[n].x - mx.x), ey = expf(mv[n].y - mx.y), ez = expf(mv[n].z - mx.z),
                  ew = expf(mv[n].w - mx.w);
      s.x += ex; s.y += ey; s.z += ez; s.w += ew;
      a.x = fmaf(ex, dv[n], a.x); a.y = fmaf(ey, dv[n], a.y); a.z = fmaf(ez, dv[n], a.z); a.w = fmaf(ew, dv[n], a.w);
    }
    float4 o = make_float4(a.x / s.x, a.y / s.y, a.z / s.z, a.w / s.w);
    *reinterpret_cast<float4*>(out + ((size_t)b * 4 * H + 4 * h + i) * (4 * W) + 4 * w) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// layout conversion between the reference's fp32 NCDHW tensors and cost planes
// ------------------------------------------------------------------------------------------------
template <int PLANES>
__global__ void __launch_bounds__(256)
planes_from_ncdhw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int C, int Cp,
                         size_t V) {
  // V = D*H*W voxels per batch item; thread = (b, c8, voxel) with voxel fastest (coalesced reads)
  const int c8n = Cp / 8;
  const size_t total = (size_t)B * c8n * V;
  const size_t plane = (size_t)B * V * Cp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % V;
    const int c8 = (int)((i / V) % c8n);
    const int b = (int)(i / (V * c8n));
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      f[k] = (c < C) ? __ldg(x + ((size_t)b * C + c) * V + v) : 0.f;
    }
    store8<PLANES>(y, plane, ((size_t)b * V + v) * Cp + c8 * 8, f);
  }
}

__global__ void __launch_bounds__(256)
planes_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ x, int planes, float* __restrict__ y, int B, int C, int Cp,
                       size_t V) {
  const size_t total = (size_t)B * C * V;
  const size_t plane = (size_t)B * V * Cp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % V;
    const int c = (int)((i / V) % C);
    const int b = (int)(i / (V * C));
    const size_t off = ((size_t)b * V + v) * Cp + c;
    float r = __bfloat162float(x[off]);
    if (planes == 2) r += __bfloat162float(x[plane + off]);
    y[i] = r;
  }
}

// weights: torch Conv [Co][Ci][taps] or ConvTranspose [Ci][Co][taps] -> [taps][Ci][CoPad] fp32 (zero pad)
__global__ void pack_weight_kernel(const float* __restrict__ w, int transposed, int Co, int Ci, int taps,
                                   float* __restrict__ out, int CoPad) {
  const int total = taps * Ci * CoPad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % CoPad, ci = (i / CoPad) % Ci, t = i / (CoPad * Ci);
    float v = 0.f;
    if (co < Co) v = transposed ? w[((size_t)ci * Co + co) * taps + t] : w[((size_t)co * Ci + ci) * taps + t];
    out[i] = v;
  }
}

// eval-mode BatchNorm -> per-channel scale/shift (length Cpad, identity padding)
__global__ void fold_bn_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C, int Cpad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cpad) return;
  if (i < C) {
    const float s = gamma[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = beta[i] - mean[i] * s;
  } else {
    scale[i] = 1.f; shift[i] = 0.f;
  }
}

static inline int grid_for(size_t total, int threads) {
  size_t g = (total + threads - 1) / threads;
  const size_t cap = 148 * 32;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace dca

using namespace dca;

extern "C" int dca_avgpool3d(const void* x, void* y, int planes, int B, int C, int Di, int Hi, int Wi, void* stream) {
  if (!x || !y || C % 8 != 0 || B <= 0 || Di <= 0 || Hi <= 0 || Wi <= 0 || planes < 1 || planes > 2)
    return DCA_ERR_ARG;
  const int Do = (Di + 1) / 2, Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  const size_t total = (size_t)B * Do * Ho * Wo * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (planes == 2)
    avgpool3d_kernel<2><<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, C, Di, Hi,
                                                               Wi, Do, Ho, Wo);
  else
    avgpool3d_kernel<1><<<grid_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, C, Di, Hi,
                                                               Wi, Do, Ho, Wo);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_class_stats(const float* logits, int* cls, float* e, float* S, int B, int D, int H, int W,
                               void* stream) {
  if (!logits || !cls || !e || !S || B <= 0 || D <= 0 || D > CS_MAXD || H <= 0 || W <= 0) return DCA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(S, 0, (size_t)B * D * sizeof(float), st) != cudaSuccess) return DCA_ERR_LAUNCH;
  const int HW = H * W;
  class_stats_kernel<<<dim3((HW + 255) / 256, B), 256, 0, st>>>(logits, cls, e, S, D, HW);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_disp_attention(const void* x, const int* cls, const float* e, const float* S, const float* weights,
                                  int has_wa, void* y, int pad, int planes, int B, int C, int D, int H, int W,
                                  void* stream) {
  if (pad != 0 && pad != 1) return DCA_ERR_ARG;
  if (!x || !cls || !e || !S || !weights || !y || planes < 1 || planes > 2 || B <= 0 || D <= 0) return DCA_ERR_ARG;
  if (C != AT_C) return DCA_ERR_UNSUPPORTED;
  const int Dp = (D + 7) & ~7;
  int warps = 8;
  size_t wbytes = (size_t)((AT_WFLOATS + 3) & ~3) * sizeof(float);
  while (warps > 1 && wbytes + (size_t)warps * 4 * Dp * AT_C * sizeof(float) > 227 * 1024) warps >>= 1;
  const size_t smem = wbytes + (size_t)warps * 4 * Dp * AT_C * sizeof(float);
  if (smem > 227 * 1024) return DCA_ERR_UNSUPPORTED;
  const int HW = H * W;
  int grid = (B * HW + warps - 1) / warps;
  if (grid > 148 * 4) grid = 148 * 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (planes == 2) {
    cudaFuncSetAttribute(disp_attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    disp_attention_kernel<2><<<grid, warps * 32, smem, st>>>((const __nv_bfloat16*)x, cls, e, S, weights, has_wa,
                                                             (__nv_bfloat16*)y, B, D, H, W, pad);
  } else {
    cudaFuncSetAttribute(disp_attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    disp_attention_kernel<1><<<grid, warps * 32, smem, st>>>((const __nv_bfloat16*)x, cls, e, S, weights, has_wa,
                                                             (__nv_bfloat16*)y, B, D, H, W, pad);
  }
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_upsample_fuse(const void* t, const void* cost, const float* WcT, const float* scale,
                                 const float* shift, void* y, int planes, int B, int C, int Dl, int Hl, int Wl,
                                 void* stream) {
  if (!t || !cost || !WcT || !scale || !shift || !y || planes < 1 || planes > 2 || B <= 0) return DCA_ERR_ARG;
  if (C != AT_C) return DCA_ERR_UNSUPPORTED;
  const long long nblocks = (long long)B * (Dl + 1) * (Hl + 1) * (Wl + 1);
  int grid = (int)((nblocks + 7) / 8 < 148 * 12 ? (nblocks + 7) / 8 : 148 * 12);
  cudaStream_t st = (cudaStream_t)stream;
  if (planes == 2)
    upsample_fuse_kernel<2><<<grid, 256, 0, st>>>((const __nv_bfloat16*)t, (const __nv_bfloat16*)cost, WcT, scale,
                                                  shift, (__nv_bfloat16*)y, B, Dl, Hl, Wl);
  else
    upsample_fuse_kernel<1><<<grid, 256, 0, st>>>((const __nv_bfloat16*)t, (const __nv_bfloat16*)cost, WcT, scale,
                                                  shift, (__nv_bfloat16*)y, B, Dl, Hl, Wl);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_softmax_regress(const float* logits, float* pred, int B, int D, int H, int W, void* stream) {
  if (!logits || !pred || B <= 0 || D <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  const int HW = H * W;
  softmax_regress_kernel<<<dim3((HW + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(logits, pred, D, HW);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_convex_upsample(const float* mask, const float* disp, float* out, int B, int H, int W,
                                   void* stream) {
  if (!mask || !disp || !out || B <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  const size_t total = (size_t)B * H * W * 4;
  convex_upsample_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mask, disp, out, B, H, W);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_planes_from_ncdhw(const float* x, void* y, int planes, int B, int C, int Cp, int D, int H, int W,
                                     void* stream) {
  if (!x || !y || planes < 1 || planes > 2 || B <= 0 || C <= 0 || Cp < C || Cp % 8 != 0) return DCA_ERR_ARG;
  const size_t V = (size_t)D * H * W, total = (size_t)B * (Cp / 8) * V;
  cudaStream_t st = (cudaStream_t)stream;
  if (planes == 2)
    planes_from_ncdhw_kernel<2><<<grid_for(total, 256), 256, 0, st>>>(x, (__nv_bfloat16*)y, B, C, Cp, V);
  else
    planes_from_ncdhw_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(x, (__nv_bfloat16*)y, B, C, Cp, V);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_planes_to_ncdhw(const void* x, int planes, float* y, int B, int C, int Cp, int D, int H, int W,
                                   void* stream) {
  if (!x || !y || planes < 1 || planes > 2 || B <= 0 || C <= 0 || Cp < C) return DCA_ERR_ARG;
  const size_t V = (size_t)D * H * W, total = (size_t)B * C * V;
  planes_to_ncdhw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, planes, y, B,
                                                                                 C, Cp, V);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_pack_weights(const float* w, int transposed, int Co, int Ci, int taps, float* out, int CoPad,
                                void* stream) {
  if (!w || !out || Co <= 0 || Ci <= 0 || taps <= 0 || CoPad < Co) return DCA_ERR_ARG;
  pack_weight_kernel<<<grid_for((size_t)taps * Ci * CoPad, 256), 256, 0, (cudaStream_t)stream>>>(w, transposed, Co, Ci,
                                                                                                 taps, out, CoPad);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                           float* scale, float* shift, int C, int Cpad, void* stream) {
  if (!gamma || !beta || !mean || !var || !scale || !shift || C <= 0 || Cpad < C) return DCA_ERR_ARG;
  fold_bn_kernel<<<(Cpad + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, scale, shift, C,
                                                                       Cpad);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_version(void) { return 100; }
