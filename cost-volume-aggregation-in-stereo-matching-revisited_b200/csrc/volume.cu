// Fused group-wise-correlation + concat cost volume (SURVEY 8a rows a1+a2).
// Replaces the reference's two Python loops over disparities
//   build_gwc_volume     /root/reference/models/submodule.py:157-167 (groupwise_correlation :148-154)
//   build_concat_volume  /root/reference/models/submodule.py:134-145
// and the torch.cat at models/gwcnet_dca_g.py:220.
//
// HBM-bound: reads 2 x (C + Cc) x H x W fp32 once (re-reads of the right-image window hit L2),
// writes [B][D][H][W][Cv] bf16 (x planes) once with sector-complete stores.
//
// Fused kernel layout
//   CTA  = one (b, h) row, TW=16 columns, one chunk of DC<=48 disparities; 256 threads.
//   smem = left tile  Ls[c][w][g]  (c = channel inside its group, g = group, pitch G+1)
//          right tile Rs[c][u][g]  for the u = w-d window (zero-filled outside the image)
//          concat tiles (channel-major).
//   warp = 4 columns x 8 disparities, lane = output channel slot (0..31 then 32..63):
//          slot < G  -> group correlation (register tile 4w x 8d: 15 LDS per 32 FMA)
//          G <= slot < G+Cc -> left concat copy, < G+2Cc -> right concat copy, rest zero pad.
//   The 32 fp32 results of a lane are bf16-split, pair-exchanged by shuffle and stored as 32-bit
//   words so that a warp store covers whole 32-byte sectors of two voxels.
#include "dca_common.cuh"

namespace dca {

constexpr int VOL_TW = 16;   // columns per CTA
constexpr int VOL_DC = 48;   // disparities per CTA
constexpr int VOL_THREADS = 256;

// 4-byte async global->shared copy (LDGSTS); src_size 0 zero-fills the destination (out-of-image columns)
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(valid ? 4 : 0)
               : "memory");
}

__device__ __forceinline__ void vol_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// CVT = compile-time channel pitch of the volume (64 for DCANet's 40 groups + 2x12 concat); 0 = runtime Cv.
// GT / CPGT = compile-time group count / channels per group (40 / 8 for DCANet); 0 = runtime.  With them fixed
// every shared-memory offset of the inner product loop is an immediate and the channel loop unrolls fully.
template <int PLANES, int CVT, int GT, int CPGT>
__global__ void __launch_bounds__(VOL_THREADS, 2)
volume_fused_kernel(const float* __restrict__ gl, const float* __restrict__ gr,
                    const float* __restrict__ cl, const float* __restrict__ cr,
                    __nv_bfloat16* __restrict__ vol, int B, int C, int G_rt, int Cc, int D, int H, int W, int Cv_rt) {
  extern __shared__ float smem[];
  const int Cv = CVT ? CVT : Cv_rt;
  const int G = GT ? GT : G_rt;
  const int cpg = CPGT ? CPGT : C / G;
  const int gp = G + 1;                       // pitch of the group axis (bank-conflict-free transposing stores)
  constexpr int UW = VOL_TW + VOL_DC - 1;     // right window width (63)
  float* Ls = smem;                           // [cpg][TW][gp]
  float* Rs = Ls + cpg * VOL_TW * gp;         // [cpg][UW][gp]
  float* cLs = Rs + cpg * UW * gp;            // [Cc][TW]
  float* cRs = cLs + Cc * VOL_TW;             // [Cc][UW]

  const int wtiles = (W + VOL_TW - 1) / VOL_TW;
  const int w0 = (blockIdx.x % wtiles) * VOL_TW;
  const int dchunk = blockIdx.x / wtiles;
  const int d0 = dchunk * VOL_DC;
  const int h = blockIdx.y, b = blockIdx.z;
  const int dcount = min(VOL_DC, D - d0);
  const int u0 = w0 - (d0 + VOL_DC - 1);      // image column of window index 0
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const size_t HW = (size_t)H * W;

  // ---- stage: coalesced along w in global, transposed into [c][w][g] in smem (no div/mod in the loops) ----
  {
    // all copies are asynchronous (no register round trip): every thread has ~100 loads in flight at once
    const float* src = gl + ((size_t)b * C * H + h) * W;   // + ch*HW + w
    for (int i = tid; i < C * VOL_TW; i += VOL_THREADS) {
      const int ch = i / VOL_TW, wl = i % VOL_TW, w = w0 + wl;
      const bool ok = w < W;
      cp_async4(&Ls[((ch % cpg) * VOL_TW + wl) * gp + ch / cpg], ok ? src + (size_t)ch * HW + w : src, ok);
    }
    const float* srcr = gr + ((size_t)b * C * H + h) * W;
    for (int i = tid; i < C * UW; i += VOL_THREADS) {
      const int ch = i / UW, ul = i % UW, u = u0 + ul;
      const bool ok = u >= 0 && u < W;
      cp_async4(&Rs[((ch % cpg) * UW + ul) * gp + ch / cpg], ok ? srcr + (size_t)ch * HW + u : srcr, ok);
    }
    if (Cc > 0) {
      const float* s2 = cl + ((size_t)b * Cc * H + h) * W;
      for (int i = tid; i < Cc * VOL_TW; i += VOL_THREADS) {
        const int ch = i / VOL_TW, wl2 = i % VOL_TW, w = w0 + wl2;
        const bool ok = w < W;
        cp_async4(&cLs[i], ok ? s2 + (size_t)ch * HW + w : s2, ok);
      }
      const float* s3 = cr + ((size_t)b * Cc * H + h) * W;
      for (int i = tid; i < Cc * UW; i += VOL_THREADS) {
        const int ch = i / UW, ul = i % UW, u = u0 + ul;
        const bool ok = u >= 0 && u < W;
        cp_async4(&cRs[i], ok ? s3 + (size_t)ch * HW + u : s3, ok);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();

  const float inv_cpg = 1.0f / (float)cpg;
  const size_t plane_stride = (size_t)B * D * H * W * Cv;
  const int n_wt = VOL_TW / 4, n_dt = (dcount + 7) / 8;
  const int odd = lane & 1;

  for (int t = warp; t < n_wt * n_dt; t += VOL_THREADS / 32) {
    const int wt = (t % n_wt) * 4;            // local column of the 4-wide tile
    const int dt = (t / n_wt) * 8;            // local disparity of the 8-deep tile
    // window index of (w = w0+wt+i, d = d0+dt+j):  ul = wt + i - dt - j + (VOL_DC-1)
    const int ubase = wt - dt + (VOL_DC - 1) - 7;   // ul = ubase + (i - j + 7), i-j+7 in [0,10]
    for (int pass = 0; pass < 2; ++pass) {
      const int slot = lane + 32 * pass;
      if (32 * pass >= Cv) break;
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

      if (slot < G) {
#pragma unroll
        for (int c = 0; c < (CPGT ? CPGT : cpg); ++c) {
          float l[4], r[11];
          const float* lp = Ls + (c * VOL_TW + wt) * gp + slot;
          const float* rp = Rs + (c * UW + ubase) * gp + slot;
#pragma unroll
          for (int i = 0; i < 4; ++i) l[i] = lp[i * gp];
#pragma unroll
          for (int k = 0; k < 11; ++k) r[k] = rp[k * gp];
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(l[i], r[i - j + 7], acc[j][i]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j][i] *= inv_cpg;
      } else if (slot < G + Cc) {
        const int cc = slot - G;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc[j][i] = (w0 + wt + i >= d0 + dt + j) ? cLs[cc * VOL_TW + wt + i] : 0.f;
      } else if (slot < G + 2 * Cc) {
        const int cc = slot - G - Cc;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j][i] = cRs[cc * UW + ubase + (i - j + 7)];
      }
      // ---- pair exchange (even lane keeps column ip, odd lane column ip+1), packed bf16 split, 32-bit stores ----
      const int ch = slot & ~1;
      const bool ch_ok = ch < Cv;
      // this lane's column inside each pair is ip + odd
      __nv_bfloat16* colp = vol + ((((size_t)b * D + d0 + dt) * H + h) * W + (w0 + wt + odd)) * Cv + ch;
      const size_t dstep = HW * Cv;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool d_ok = (dt + j) < dcount;
#pragma unroll
        for (int ip = 0; ip < 4; ip += 2) {
          const float mine_a = acc[j][ip], mine_b = acc[j][ip + 1];
          const float got = __shfl_xor_sync(0xffffffffu, odd ? mine_a : mine_b, 1);
          // even lane: voxel a, channels (slot, slot+1) = (mine_a, got); odd lane: voxel b, (slot-1, slot) = (got, mine_b)
          const float lo_ch = odd ? got : mine_a, hi_ch = odd ? mine_b : got;
          if (d_ok && ch_ok && (w0 + wt + ip + odd) < W) {
            uint32_t hw, lw;
            vol_split2(lo_ch, hi_ch, hw, lw);
            __nv_bfloat16* dst = colp + (size_t)j * dstep + ip * Cv;
            *reinterpret_cast<uint32_t*>(dst) = hw;
            if (PLANES == 2) *reinterpret_cast<uint32_t*>(dst + plane_stride) = lw;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Reference-API variants: fp32 NCDHW outputs, exactly the tensors build_gwc_volume /
// build_concat_volume return.  CTA = one (b, group, h) row staged in smem, lane = column,
// every thread walks the disparities; stores are coalesced along w.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gwc_volume_ncdhw_kernel(const float* __restrict__ gl, const float* __restrict__ gr, float* __restrict__ vol,
                        int B, int C, int G, int D, int H, int W) {
  extern __shared__ float smem[];
  const int cpg = C / G;
  float* Ls = smem;            // [cpg][W]
  float* Rs = smem + cpg * W;  // [cpg][W]
  const int h = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const float* lsrc = gl + ((size_t)(b * C + g * cpg) * H + h) * W;
  const float* rsrc = gr + ((size_t)(b * C + g * cpg) * H + h) * W;
  for (int i = threadIdx.x; i < cpg * W; i += blockDim.x) {
    int c = i / W, w = i % W;
    Ls[i] = __ldg(lsrc + (size_t)c * HW + w);
    Rs[i] = __ldg(rsrc + (size_t)c * HW + w);
  }
  __syncthreads();
  const float inv = 1.0f / (float)cpg;
  for (int i = threadIdx.x; i < D * W; i += blockDim.x) {
    int d = i / W, w = i % W;
    float acc = 0.f;
    if (w >= d) {
      for (int c = 0; c < cpg; ++c) acc = fmaf(Ls[c * W + w], Rs[c * W + w - d], acc);
      acc *= inv;
    }
    vol[((((size_t)b * G + g) * D + d) * H + h) * W + w] = acc;
  }
}

__global__ void __launch_bounds__(256)
concat_volume_ncdhw_kernel(const float* __restrict__ cl, const float* __restrict__ cr, float* __restrict__ vol,
                           int B, int C, int D, int H, int W) {
  // one thread per output element of the [B,2C,D,H,W] volume, w fastest
  const size_t total = (size_t)B * 2 * C * D * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    size_t r = i / W;
    int h = (int)(r % H); r /= H;
    int d = (int)(r % D); r /= D;
    int c2 = (int)(r % (2 * C));
    int b = (int)(r / (2 * C));
    float v = 0.f;
    if (w >= d) {
      if (c2 < C) v = __ldg(cl + (((size_t)b * C + c2) * H + h) * W + w);
      else v = __ldg(cr + (((size_t)b * C + (c2 - C)) * H + h) * W + w - d);
    }
    vol[i] = v;
  }
}

}  // namespace dca

using namespace dca;

extern "C" int dca_volume_gwc_concat(const float* gwc_l, const float* gwc_r, const float* cat_l, const float* cat_r,
                                     void* vol, int B, int C, int G, int Cc, int D, int H, int W, int Cv, int planes,
                                     void* stream) {
  if (!gwc_l || !gwc_r || !vol || B <= 0 || C <= 0 || G <= 0 || D <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  if (C % G != 0 || (Cc > 0 && (!cat_l || !cat_r)) || Cc < 0) return DCA_ERR_ARG;
  if (Cv < G + 2 * Cc || Cv > 64 || (Cv % 8) != 0 || (planes != 1 && planes != 2)) return DCA_ERR_ARG;
  const int cpg = C / G, gp = G + 1, UW = VOL_TW + VOL_DC - 1;
  size_t smem = ((size_t)cpg * (VOL_TW + UW) * gp + (size_t)Cc * (VOL_TW + UW)) * sizeof(float);
  if (smem > 220 * 1024) return DCA_ERR_UNSUPPORTED;
  const int wtiles = (W + VOL_TW - 1) / VOL_TW, dchunks = (D + VOL_DC - 1) / VOL_DC;
  dim3 grid(wtiles * dchunks, H, B);
  cudaStream_t st = (cudaStream_t)stream;
#define DCA_VOL_LAUNCH(P, CVT)                                                                                   \
  do {                                                                                                           \
    auto kern = volume_fused_kernel<P, CVT, GT_, CPG_>;                                                          \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                          \
    kern<<<grid, VOL_THREADS, smem, st>>>(gwc_l, gwc_r, cat_l, cat_r, (__nv_bfloat16*)vol, B, C, G, Cc, D, H, W, Cv); \
  } while (0)
  const bool dcanet_shape = (Cv == 64 && G == 40 && cpg == 8);
#define GT_ 40
#define CPG_ 8
  if (dcanet_shape) { if (planes == 2) DCA_VOL_LAUNCH(2, 64); else DCA_VOL_LAUNCH(1, 64); }
#undef GT_
#undef CPG_
#define GT_ 0
#define CPG_ 0
  else { if (planes == 2) DCA_VOL_LAUNCH(2, 0); else DCA_VOL_LAUNCH(1, 0); }
#undef GT_
#undef CPG_
#undef DCA_VOL_LAUNCH
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_build_gwc_volume_f32(const float* l, const float* r, float* vol, int B, int C, int G, int D, int H,
                                        int W, void* stream) {
  if (!l || !r || !vol || B <= 0 || C <= 0 || G <= 0 || C % G != 0 || D <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  size_t smem = (size_t)2 * (C / G) * W * sizeof(float);
  if (smem > 200 * 1024) return DCA_ERR_UNSUPPORTED;
  cudaFuncSetAttribute(gwc_volume_ncdhw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gwc_volume_ncdhw_kernel<<<dim3(H, G, B), 256, smem, (cudaStream_t)stream>>>(l, r, vol, B, C, G, D, H, W);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

extern "C" int dca_build_concat_volume_f32(const float* l, const float* r, float* vol, int B, int C, int D, int H, int W,
                                           void* stream) {
  if (!l || !r || !vol || B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  size_t total = (size_t)B * 2 * C * D * H * W;
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  concat_volume_ncdhw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(l, r, vol, B, C, D, H, W);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}
