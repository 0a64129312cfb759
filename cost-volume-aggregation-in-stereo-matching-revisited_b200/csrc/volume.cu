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

template <int PLANES>
__global__ void __launch_bounds__(VOL_THREADS, 2)
volume_fused_kernel(const float* __restrict__ gl, const float* __restrict__ gr,
                    const float* __restrict__ cl, const float* __restrict__ cr,
                    __nv_bfloat16* __restrict__ vol, int B, int C, int G, int Cc, int D, int H, int W, int Cv) {
  extern __shared__ float smem[];
  const int cpg = C / G;
  const int gp = G + 1;                       // pitch of the group axis (bank-conflict-free transposing stores)
  const int UW = VOL_TW + VOL_DC - 1;         // right window width
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
  const size_t HW = (size_t)H * W;

  // ---- stage: coalesced along w in global, transposed into [c][w][g] in smem ----
  {
    const float* src = gl + ((size_t)b * C * H + h) * W;   // + ch*HW + w
    for (int i = tid; i < C * VOL_TW; i += VOL_THREADS) {
      int ch = i / VOL_TW, wl = i % VOL_TW, w = w0 + wl;
      float v = (w < W) ? __ldg(src + (size_t)ch * HW + w) : 0.f;
      Ls[((ch % cpg) * VOL_TW + wl) * gp + ch / cpg] = v;
    }
    const float* srcr = gr + ((size_t)b * C * H + h) * W;
    for (int i = tid; i < C * UW; i += VOL_THREADS) {
      int ch = i / UW, ul = i % UW, u = u0 + ul;
      float v = (u >= 0 && u < W) ? __ldg(srcr + (size_t)ch * HW + u) : 0.f;
      Rs[((ch % cpg) * UW + ul) * gp + ch / cpg] = v;
    }
    if (Cc > 0) {
      const float* s2 = cl + ((size_t)b * Cc * H + h) * W;
      for (int i = tid; i < Cc * VOL_TW; i += VOL_THREADS) {
        int ch = i / VOL_TW, wl = i % VOL_TW, w = w0 + wl;
        cLs[i] = (w < W) ? __ldg(s2 + (size_t)ch * HW + w) : 0.f;
      }
      const float* s3 = cr + ((size_t)b * Cc * H + h) * W;
      for (int i = tid; i < Cc * UW; i += VOL_THREADS) {
        int ch = i / UW, ul = i % UW, u = u0 + ul;
        cRs[i] = (u >= 0 && u < W) ? __ldg(s3 + (size_t)ch * HW + u) : 0.f;
      }
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const float inv_cpg = 1.0f / (float)cpg;
  const size_t plane_stride = (size_t)B * D * H * W * Cv;
  const int n_wt = VOL_TW / 4, n_dt = (dcount + 7) / 8;

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
        for (int c = 0; c < cpg; ++c) {
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
      // ---- bf16 split, pair exchange, 32-bit stores: even lane stores column i of the pair
      //      (i, i+1) ... we pair voxels (j, i) and (j, i+1) ----
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int ip = 0; ip < 4; ip += 2) {
          float mine_a = acc[j][ip], mine_b = acc[j][ip + 1];
          float send = (lane & 1) ? mine_a : mine_b;          // odd lanes give away voxel a, even give b
          float got = __shfl_xor_sync(0xffffffffu, send, 1);
          // even lane: voxel a, channels (slot, slot+1) = (mine_a, got)
          // odd  lane: voxel b, channels (slot-1, slot) = (got, mine_b)
          float lo_ch = (lane & 1) ? got : mine_a;
          float hi_ch = (lane & 1) ? mine_b : got;
          const int i = ip + (lane & 1);
          const int w = w0 + wt + i, d = d0 + dt + j;
          const int ch = slot & ~1;
          if (w < W && d < D && (dt + j) < dcount && ch < Cv) {
            uint32_t l0, l1;
            uint32_t h0 = split_bf16(lo_ch, l0), h1 = split_bf16(hi_ch, l1);
            size_t off = ((((size_t)b * D + d) * H + h) * W + w) * Cv + ch;
            *reinterpret_cast<uint32_t*>(vol + off) = h0 | (h1 << 16);
            if (PLANES == 2) *reinterpret_cast<uint32_t*>(vol + plane_stride + off) = l0 | (l1 << 16);
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
  if (planes == 2) {
    cudaFuncSetAttribute(volume_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    volume_fused_kernel<2><<<grid, VOL_THREADS, smem, st>>>(gwc_l, gwc_r, cat_l, cat_r, (__nv_bfloat16*)vol, B, C, G,
                                                             Cc, D, H, W, Cv);
  } else {
    cudaFuncSetAttribute(volume_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    volume_fused_kernel<1><<<grid, VOL_THREADS, smem, st>>>(gwc_l, gwc_r, cat_l, cat_r, (__nv_bfloat16*)vol, B, C, G,
                                                             Cc, D, H, W, Cv);
  }
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
