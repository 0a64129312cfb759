// tcgen05 / TMEM / TMA implicit-GEMM 3-D convolution for sm_100a.
//
// Replaces nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d + ReLU + residual adds of the reference
// (/root/reference/models/submodule.py:121-124, models/gwcnet_dca_g.py:141-148,166-168,224-225,
//  models/augment/cva.py:16-29,39-53) for the heavy layers.
//
// GEMM view:  D[M=128 voxels, N] += A[128, K=Cin] * B[N, K]^T   once per filter tap
//   A  = one TMA box [Cin, 8 w, 16 h, 1 d, 1 b] of the channels-last bf16 cost planes, shifted by the
//        tap offset; out-of-tensor coordinates are zero-filled by TMA (= the conv's zero padding).
//        The box lands as 128 K-major rows (row = h*8+w) in the 64B/128B swizzled UMMA layout.
//   B  = that tap's weights, K-major [N rows][Cin], streamed by TMA from a pre-packed bf16 table.
//   D  = fp32 accumulator in TMEM (double buffered so the epilogue overlaps the next tile's MMAs).
// Precision (SURVEY 7 hard part 2):
//   planes == 2 ("parity"): x = hi + lo, W = Whi + Wlo (all bf16).  Per K-step
//        MMA#1  A=hi, B=[Whi;Wlo] (N = 2*Cout)  -> cols [0,Cout) += hi*Whi, cols [Cout,2Cout) += hi*Wlo
//        MMA#2  A=lo, B=[Whi]     (N = Cout)    -> cols [0,Cout) += lo*Whi
//     and the epilogue adds the two column halves: ~16-bit operand significand, fp32 accumulate.
//   planes == 1 ("fast"): one MMA, plain bf16 operands.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
//   warps 2-5 = epilogue (TMEM -> registers -> BN scale/shift, residuals, activation -> bf16 planes).
// Persistent CTAs (grid = #SMs) walk output tiles round-robin.
#include <cuda.h>
#include <cstring>

#include "dca_common.cuh"

namespace dca {

constexpr int TC_TW = 8, TC_TH = 16, TC_M = 128;
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_TAPS = 27;

struct TcMaps {
  CUtensorMap a[8];   // input views (index = parity class for stride-2 convs, else only [0])
  CUtensorMap w;      // packed weights [taps*planes*Cout rows][Cin]
};

struct TcParams {
  int B, Do, Ho, Wo;
  int Dt, Ht, Wt;                 // tile-space extent
  int tiles_w, tiles_h;
  int out_stride, out_off[3];
  int ntaps;
  signed char tap_off[TC_MAX_TAPS][3];
  signed char tap_map[TC_MAX_TAPS];
  signed char tap_w[TC_MAX_TAPS];
  const float* scale; const float* shift;
  const __nv_bfloat16* res_pre; const __nv_bfloat16* res_post; size_t res_plane; int planes_res;
  __nv_bfloat16* y; size_t y_plane; int planes_out; int act;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major smem matrix descriptor: rows of ROWB bytes (64 -> SWIZZLE_64B, 128 -> SWIZZLE_128B), 8-row groups
// ROWB*8 bytes apart.  Bit layout: cute::UMMA::SmemDescriptor (start>>4 @0, LBO>>4 @16, SBO>>4 @32,
// version=1 @46, layout_type @61).
template <int ROWB>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  constexpr uint64_t layout = (ROWB == 128) ? 2ull : 4ull;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((ROWB * 8) >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}
// instruction descriptor, kind::f16: D=f32 (bit4), A=B=bf16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int CIN, int COUT, int PLANES>
struct TcCfg {
  static constexpr int ROWB = CIN * 2;                       // bytes per K-major row
  static constexpr int A_BYTES = TC_M * ROWB;                // one plane of the activation tile
  static constexpr int B_ROWS = PLANES * COUT;
  static constexpr int B_BYTES = B_ROWS * ROWB;
  static constexpr int STAGE_BYTES = PLANES * A_BYTES + B_BYTES;
  static constexpr int STAGES = (196 * 1024 / STAGE_BYTES) > 8 ? 8 : (196 * 1024 / STAGE_BYTES);
  static constexpr int NACC = PLANES * COUT;                 // TMEM columns per accumulator buffer
  static constexpr int TMEM_COLS = 2 * NACC < 32 ? 32 : 2 * NACC;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * COUT * 4;
};

template <int CIN, int COUT, int PLANES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  using Cfg = TcCfg<CIN, COUT, PLANES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                          // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;           // [STAGES]
  uint64_t* tfull = bars + 2 * Cfg::STAGES;       // [2]
  uint64_t* tempty = tfull + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.Dt * p.tiles_h * p.tiles_w;

  if (threadIdx.x < COUT) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.w);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int td = r % p.Dt;
        const int b = r / p.Dt;
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* st = stage_base + (size_t)s * Cfg::STAGE_BYTES;
          const CUtensorMap* am = &maps.a[p.tap_map[t]];
          const int cw = tw * TC_TW + p.tap_off[t][2], chh = th * TC_TH + p.tap_off[t][1], cd = td + p.tap_off[t][0];
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_5d(st + pl * Cfg::A_BYTES, am, &full[s], 0, cw, chh, cd, pl * p.B + b);
          tma_load_2d(st + PLANES * Cfg::A_BYTES, &maps.w, &full[s], 0, p.tap_w[t] * Cfg::B_ROWS);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_full = make_idesc(TC_M, Cfg::NACC);
      constexpr uint32_t idesc_half = make_idesc(TC_M, COUT);
      uint32_t s = 0, ph = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * Cfg::NACC;
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(stage_base + (size_t)s * Cfg::STAGE_BYTES);
          const uint32_t b0 = a0 + PLANES * Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k) {
            const uint64_t da = make_desc<Cfg::ROWB>(a0 + k * 32);
            const uint64_t db = make_desc<Cfg::ROWB>(b0 + k * 32);
            umma_bf16(d_addr, da, db, idesc_full, (t > 0 || k > 0) ? 1u : 0u);
            if (PLANES == 2) {
              const uint64_t dl = make_desc<Cfg::ROWB>(a0 + Cfg::A_BYTES + k * 32);
              umma_bf16(d_addr, dl, db, idesc_half, 1u);
            }
          }
          umma_commit(&empty[s]);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are reachable from this warp
    const int row = quarter * 32 + lane;          // tile row = voxel
    const int hh = row / TC_TW, ww = row % TC_TW;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int r = tile;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; r /= p.tiles_h;
      const int td = r % p.Dt;
      const int b = r / p.Dt;
      const uint32_t acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * Cfg::NACC;
      const int tz = td, ty = th * TC_TH + hh, tx = tw * TC_TW + ww;
      const int oz = tz * p.out_stride + p.out_off[0], oy = ty * p.out_stride + p.out_off[1],
                ox = tx * p.out_stride + p.out_off[2];
      const bool valid = (ty < p.Ht) && (tx < p.Wt) && (oz < p.Do) && (oy < p.Ho) && (ox < p.Wo);
      const size_t vox = (((size_t)b * p.Do + oz) * p.Ho + oy) * p.Wo + ox;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t rh[32];
        float v[32];
        tmem_ld32(taddr + c0, rh);
        if (PLANES == 2) {
          uint32_t rl[32];
          tmem_ld32(taddr + COUT + c0, rl);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rh[j]) + __uint_as_float(rl[j]);
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rh[j]);
        }
        if (c0 + 32 >= COUT) {           // all TMEM reads of this tile are done: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (valid) {
          const size_t off = vox * COUT + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] * s_scale[c0 + j] + s_shift[c0 + j];
          if (p.res_pre) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float f[8];
              load8_rt(p.res_pre, p.res_plane, p.planes_res, off + q * 8, f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
          if (p.res_post) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float f[8];
              load8_rt(p.res_post, p.res_plane, p.planes_res, off + q * 8, f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) store8_rt(p.y, p.y_plane, p.planes_out, off + q * 8, v + q * 8);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------ weight pack: [tap][plane][Cout][Cin] bf16
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, int transposed, int Co, int Ci, int taps,
                                      __nv_bfloat16* __restrict__ out, int planes) {
  const int total = taps * Co * Ci;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Ci, co = (i / Ci) % Co, t = i / (Ci * Co);
    const float v = transposed ? w[((size_t)ci * Co + co) * taps + t] : w[((size_t)co * Ci + ci) * taps + t];
    uint32_t lo;
    const uint32_t hi = split_bf16(v, lo);
    out[((size_t)(t * planes + 0) * Co + co) * Ci + ci] = __ushort_as_bfloat16((unsigned short)hi);
    if (planes == 2) out[((size_t)(t * planes + 1) * Co + co) * Ci + ci] = __ushort_as_bfloat16((unsigned short)lo);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 5-D view of cost planes: dims (C, W, H, D, planes*B) with arbitrary element strides per axis
static bool make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int D, int NB, size_t sW, size_t sH,
                         size_t sD, size_t sB) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NB};
  cuuint64_t strides[4] = {sW * 2, sH * 2, sD * 2, sB * 2};
  cuuint32_t box[5] = {(cuuint32_t)C, TC_TW, TC_TH, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = (C * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_w_map(CUtensorMap* m, const void* base, int Cin, int rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  cuuint32_t box[2] = {(cuuint32_t)Cin, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUtensorMapSwizzle sw = (Cin * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_num_sms = 0;

template <int CIN, int COUT, int PLANES>
static int launch_tc(const TcMaps& maps, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<CIN, COUT, PLANES>;
  static_assert(Cfg::STAGES >= 2, "pipeline too shallow");
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  cudaFuncSetAttribute(conv_tc_kernel<CIN, COUT, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       Cfg::SMEM_BYTES);
  const int total = p.B * p.Dt * p.tiles_h * p.tiles_w;
  const int grid = total < g_num_sms ? total : g_num_sms;
  conv_tc_kernel<CIN, COUT, PLANES><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(maps, p);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

}  // namespace dca

using namespace dca;

extern "C" long long dca_pack_weights_tc_bytes(int Co, int Ci, int taps, int planes) {
  if (Co <= 0 || Ci <= 0 || taps <= 0 || planes < 1 || planes > 2) return 0;
  if (!((Ci == 32 || Ci == 64) && (Co == 32 || Co == 64))) return 0;
  return (long long)taps * planes * Co * Ci * 2;
}

extern "C" int dca_pack_weights_tc(const float* w, int transposed, int Co, int Ci, int taps, void* out, int planes,
                                   void* stream) {
  if (!w || !out || dca_pack_weights_tc_bytes(Co, Ci, taps, planes) == 0) return DCA_ERR_ARG;
  const int total = taps * Co * Ci;
  pack_weight_tc_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, transposed, Co, Ci, taps,
                                                                               (__nv_bfloat16*)out, planes);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

// mode: DCA_CONV_K3S1 (0), DCA_CONV_K3S2 (1), DCA_CONV_T3S2 (2), DCA_CONV_K1 (3)
extern "C" int dca_conv3d_tc(int mode, const void* x, int planes_in, const void* w_tc, const float* scale,
                             const float* shift, const void* res_pre, const void* res_post, int planes_res, void* y,
                             int planes_out, int act, int B, int Cin, int Cout, int Di, int Hi, int Wi, int Do, int Ho,
                             int Wo, void* stream) {
  if (!x || !w_tc || !y || B <= 0 || planes_in < 1 || planes_in > 2 || planes_out < 1 || planes_out > 2)
    return DCA_ERR_ARG;
  if (!((Cin == 32 || Cin == 64) && (Cout == 32 || Cout == 64)) || mode < 0 || mode > 3) return DCA_ERR_UNSUPPORTED;
  if ((mode == 0 || mode == 3) && (Do != Di || Ho != Hi || Wo != Wi)) return DCA_ERR_ARG;
  if (mode == 1 && (Do != (Di + 1) / 2 || Ho != (Hi + 1) / 2 || Wo != (Wi + 1) / 2)) return DCA_ERR_ARG;
  if (mode == 2 && (Do != 2 * Di || Ho != 2 * Hi || Wo != 2 * Wi)) return DCA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = planes_in;
  TcMaps maps;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  p.scale = scale; p.shift = shift;
  p.res_pre = (const __nv_bfloat16*)res_pre; p.res_post = (const __nv_bfloat16*)res_post;
  p.res_plane = (size_t)B * Do * Ho * Wo * Cout; p.planes_res = planes_res;
  p.y = (__nv_bfloat16*)y; p.y_plane = p.res_plane; p.planes_out = planes_out; p.act = act;
  const size_t sW = Cin, sH = (size_t)Wi * Cin, sD = (size_t)Hi * Wi * Cin, sB = (size_t)Di * Hi * Wi * Cin;
  const int ntaps_total = (mode == 3) ? 1 : 27;
  if (!make_w_map(&maps.w, w_tc, Cin, ntaps_total * P * Cout, P * Cout)) return DCA_ERR_LAUNCH;

  auto run = [&]() -> int {
    p.tiles_w = (p.Wt + TC_TW - 1) / TC_TW;
    p.tiles_h = (p.Ht + TC_TH - 1) / TC_TH;
#define DCA_TC_CASE(CI, CO)                                                        \
  if (Cin == CI && Cout == CO)                                                     \
    return P == 2 ? launch_tc<CI, CO, 2>(maps, p, st) : launch_tc<CI, CO, 1>(maps, p, st);
    DCA_TC_CASE(32, 32)
    DCA_TC_CASE(64, 32)
    DCA_TC_CASE(32, 64)
    DCA_TC_CASE(64, 64)
#undef DCA_TC_CASE
    return DCA_ERR_UNSUPPORTED;
  };

  if (mode == 0 || mode == 3) {
    if (!make_act_map(&maps.a[0], x, Cin, Wi, Hi, Di, P * B, sW, sH, sD, sB)) return DCA_ERR_LAUNCH;
    for (int i = 1; i < 8; ++i) maps.a[i] = maps.a[0];
    p.Dt = Do; p.Ht = Ho; p.Wt = Wo; p.out_stride = 1;
    if (mode == 3) {
      p.ntaps = 1; p.tap_map[0] = 0; p.tap_w[0] = 0;
    } else {
      for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
        const int t = p.ntaps++;
        p.tap_off[t][0] = (signed char)(kd - 1); p.tap_off[t][1] = (signed char)(kh - 1);
        p.tap_off[t][2] = (signed char)(kw - 1);
        p.tap_map[t] = 0; p.tap_w[t] = (signed char)((kd * 3 + kh) * 3 + kw);
      }
    }
    return run();
  }
  if (mode == 1) {
    // stride 2: output o reads input 2o + k - 1.  Tap k reads the parity view q = (k != 1) of that axis at
    // view index o + (k == 0 ? -1 : 0).  8 parity views = 8 tensor maps with doubled strides.
    for (int pc = 0; pc < 8; ++pc) {
      const int pz = (pc >> 2) & 1, py = (pc >> 1) & 1, px = pc & 1;
      const int vd = (Di - pz + 1) / 2, vh = (Hi - py + 1) / 2, vw = (Wi - px + 1) / 2;
      const __nv_bfloat16* base = (const __nv_bfloat16*)x + pz * sD + py * sH + px * sW;
      if (vd <= 0 || vh <= 0 || vw <= 0) { maps.a[pc] = maps.a[0]; continue; }
      if (!make_act_map(&maps.a[pc], base, Cin, vw, vh, vd, P * B, 2 * sW, 2 * sH, 2 * sD, sB)) return DCA_ERR_LAUNCH;
    }
    p.Dt = Do; p.Ht = Ho; p.Wt = Wo; p.out_stride = 1;
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
      const int t = p.ntaps++;
      p.tap_off[t][0] = (signed char)(kd == 0 ? -1 : 0); p.tap_off[t][1] = (signed char)(kh == 0 ? -1 : 0);
      p.tap_off[t][2] = (signed char)(kw == 0 ? -1 : 0);
      p.tap_map[t] = (signed char)(((kd != 1) << 2) | ((kh != 1) << 1) | (kw != 1));
      p.tap_w[t] = (signed char)((kd * 3 + kh) * 3 + kw);
    }
    return run();
  }
  // mode 2: transposed conv, one launch per output parity class (o = 2i - 1 + k)
  if (!make_act_map(&maps.a[0], x, Cin, Wi, Hi, Di, P * B, sW, sH, sD, sB)) return DCA_ERR_LAUNCH;
  for (int i = 1; i < 8; ++i) maps.a[i] = maps.a[0];
  p.Dt = Di; p.Ht = Hi; p.Wt = Wi; p.out_stride = 2;
  for (int pc = 0; pc < 8; ++pc) {
    const int par[3] = {(pc >> 2) & 1, (pc >> 1) & 1, pc & 1};
    p.out_off[0] = par[0]; p.out_off[1] = par[1]; p.out_off[2] = par[2];
    p.ntaps = 0;
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
      const int k[3] = {kd, kh, kw};
      int off[3]; bool ok = true;
      for (int a = 0; a < 3; ++a) { const int n = par[a] + 1 - k[a]; if (n & 1) ok = false; off[a] = n / 2; }
      if (!ok) continue;
      const int t = p.ntaps++;
      p.tap_off[t][0] = (signed char)off[0]; p.tap_off[t][1] = (signed char)off[1]; p.tap_off[t][2] = (signed char)off[2];
      p.tap_map[t] = 0; p.tap_w[t] = (signed char)((kd * 3 + kh) * 3 + kw);
    }
    const int rc = run();
    if (rc != DCA_OK) return rc;
  }
  return DCA_OK;
}
