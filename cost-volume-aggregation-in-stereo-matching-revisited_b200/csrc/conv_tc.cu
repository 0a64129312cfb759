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
//        MMA#2  A=lo, B=[Whi]     (N = Cout)    -> cols [Cout,2Cout) += lo*Whi
//     and the epilogue adds the two column halves.  The small correction terms (hi*Wlo, lo*Whi) share one block so that
//     the main block sees one third of the accumulation steps: the tensor core TRUNCATES its fp32 accumulator toward
//     zero at every MMA (measured: a systematic -1.5e-8 relative per step, benchmarks/tc_bias_probe.py), which is the
//     dominant error of this path, and the truncation of the correction block is 2^-11 smaller.
//   planes == 1 ("fast"): one MMA, plain bf16 operands.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
//   warps 2-9 = epilogue (TMEM -> registers -> [+ fused trilinear x2 term] -> BN scale/shift, residuals,
//   activation -> bf16 planes).
// Persistent CTAs (grid = #SMs) walk output tiles round-robin.
#include <cuda.h>
#include <cstring>

#include "dca_common.cuh"

namespace dca {

constexpr int TC_TW = 8, TC_TH = 16, TC_M = 128;
constexpr int TC_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int MARCH_CORR = 256;       // march kernel, parity precision: TMEM column offset of the correction accumulators
constexpr int TC_MAX_TAPS = 36;       // 27 + one fused 1x1x1 side tap per transposed-conv parity class

struct TcMaps {
  CUtensorMap a[9];   // input views: [0..7] parity classes of stride-2 convs (else only [0]); for the transposed conv
                      // with a fused 1x1x1 side input, [1+c] = parity view c of that side input
  CUtensorMap w;      // packed weights [taps*planes*Cout rows][Cin]
};

struct TcParams {
  int B, Do, Ho, Wo;
  int Dt, Ht, Wt;                 // tile-space extent
  int tiles_w, tiles_h;
  int out_stride, out_off[3];
  int out_stride_d;                // depth stride of the output mapping (0 = same as out_stride)
  int ncls;                        // output classes sharing one launch (8 parity classes of a transposed conv, else 1)
  unsigned char cls_tap0[9];       // taps [cls_tap0[c], cls_tap0[c+1]) belong to class c
  signed char cls_off[8][3];       // output offset of class c
  int ntaps;
  signed char tap_off[TC_MAX_TAPS][3];
  signed char tap_map[TC_MAX_TAPS];
  signed char tap_w[TC_MAX_TAPS];
  const float* scale; const float* shift;
  const __nv_bfloat16* res_pre; const __nv_bfloat16* res_post; size_t res_plane; int planes_res;
  __nv_bfloat16* y; size_t y_plane; int planes_out; int act;
  const __nv_bfloat16* up; int planes_up;   // optional half-resolution tensor added (trilinear x2) before BN
  int nslab; int slab_c0[5]; int slab_dz[5];   // halo kernel: slabs per tile (3 depth slabs, or up to 5 64-channel slabs in 2-D)
  int slab_map[5];                 // ... and the input view (TcMaps::a index) each slab is read from (2-D channel concat)
  int par2d;                       // 2-D dilation-2 conv: the tile-space depth index td in [0, 4) is the (row, column) PARITY of a
                                   // sub-image: input view maps.a[td], output pixel (2 ty + (td >> 1), 2 tx + (td & 1)), depth 0
  int act_post;                    // activation applied once more AFTER res_post (ResidualBlock: relu(x + relu(bn(conv))))
  int ldc, co_base, cout_valid, out_f32;        // output row pitch / first channel / valid channels / fp32 output
  float inv_tiles_w, inv_tiles_h, inv_Dt, inv_ncls;   // reciprocals for the epilogue's tile decode (fast_divmod)
  int march_n;                     // > 0: depth-marching kernel, a work item = march_n consecutive output planes
  int cls_inner;                   // epilogue/work order: CTA owns whole tiles, classes inside (up2 kernel)
  int pair;                        // cls_inner + TWO depth-adjacent tiles per work item (deconv pair kernel): items run
                                   // (pair, class, sub-tile), 4 accumulator buffers, Dt counts depth PAIRS
  float cls_comp[8];               // per class: 1 + kappa * (MMA steps accumulated into the main block), see fill_comp()
  int linear;                      // 1x1x1 convs: tiles are 128 CONSECUTIVE voxels (8 KB bursts); rB/rD/rH/rW = real dims
  int rB, rD, rH, rW;
  int dbg;       // diagnostics: bit0 = skip epilogue stores, bit1 = skip MMAs (timing experiments only)
};

static inline void fill_recips(TcParams& p) {
  p.inv_tiles_w = 1.0f / (float)(p.tiles_w > 0 ? p.tiles_w : 1);
  p.inv_tiles_h = 1.0f / (float)(p.tiles_h > 0 ? p.tiles_h : 1);
  p.inv_Dt = 1.0f / (float)(p.Dt > 0 ? p.Dt : 1);
  const int nc = p.march_n > 0 ? p.march_n : (p.ncls > 0 ? p.ncls : 1);
  p.inv_ncls = 1.0f / (float)nc;
}

// The tensor core truncates its fp32 accumulator toward zero at every MMA: a conv output comes out smaller by a
// systematic kappa = 1.56e-8 (relative) per accumulation step of its main column block (measured on B200 with
// benchmarks/tc_bias_probe.py: -8.5e-7 after 54 steps, -1.66e-6 after 108; the correction block is 2^-11 smaller and
// needs nothing).  The epilogue multiplies the main block by 1 + kappa * steps.  dca_tc_set_trunc_comp(0) disables it.
static float g_trunc_kappa = 1.56e-8f;
static inline void fill_comp(TcParams& p, int cls, int steps) { p.cls_comp[cls] = 1.0f + g_trunc_kappa * (float)steps; }

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
// L2 prefetch of a tensor-map box (no shared memory, no barrier): decouples the DRAM latency of an operand from the
// availability of its smem slot
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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
// descriptor with the start-address field left zero: desc(addr) = desc_hi_const<ROWB>(sbo) + (addr >> 4)
template <int ROWB>
__host__ __device__ constexpr uint64_t desc_const(uint32_t sbo_bytes) {
  return (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((ROWB == 128 ? 2ull : 4ull) << 61);
}
// instruction descriptor, kind::f16: D=f32 (bit4), A=B=bf16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  // A/B format field: 0 = fp16, 1 = bf16 (DCA_F16_PLANES selects the plane format, dca_common.cuh)
  return (1u << 4) | ((uint32_t)(DCA_F16_PLANES ? 0 : 1) << 7) | ((uint32_t)(DCA_F16_PLANES ? 0 : 1) << 10) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// n / d and n % d for 0 <= n < 2^24 with a precomputed float reciprocal (exact after one correction step);
// replaces the ~25-instruction integer division sequences of the per-tile index decode
__device__ __forceinline__ int fast_divmod(int n, int d, float inv, int& rem) {
  int q = __float2int_rz(__int2float_rn(n) * inv);
  int r = n - q * d;
  if (r >= d) { ++q; r -= d; }
  if (r < 0) { --q; r += d; }
  rem = r;
  return q;
}

// ------------------------------------------------------------------ shared epilogue (warps 2..9)
// TMEM accumulator -> registers -> (+lo half) -> [+ trilinear x2 of `up`] -> BN scale/shift -> +res_pre -> act
// -> +res_post -> bf16 planes.  Eight warps: warp w reads TMEM lanes [32*(w%4), +32) (hardware restriction) and
// the 16-column half (w-2)/4 of every 32-channel chunk, so a thread finishes 16 channels of one voxel.
// Residual / upsample source rows are fetched BEFORE waiting for the accumulator (latency overlaps the MMAs).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z) : "memory");
}
// (a, b) -> packed bf16x2 hi and bf16x2 lo (= bf16 of the remainders)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) { split_pair(a, b, hi, lo); }
__device__ __forceinline__ void add_raw16(float* v, const uint4* raw, int planes) {   // raw[0..1] hi, raw[2..3] lo
  float f[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    unpack8(raw[q], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
    if (planes == 2) {
      unpack8(raw[2 + q], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
    }
  }
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread's 16 channels of one plane are exactly one 32-byte
// sector, so one instruction moves a whole sector instead of two half-sector requests
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void load_raw16(uint4* raw, const __nv_bfloat16* base, size_t plane, int planes, size_t off) {
  ldg256(base + off, raw[0], raw[1]);
  if (planes == 2) ldg256(base + plane + off, raw[2], raw[3]);
}

// BN scale / shift live in shared memory; the epilogues get them as generic pointers, and a generic LD of shared memory
// goes through the global-load path (long scoreboard: ncu showed 24 % of all stall samples of the deconv kernel on the
// first FFMA after these loads).  Explicit ld.shared keeps them on the 20-cycle LDS path.
__device__ __forceinline__ float4 lds_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

template <int COUT, int PLANES>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, int total_tiles, uint32_t tmem_base, uint64_t* tfull,
                                            uint64_t* tempty, const float* s_scale, const float* s_shift, int warp,
                                            int lane) {
  const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are reachable from this warp
  const int half = (warp - 2) >> 2;             // which 16 columns of each 32-channel chunk
  const int row = quarter * 32 + lane;          // tile row = voxel
  const int hh = row / TC_TW, ww = row % TC_TW;
  uint32_t it = 0;
  // flat mode: work item = (tile, class) pairs round-robin over CTAs; cls_inner mode: a CTA owns whole tiles and
  // walks the ncls classes inside each (total_tiles then counts tiles only)
  const int step = p.cls_inner ? 1 : (int)gridDim.x;
  const int first = p.cls_inner ? 0 : (int)blockIdx.x;
  const int my_tiles = p.cls_inner ? ((total_tiles > (int)blockIdx.x) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0) : 0;
  // march mode: a CTA owns whole work items (tile column x chunk of march_n output planes) and walks the planes inside;
  // the accumulator of plane j sits in TMEM columns 32*(n-1-j) and is announced by its own mbarrier tfull[j]
  const int mn = p.march_n;
  const int my_march = mn ? ((total_tiles > (int)blockIdx.x) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0) : 0;
  const int n_items = mn ? my_march * mn : (p.cls_inner ? my_tiles * p.ncls * (p.pair ? 2 : 1) : total_tiles);
  const uint32_t acc_mask = p.pair ? 3u : 1u, acc_shift = p.pair ? 2u : 1u;
  if (mn) {
    // march mode: clear this thread's slice of every accumulator plane once; afterwards a plane is cleared again right after
    // it has been drained (below) and handed back through ITS OWN mbarrier tempty[j], so the MMA warp starts the next
    // work item while the last planes of this one are still in the epilogue (no per-item drain bubble)
    const uint32_t zaddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 16;
    for (int jj = 0; jj < mn; ++jj) {
      tmem_st16_zero(zaddr + 32 * jj);
      if (PLANES == 2) tmem_st16_zero(zaddr + MARCH_CORR + 32 * jj);     // correction block (hi.Wlo + lo.Whi)
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int jj = 0; jj < mn; ++jj) mbar_arrive(&tempty[jj]);
  }
  for (int item = (mn ? 0 : first); item < n_items; item += (mn ? 1 : step), ++it) {
    int r, cls, mj = 0, mwi = 0, sub = 0;
    if (mn) { cls = 0; mwi = fast_divmod(item, mn, p.inv_ncls, mj); r = (int)blockIdx.x + mwi * (int)gridDim.x; }
    else if (p.pair) { sub = item & 1; cls = (item >> 1) & 7; r = (int)blockIdx.x + (item >> 4) * (int)gridDim.x; }   // ncls == 8
    else if (p.cls_inner) { const int q = fast_divmod(item, p.ncls, p.inv_ncls, cls); r = (int)blockIdx.x + q * (int)gridDim.x; }
    else if (p.ncls > 1) { r = fast_divmod(item, p.ncls, p.inv_ncls, cls); }
    else { cls = 0; r = item; }
    int tw, th, td;
    r = fast_divmod(r, p.tiles_w, p.inv_tiles_w, tw);
    r = fast_divmod(r, p.tiles_h, p.inv_tiles_h, th);
    const int b = fast_divmod(r, p.Dt, p.inv_Dt, td);
    if (p.pair) td = 2 * td + sub;
    const uint32_t acc = it & acc_mask;
    const int ty = th * TC_TH + hh, tx = tw * TC_TW + ww;
    int oz = mn ? td * mn + mj : td * (p.out_stride_d ? p.out_stride_d : p.out_stride) + p.cls_off[cls][0],
        oy = ty * p.out_stride + p.cls_off[cls][1],
        ox = tx * p.out_stride + p.cls_off[cls][2];
    if (p.par2d) { oz = 0; oy += td >> 1; ox += td & 1; }
    const bool valid = (ty < p.Ht) && (tx < p.Wt) && (oz < p.Do) && (oy < p.Ho) && (ox < p.Wo);
    const size_t vox = (((size_t)b * p.Do + oz) * p.Ho + oy) * p.Wo + ox;
    int bb = b, uD = p.Do, uH = p.Ho, uW = p.Wo;
    if (p.linear && p.up != nullptr) {   // only the fused-upsample term needs the real voxel coordinates of a linear tile
      // (32-bit arithmetic: the 64-bit div/mod chain that used to sit here cost ~2000 cycles per tile and bounded every
      //  linear 1x1x1 launch; linear tiles index fewer than 2^31 voxels by construction)
      unsigned r2 = (unsigned)vox;
      const unsigned q1 = r2 / (unsigned)p.rW; ox = (int)(r2 - q1 * (unsigned)p.rW);
      const unsigned q2 = q1 / (unsigned)p.rH; oy = (int)(q1 - q2 * (unsigned)p.rH);
      const unsigned q3 = q2 / (unsigned)p.rD; oz = (int)(q2 - q3 * (unsigned)p.rD);
      bb = (int)q3;
      uD = p.rD; uH = p.rH; uW = p.rW;
    }
    const size_t off0 = vox * (size_t)p.ldc + p.co_base + half * 16;   // first chunk's 16 channels of this thread
    uint4 pre_raw[4], post_raw[4];
    const bool has_pre = valid && p.res_pre != nullptr, has_post = valid && p.res_post != nullptr;
    if (has_pre) load_raw16(pre_raw, p.res_pre, p.res_plane, p.planes_res, off0);
    if (has_post) load_raw16(post_raw, p.res_post, p.res_plane, p.planes_res, off0);
    // fused trilinear x2 (align_corners=False) of a half-resolution tensor `up` with COUT channels
    float upv[16];
    const bool has_up = valid && p.up != nullptr;
    if (has_up) {
#pragma unroll
      for (int j = 0; j < 16; ++j) upv[j] = 0.f;
      const int Dl = uD >> 1, Hl = uH >> 1, Wl = uW >> 1;
      // even o = 2i: .25*in[i-1] + .75*in[i];  odd o = 2i+1: .75*in[i] + .25*in[i+1]   (indices clamped)
      const int z0 = (oz >> 1) - ((oz & 1) ? 0 : 1), y0 = (oy >> 1) - ((oy & 1) ? 0 : 1), x0 = (ox >> 1) - ((ox & 1) ? 0 : 1);
      const float wz = (oz & 1) ? 0.75f : 0.25f, wy = (oy & 1) ? 0.75f : 0.25f, wx = (ox & 1) ? 0.75f : 0.25f;
      const size_t uplane = (size_t)(p.linear ? p.rB : p.B) * Dl * Hl * Wl * COUT;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int iz = min(max(z0 + (k >> 2), 0), Dl - 1), iy = min(max(y0 + ((k >> 1) & 1), 0), Hl - 1),
                  ix = min(max(x0 + (k & 1), 0), Wl - 1);
        const float wgt = ((k >> 2) ? 1.f - wz : wz) * (((k >> 1) & 1) ? 1.f - wy : wy) * ((k & 1) ? 1.f - wx : wx);
        uint4 raw[4];
        load_raw16(raw, p.up, uplane, p.planes_up, ((((size_t)bb * Dl + iz) * Hl + iy) * Wl + ix) * COUT + half * 16);
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = 0.f;
        add_raw16(f, raw, p.planes_up);
#pragma unroll
        for (int j = 0; j < 16; ++j) upv[j] = fmaf(wgt, f[j], upv[j]);
      }
    }
    if (mn) mbar_wait(&tfull[mj], (uint32_t)(mwi & 1));
    else mbar_wait(&tfull[acc], (it >> acc_shift) & 1);
    tc_fence_after();
    const float comp = p.cls_comp[cls];      // undo the accumulator's truncation shrink of the main block
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 16 +
                           (mn ? (uint32_t)(32 * (mn - 1 - mj)) : acc * (uint32_t)(PLANES * COUT));
#pragma unroll 1
    for (int c0 = 0; c0 < COUT; c0 += 32) {
      uint32_t rh[16];
      float v[16];
      tmem_ld16(taddr + c0, rh);
      if (PLANES == 2) {                 // main block + correction block (march mode: MARCH_CORR columns further)
        uint32_t rl[16];
        tmem_ld16(taddr + (mn ? (uint32_t)MARCH_CORR : (uint32_t)COUT) + c0, rl);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(rh[j]), comp, __uint_as_float(rl[j]));
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rh[j]) * comp;
      }
      if (c0 + 32 >= COUT && !mn) {    // all TMEM reads of this tile are done: hand the buffer back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
      if (mn) {                        // plane drained: clear this thread's slice of it and hand it back (tempty[mj])
        tmem_st16_zero(taddr);
        if (PLANES == 2) tmem_st16_zero(taddr + MARCH_CORR);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[mj]);
      }
      if (valid && !(p.dbg & 1)) {
        const size_t off = off0 + c0;
        const int cb = c0 + half * 16;
        if (has_up) {
          if (c0 == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += upv[j];
          }   // (the fused upsample is only used with COUT == 32)
        }
        {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 a = lds_f4(s_scale + cb + 4 * j4), c = lds_f4(s_shift + cb + 4 * j4);   // (16-float aligned)
            v[4 * j4 + 0] = fmaf(v[4 * j4 + 0], a.x, c.x); v[4 * j4 + 1] = fmaf(v[4 * j4 + 1], a.y, c.y);
            v[4 * j4 + 2] = fmaf(v[4 * j4 + 2], a.z, c.z); v[4 * j4 + 3] = fmaf(v[4 * j4 + 3], a.w, c.w);
          }
        }
        if (p.res_pre) {
          if (c0 != 0) load_raw16(pre_raw, p.res_pre, p.res_plane, p.planes_res, off);
          add_raw16(v, pre_raw, p.planes_res);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
        if (p.res_post) {
          if (c0 != 0) load_raw16(post_raw, p.res_post, p.res_plane, p.planes_res, off);
          add_raw16(v, post_raw, p.planes_res);
          if (p.act_post) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act_post);
          }
        }
        if (p.out_f32 == 2) {                 // fp32 CHANNEL-major output [Cout][nvox] (per-tap partial sums of the 32->1 convs)
          const size_t nvox = (size_t)p.B * p.Do * p.Ho * p.Wo;
          float* yf = reinterpret_cast<float*>(p.y) + (size_t)(p.co_base + cb) * nvox + vox;   // one 64-bit product per tile
          const int nval = p.cout_valid - (p.co_base + cb);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nval) *yf = v[j];
            yf += nvox;
          }
        } else if (p.out_f32) {               // fp32 channels-last output (mask logits of the propagation net)
          float* yf = reinterpret_cast<float*>(p.y) + off;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (p.co_base + cb + j < p.cout_valid) yf[j] = v[j];
        } else {
          uint32_t hw[8], lw[8];
          if (PLANES == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) split2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              hw[j] = f2_to_h16x2(v[2 * j], v[2 * j + 1]);
            }
          }
          stg256(p.y + off, hw);
          if (PLANES == 2 && p.planes_out == 2) stg256(p.y + p.y_plane + off, lw);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ fused tail: conv(32ch, BN, ReLU) -> Conv3d(32 -> 1, k3)
// classif3 (gwcnet_dca_g.py:166-168) and cva.classify (cva.py:51-53) are a 3x3x3 conv to 32 channels followed by a 3x3x3
// conv to ONE channel.  The second conv is linear in its 27 taps: logit[v] = sum_tap P[tap][v + off(tap)] with
// P[tap][v] = sum_c w27[tap][c] * h[v][c].  The first conv's epilogue holds h[v][0..32) of one voxel in the registers of
// one thread (fp32, BEFORE any rounding to 16-bit planes), so it computes the 27 per-tap products right there on the
// CUDA cores (864 FFMA per voxel, weights as constant-bank operands: they are launch parameters) and stores P (fp32,
// tap-major [27][B*D*H*W]) INSTEAD of h: h is never written to or re-read from HBM, and the separate per-tap GEMM launch
// (conv_tc_kernel<32,32>, 4.5 % tensor pipe) disappears.  dca_tap_gather_* (dca_ops.cu) then sums the shifted taps.
// Thread = voxel with all 32 channels; the two sets of four epilogue warps (TMEM lane quarters 0..3 each) alternate over
// output planes (march kernel) / tiles (halo kernel).
struct TapW { float w[27 * 32]; };      // [tap = (kd*3+kh)*3+kw][ci]
struct NoTapW { int unused; };
template <int TAPS> struct TapArg { typedef NoTapW type; };
template <> struct TapArg<1> { typedef TapW type; };

template <int PLANES>
__device__ __forceinline__ void taps_load_act(uint32_t t_main, uint32_t t_corr, float comp, const float* s_scale,
                                              const float* s_shift, int act, float* v) {
  uint32_t rh[32];
  tmem_ld32(t_main, rh);
  if (PLANES == 2) {
    uint32_t rl[32];
    tmem_ld32(t_corr, rl);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(rh[j]), comp, __uint_as_float(rl[j]));
  } else {
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rh[j]) * comp;
  }
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 a = lds_f4(s_scale + 4 * j4), c = lds_f4(s_shift + 4 * j4);
    v[4 * j4 + 0] = apply_act(fmaf(v[4 * j4 + 0], a.x, c.x), act); v[4 * j4 + 1] = apply_act(fmaf(v[4 * j4 + 1], a.y, c.y), act);
    v[4 * j4 + 2] = apply_act(fmaf(v[4 * j4 + 2], a.z, c.z), act); v[4 * j4 + 3] = apply_act(fmaf(v[4 * j4 + 3], a.w, c.w), act);
  }
}

__device__ __forceinline__ void taps_store(const float* v, const TapW& tw, float* __restrict__ P, size_t nvox, size_t vox) {
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      a0 = fmaf(v[c], tw.w[t * 32 + c], a0);
      a1 = fmaf(v[c + 1], tw.w[t * 32 + c + 1], a1);
    }
    P[(size_t)t * nvox + vox] = a0 + a1;
  }
}

template <int PLANES>
__device__ __forceinline__ void tc_epilogue_taps(const TcParams& p, const TapW& tw, int total_tiles, uint32_t tmem_base,
                                                 uint64_t* tfull, uint64_t* tempty, const float* s_scale,
                                                 const float* s_shift, int warp, int lane) {
  constexpr int COUT = 32;
  const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are reachable from this warp
  const int set = (warp - 2) >> 2;              // which half of the planes / tiles this warp drains
  const int row = quarter * 32 + lane;
  const int hh = row / TC_TW, ww = row % TC_TW;
  const int mn = p.march_n;
  const size_t nvox = (size_t)p.B * p.Do * p.Ho * p.Wo;
  float* __restrict__ Pout = reinterpret_cast<float*>(p.y);
  const float comp = p.cls_comp[0];
  const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
  const int my_items = (total_tiles > (int)blockIdx.x) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (mn) {
    // clear the accumulator planes this thread drains (all 32 columns of its lanes) once; afterwards each plane is cleared
    // right after it has been drained and handed back through its own mbarrier (rolling: no per-item drain bubble)
    for (int jj = set; jj < mn; jj += 2) {
      const uint32_t z = lane_base + (uint32_t)(32 * (mn - 1 - jj));
      tmem_st16_zero(z); tmem_st16_zero(z + 16);
      if (PLANES == 2) { tmem_st16_zero(z + MARCH_CORR); tmem_st16_zero(z + MARCH_CORR + 16); }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int jj = set; jj < mn; jj += 2) mbar_arrive(&tempty[jj]);
    for (int wi = 0; wi < my_items; ++wi) {
      int r = (int)blockIdx.x + wi * (int)gridDim.x, tw_, th, td;
      r = fast_divmod(r, p.tiles_w, p.inv_tiles_w, tw_);
      r = fast_divmod(r, p.tiles_h, p.inv_tiles_h, th);
      const int b = fast_divmod(r, p.Dt, p.inv_Dt, td);
      const int ty = th * TC_TH + hh, tx = tw_ * TC_TW + ww;
      for (int mj = set; mj < mn; mj += 2) {
        const int oz = td * mn + mj;
        const bool valid = (ty < p.Ho) && (tx < p.Wo) && (oz < p.Do);
        const size_t vox = (((size_t)b * p.Do + oz) * p.Ho + ty) * p.Wo + tx;
        mbar_wait(&tfull[mj], (uint32_t)(wi & 1));
        tc_fence_after();
        const uint32_t ta = lane_base + (uint32_t)(32 * (mn - 1 - mj));
        float v[32];
        taps_load_act<PLANES>(ta, ta + MARCH_CORR, comp, s_scale, s_shift, p.act, v);
        // plane drained: clear it and hand it back to the MMA warp (its own mbarrier, 4 arrivals: this set of warps)
        tmem_st16_zero(ta); tmem_st16_zero(ta + 16);
        if (PLANES == 2) { tmem_st16_zero(ta + MARCH_CORR); tmem_st16_zero(ta + MARCH_CORR + 16); }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[mj]);
        if (valid) taps_store(v, tw, Pout, nvox, vox);
      }
    }
  } else {
    for (int it = set; it < my_items; it += 2) {
      int r = (int)blockIdx.x + it * (int)gridDim.x, tw_, th, td;
      r = fast_divmod(r, p.tiles_w, p.inv_tiles_w, tw_);
      r = fast_divmod(r, p.tiles_h, p.inv_tiles_h, th);
      const int b = fast_divmod(r, p.Dt, p.inv_Dt, td);
      const int ty = th * TC_TH + hh, tx = tw_ * TC_TW + ww;
      const bool valid = (ty < p.Ho) && (tx < p.Wo) && (td < p.Do);
      const size_t vox = (((size_t)b * p.Do + td) * p.Ho + ty) * p.Wo + tx;
      const uint32_t acc = (uint32_t)(it & 1);
      mbar_wait(&tfull[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t ta = lane_base + acc * (uint32_t)(PLANES * COUT);
      float v[32];
      taps_load_act<PLANES>(ta, ta + COUT, comp, s_scale, s_shift, p.act, v);
      tc_fence_before();                         // all TMEM reads of this tile are done: hand the buffer back
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);  // (barrier count 4: one set of warps drains a tile)
      if (valid) taps_store(v, tw, Pout, nvox, vox);
    }
  }
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
  const int total_tiles = p.B * p.Dt * p.tiles_h * p.tiles_w * p.ncls;

  if (threadIdx.x < COUT) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
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
  pdl_trigger();                 // persistent single-wave grid: let the next kernel's CTAs take over SMs as ours retire
  if (warp != 0) pdl_wait();     // (the producer warp waits after its loads of static weights, see below)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      // single-tap launches (1x1x1 convs, per-tap GEMMs): every stage slot always holds the same weight tile, so it is
      // loaded only on the first pass over the ring (148 SMs re-reading one 4 KB tile per 128 voxels is an L2 hot spot)
      const bool one_tap = (p.ntaps == 1 && p.ncls == 1);
      pdl_wait();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int cls = r % p.ncls; r /= p.ncls;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int td = r % p.Dt;
        const int b = r / p.Dt;
        for (int t = p.cls_tap0[cls]; t < p.cls_tap0[cls + 1]; ++t) {
          mbar_wait(&empty[s], ph ^ 1);
          const bool need_w = !(one_tap && ph != 0);
          mbar_expect_tx(&full[s], need_w ? Cfg::STAGE_BYTES : PLANES * Cfg::A_BYTES);
          uint8_t* st = stage_base + (size_t)s * Cfg::STAGE_BYTES;
          const CUtensorMap* am = &maps.a[p.tap_map[t]];
          const int cw = tw * TC_TW + p.tap_off[t][2], chh = th * TC_TH + p.tap_off[t][1], cd = td + p.tap_off[t][0];
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_5d(st + pl * Cfg::A_BYTES, am, &full[s], 0, cw, chh, cd, pl * p.B + b);
          if (need_w) tma_load_2d(st + PLANES * Cfg::A_BYTES, &maps.w, &full[s], 0, p.tap_w[t] * Cfg::B_ROWS);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_full = make_idesc(TC_M, Cfg::NACC);
      constexpr uint32_t idesc_half = make_idesc(TC_M, COUT);
      constexpr uint64_t DAB = desc_const<Cfg::ROWB>(8 * Cfg::ROWB);
      const uint32_t stage_u32 = smem_u32(stage_base);
      uint32_t s = 0, ph = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * Cfg::NACC;
        const int cls = tile % p.ncls;
        const int t_first = p.cls_tap0[cls], t_last = p.cls_tap0[cls + 1];
        for (int t = t_first; t < t_last; ++t) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = stage_u32 + s * Cfg::STAGE_BYTES;
          const uint64_t da0 = DAB + (a0 >> 4);
          const uint64_t db0 = DAB + ((a0 + PLANES * Cfg::A_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k) {
            if (leader) {
              umma_bf16(d_addr, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc_full, (t > t_first || k > 0) ? 1u : 0u);
              if (PLANES == 2)
                umma_bf16(d_addr + COUT, da0 + (uint64_t)((Cfg::A_BYTES >> 4) + k * 2), db0 + (uint64_t)(k * 2), idesc_half, 1u);
            }
          }
          __syncwarp();
          if (leader) umma_commit(&empty[s]);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        if (leader) umma_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    tc_epilogue<COUT, PLANES>(p, total_tiles, tmem_base, tfull, tempty, s_scale, s_shift, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================================
// v2 main loop for 3x3x3 stride-1 convs: halo'd slab reuse.
// Instead of one shifted 128-row box per tap (27 TMA boxes / tile), ONE halo'd box [Cin, 10 w, 18 h] is
// loaded per depth slab kd (3 boxes / tile / plane) and the nine (kh,kw) taps of that slab read it through
// shifted UMMA descriptors: start = slab + (kh*10 + kw)*ROWB, 8-row groups (one h row of 8 voxels) are
// 10*ROWB apart (SBO).  TMA and UMMA both derive the 64B/128B swizzle XOR from absolute smem address
// bits, so any row-aligned start inside the slab addresses the right 16-byte chunks.
// Weights: resident in smem for the whole persistent CTA when 27 taps fit (WRES), else streamed in
// chunks of 3 taps (one (kd,kh) row) through their own mbarrier ring.
// =====================================================================================================
constexpr int HB_W = TC_TW + 2, HB_H = TC_TH + 2;     // halo box 10 x 18
static int g_num_sms = 0;

template <int CIN, int COUT, int PLANES>
struct HaloCfg {
  static constexpr int ROWB = CIN * 2;
  static constexpr int SLAB_BYTES = HB_W * HB_H * ROWB;                        // one plane of one slab
  static constexpr int SLAB_PITCH = (SLAB_BYTES + 1023) / 1024 * 1024;
  static constexpr int A_SLOT = PLANES * SLAB_PITCH;
  static constexpr int B_ROWS = PLANES * COUT;
  static constexpr int B_BYTES = B_ROWS * ROWB;                                // one tap
  static constexpr int BUDGET = 200 * 1024;   // (a 4th slab slot was measured: same TMA-only time, the ring is not latency bound)
  static constexpr bool WRES = (27 * B_BYTES + 2 * A_SLOT) <= BUDGET;
  static constexpr int W_CHUNK = 3 * B_BYTES;
  static constexpr int A_SLOTS_RES = ((BUDGET - 27 * B_BYTES) / A_SLOT) > 4 ? 4 : ((BUDGET - 27 * B_BYTES) / A_SLOT);
  static constexpr int A_SLOTS = WRES ? A_SLOTS_RES : 2;
  static constexpr int W_SLOTS_STR = ((BUDGET - 2 * A_SLOT) / W_CHUNK) > 4 ? 4 : ((BUDGET - 2 * A_SLOT) / W_CHUNK);
  static constexpr int W_SLOTS = WRES ? 1 : W_SLOTS_STR;
  static constexpr int W_BYTES_TOTAL = WRES ? 27 * B_BYTES : W_SLOTS * W_CHUNK;
  static constexpr int TMEM_COLS = 2 * PLANES * COUT < 32 ? 32 : 2 * PLANES * COUT;
  static constexpr int SMEM_BYTES = A_SLOTS * A_SLOT + W_BYTES_TOTAL + 1024 + 256 + 2 * COUT * 4;
};

template <int ROWB>
__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  constexpr uint64_t layout = (ROWB == 128) ? 2ull : 4ull;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}

template <int CIN, int COUT, int PLANES, int TAPS = 0>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_halo_kernel(const __grid_constant__ TcMaps maps, const TcParams p,
                    const __grid_constant__ typename TapArg<TAPS>::type tw) {
  using Cfg = HaloCfg<CIN, COUT, PLANES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;
  uint8_t* w_base = smem + Cfg::A_SLOTS * Cfg::A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + Cfg::W_BYTES_TOTAL);
  uint64_t* afull = bars;                       // [A_SLOTS]
  uint64_t* aempty = afull + Cfg::A_SLOTS;      // [A_SLOTS]
  uint64_t* wfull = aempty + Cfg::A_SLOTS;      // [W_SLOTS]
  uint64_t* wempty = wfull + Cfg::W_SLOTS;      // [W_SLOTS]
  uint64_t* tfull = wempty + Cfg::W_SLOTS;      // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(w_base + Cfg::W_BYTES_TOTAL + 256);
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.Dt * p.tiles_h * p.tiles_w;   // halo kernel: ncls == 1

  if (threadIdx.x < COUT) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TAPS ? 4 : 8); }
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
  pdl_trigger();                 // persistent single-wave grid: let the next kernel's CTAs take over SMs as ours retire
  if (warp != 0) pdl_wait();     // (the producer warp waits after its loads of static weights, see below)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (Cfg::WRES) {
        mbar_expect_tx(&wfull[0], p.nslab * 9 * Cfg::B_BYTES);
        for (int t = 0; t < p.nslab * 9; ++t) tma_load_2d(w_base + t * Cfg::B_BYTES, &maps.w, &wfull[0], 0, t * Cfg::B_ROWS);
      }
      pdl_wait();                // weights are static; everything below reads the previous kernel's output
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int td = r % p.Dt;
        const int b = r / p.Dt;
        for (int kd = 0; kd < p.nslab; ++kd) {
          mbar_wait(&aempty[sa], pa ^ 1);
          mbar_expect_tx(&afull[sa], PLANES * Cfg::SLAB_BYTES);
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_5d(a_base + sa * Cfg::A_SLOT + pl * Cfg::SLAB_PITCH, &maps.a[p.slab_map[kd] + (p.par2d ? td : 0)],
                        &afull[sa], p.slab_c0[kd], tw * TC_TW - 1, th * TC_TH - 1, (p.par2d ? 0 : td) + p.slab_dz[kd],
                        pl * p.B + b);
          if (++sa == Cfg::A_SLOTS) { sa = 0; pa ^= 1; }
          if (!Cfg::WRES) {
            for (int kh = 0; kh < 3; ++kh) {
              mbar_wait(&wempty[sw], pw ^ 1);
              mbar_expect_tx(&wfull[sw], Cfg::W_CHUNK);
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
                tma_load_2d(w_base + sw * Cfg::W_CHUNK + kw * Cfg::B_BYTES, &maps.w, &wfull[sw], 0,
                            ((kd * 3 + kh) * 3 + kw) * Cfg::B_ROWS);
              if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_full = make_idesc(TC_M, PLANES * COUT);
      constexpr uint32_t idesc_half = make_idesc(TC_M, COUT);
      constexpr uint64_t DA = desc_const<Cfg::ROWB>(HB_W * Cfg::ROWB);     // activation slab: 8-row groups 10 rows apart
      constexpr uint64_t DB = desc_const<Cfg::ROWB>(8 * Cfg::ROWB);        // weights: dense
      constexpr uint32_t NACC = PLANES * COUT;
      if (Cfg::WRES) { mbar_wait(&wfull[0], 0); tc_fence_after(); }
      const uint32_t a_u32 = smem_u32(a_base), w_u32 = smem_u32(w_base);
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * NACC;
#pragma unroll 1
        for (int kd = 0; kd < p.nslab; ++kd) {
          mbar_wait(&afull[sa], pa);
          tc_fence_after();
          const uint64_t da_slab = DA + ((a_u32 + sa * Cfg::A_SLOT) >> 4);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            uint64_t db_row;
            if (Cfg::WRES) {
              db_row = DB + ((w_u32 + (uint32_t)((kd * 3 + kh) * 3) * Cfg::B_BYTES) >> 4);
            } else {
              mbar_wait(&wfull[sw], pw);
              tc_fence_after();
              db_row = DB + ((w_u32 + sw * Cfg::W_CHUNK) >> 4);
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int k = 0; k < CIN / 16; ++k) {
                const uint64_t da = da_slab + (uint64_t)(((kh * HB_W + kw) * Cfg::ROWB + k * 32) >> 4);
                const uint64_t db = db_row + (uint64_t)((kw * Cfg::B_BYTES + k * 32) >> 4);
                const uint32_t accum = (kh == 0 && kw == 0 && k == 0) ? (kd != 0 ? 1u : 0u) : 1u;
                if (leader && !(p.dbg & 2)) {
                  umma_bf16(d_addr, da, db, idesc_full, accum);
                  if (PLANES == 2) umma_bf16(d_addr + COUT, da + (uint64_t)(Cfg::SLAB_PITCH >> 4), db, idesc_half, 1u);
                }
              }
            }
            if (!Cfg::WRES) {
              __syncwarp();
              if (leader) umma_commit(&wempty[sw]);
              if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
            }
          }
          __syncwarp();
          if (leader) umma_commit(&aempty[sa]);
          if (++sa == Cfg::A_SLOTS) { sa = 0; pa ^= 1; }
        }
        if (leader) umma_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    if constexpr (TAPS) tc_epilogue_taps<PLANES>(p, tw, total_tiles, tmem_base, tfull, tempty, s_scale, s_shift, warp, lane);
    else tc_epilogue<COUT, PLANES>(p, total_tiles, tmem_base, tfull, tempty, s_scale, s_shift, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================================
// Stride-2 3x3x3 conv, Cin = 32 -> Cout = 64 (Multi_Aggregation.conv1, cva.py:16-17) on halo slabs.
// The per-tap kernel above needs 27 element-strided TMA boxes per tile and plane and is L2->SMEM bound (0.96 GB per
// launch at KITTI).  Here the input is viewed as PAIRS of w-adjacent voxels, [.., W/2, 64 = 2 x 32 channels], i.e. rows
// of 128 bytes (SWIZZLE_128B).  One box [64, 9 pairs, 33 rows] per input plane (3 per tile) then serves all nine
// (kh, kw) taps through shifted UMMA descriptors:
//     tap (kh, kw) of output (ho, wo) reads input row 2ho + kh - 1 and column 2wo + kw - 1
//       = slab row 2*ho_l + kh (8-row groups 2 x 9 pair-rows apart: SBO = 2304 B),
//         pair wo_l + (kw == 0 ? 0 : 1) and the EVEN voxel of the pair for kw == 1, the ODD one otherwise:
//         the odd voxel is simply K-offset +64 bytes inside the 128-byte row.
// Weights (27 x 8 KB in parity precision) stream per tap through an 8-deep ring.
// =====================================================================================================
template <int PLANES>
struct S2Cfg {
  static constexpr int CIN = 32, COUT = 64;
  static constexpr int PW = TC_TW + 1, PH = 2 * TC_TH + 1;                  // slab: 9 pairs x 33 rows of 128 bytes
  static constexpr int SLAB_BYTES = PW * PH * 128;
  static constexpr int SLAB_PITCH = (SLAB_BYTES + 1023) / 1024 * 1024;
  static constexpr int A_SLOT = PLANES * SLAB_PITCH;
  static constexpr int A_SLOTS = PLANES == 2 ? 2 : 4;
  static constexpr int B_ROWS = PLANES * COUT;
  static constexpr int B_BYTES = B_ROWS * CIN * 2;                          // one tap
  static constexpr int W_SLOTS = 8;
  static constexpr int TMEM_COLS = 2 * PLANES * COUT;
  static constexpr int SMEM_BYTES = A_SLOTS * A_SLOT + W_SLOTS * B_BYTES + 1024 + 256 + 2 * COUT * 4;
};

template <int PLANES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_s2slab_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  using Cfg = S2Cfg<PLANES>;
  constexpr int COUT = Cfg::COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;
  uint8_t* w_base = smem + Cfg::A_SLOTS * Cfg::A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + Cfg::W_SLOTS * Cfg::B_BYTES);
  uint64_t* afull = bars;                       // [A_SLOTS]
  uint64_t* aempty = afull + Cfg::A_SLOTS;      // [A_SLOTS]
  uint64_t* wfull = aempty + Cfg::A_SLOTS;      // [W_SLOTS]
  uint64_t* wempty = wfull + Cfg::W_SLOTS;      // [W_SLOTS]
  uint64_t* tfull = wempty + Cfg::W_SLOTS;      // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(w_base + Cfg::W_SLOTS * Cfg::B_BYTES + 256);
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.Dt * p.tiles_h * p.tiles_w;

  if (threadIdx.x < COUT) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
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
  pdl_trigger();                 // persistent single-wave grid: let the next kernel's CTAs take over SMs as ours retire
  if (warp != 0) pdl_wait();     // (the producer warp waits after its loads of static weights, see below)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int td = r % p.Dt;
        const int b = r / p.Dt;
        for (int kd = 0; kd < 3; ++kd) {
          mbar_wait(&aempty[sa], pa ^ 1);
          mbar_expect_tx(&afull[sa], PLANES * Cfg::SLAB_BYTES);
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_5d(a_base + sa * Cfg::A_SLOT + pl * Cfg::SLAB_PITCH, &maps.a[0], &afull[sa], 0, tw * TC_TW - 1,
                        2 * th * TC_TH - 1, 2 * td + kd - 1, pl * p.B + b);
          if (++sa == Cfg::A_SLOTS) { sa = 0; pa ^= 1; }
          for (int t9 = 0; t9 < 9; ++t9) {
            mbar_wait(&wempty[sw], pw ^ 1);
            mbar_expect_tx(&wfull[sw], Cfg::B_BYTES);
            tma_load_2d(w_base + sw * Cfg::B_BYTES, &maps.w, &wfull[sw], 0, (kd * 9 + t9) * Cfg::B_ROWS);
            if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_full = make_idesc(TC_M, PLANES * COUT);
      constexpr uint32_t idesc_half = make_idesc(TC_M, COUT);
      constexpr uint64_t DA = desc_const<128>(2 * Cfg::PW * 128);      // pair rows, 8-row groups two slab rows apart
      constexpr uint64_t DB = desc_const<64>(8 * 64);                  // weights: dense 64-byte rows
      constexpr uint32_t NACC = PLANES * COUT;
      const uint32_t a_u32 = smem_u32(a_base), w_u32 = smem_u32(w_base);
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * NACC;
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
          mbar_wait(&afull[sa], pa);
          tc_fence_after();
          const uint64_t da_slab = DA + ((a_u32 + sa * Cfg::A_SLOT) >> 4);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              mbar_wait(&wfull[sw], pw);
              tc_fence_after();
              const uint64_t db_tap = DB + ((w_u32 + sw * Cfg::B_BYTES) >> 4);
              const int a_off = (kh * Cfg::PW + (kw == 0 ? 0 : 1)) * 128 + (kw == 1 ? 0 : 64);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint64_t da = da_slab + (uint64_t)((a_off + k * 32) >> 4);
                const uint64_t db = db_tap + (uint64_t)((k * 32) >> 4);
                const uint32_t accum = (kd == 0 && kh == 0 && kw == 0 && k == 0) ? 0u : 1u;
                if (leader && !(p.dbg & 2)) {
                  umma_bf16(d_addr, da, db, idesc_full, accum);
                  if (PLANES == 2) umma_bf16(d_addr + COUT, da + (uint64_t)(Cfg::SLAB_PITCH >> 4), db, idesc_half, 1u);
                }
              }
              __syncwarp();
              if (leader) umma_commit(&wempty[sw]);
              if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
            }
          }
          __syncwarp();
          if (leader) umma_commit(&aempty[sa]);
          if (++sa == Cfg::A_SLOTS) { sa = 0; pa ^= 1; }
        }
        if (leader) umma_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    tc_epilogue<COUT, PLANES>(p, total_tiles, tmem_base, tfull, tempty, s_scale, s_shift, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// =====================================================================================================
// "up2" kernel: outputs at TWICE the resolution of the main input, 8 output parity classes per low-res
// tile, every class a short list of taps that all read the SAME halo'd slabs of the low-res tile:
//   kind 0  ConvTranspose3d k3 s2 p1 op1 (o = 2i - 1 + k): slabs dz in {0,1}, tap offsets in {0,1}^3,
//           1/2/4/8 taps per class, weights = the 27 filter taps        (Multi_Aggregation.conv3, cva.py:20-22)
//   kind 1  trilinear x2 (align_corners=False) of a border-replicated low-res tensor: slabs dz in {0,1,2},
//           8 taps per class whose weights are the 4 diagonal matrices {27,9,3,1}/64 * I     (cva.py:64)
// plus an optional 1x1x1 "side" tap per class on the parity view of a tensor at the OUTPUT resolution
// (redir of cva.py:24 for kind 0; the cost half of cva.fuse, cva.py:55,69, for kind 1).
// One slab load per tile feeds all 8 classes (the per-tap kernel reloaded a 128-row box per tap and was
// L2->SMEM bound).  Weights stream per tap through a small ring; side boxes through a 2-deep ring.
// =====================================================================================================
struct Up2Tap { unsigned char slab, dy, dx, widx; };
struct Up2Params {
  int nslab; int slab_dz[3];
  int has_side; int side_widx;
  int wres;      // 1: all weight tiles (<= W_SLOTS) are loaded once and stay resident (kind 1: 5 tiles)
  int nw;
  unsigned char cls_tap0[9];
  Up2Tap taps[64];
};

template <int CIN, int PLANES, int NSLAB, int NSIDE = 2>
struct Up2Cfg {
  static constexpr int COUT = 32;
  static constexpr int ROWB = CIN * 2;
  static constexpr int SLAB_BYTES = HB_W * HB_H * ROWB;
  static constexpr int SLAB_PITCH = (SLAB_BYTES + 1023) / 1024 * 1024;
  static constexpr int SLAB_SET = NSLAB * PLANES * SLAB_PITCH;      // slabs x planes
  static constexpr int A_BYTES = TC_M * 128;                         // one plane of a side box: 128 w-PAIR rows of 128 B
  static constexpr int SIDE_SLOT = PLANES * A_BYTES;
  static constexpr int B_ROWS = PLANES * COUT;
  static constexpr int B_BYTES = B_ROWS * ROWB;
  static constexpr int SIDE_SLOTS = NSIDE;     // pair-row boxes in flight (a DRAM stream: the deeper, the more bytes in flight)
  static constexpr int OTHER = SIDE_SLOTS * SIDE_SLOT + 1024 + 512 + 2 * COUT * 4;
  static constexpr int NBUF = (2 * SLAB_SET + OTHER + 6 * B_BYTES <= 225 * 1024) ? 2 : 1;
  // weight ring: whatever is left, 6..12 taps deep (the streamed 64->32 parity taps are 8 KB each and one tap is only
  // ~200 tensor-pipe cycles of work, so the ring depth is the L2 latency the MMA warp can ride out)
  static constexpr int W_FIT = (225 * 1024 - NBUF * SLAB_SET - OTHER) / B_BYTES;
  static constexpr int W_SLOTS = W_FIT > 12 ? 12 : (W_FIT < 6 ? 6 : W_FIT);
  static constexpr int FIXED = OTHER + W_SLOTS * B_BYTES;
  static constexpr int TMEM_COLS = 2 * PLANES * COUT < 32 ? 32 : 2 * PLANES * COUT;
  static constexpr int SMEM_BYTES = NBUF * SLAB_SET + FIXED;
};

template <int CIN, int PLANES, int NSLAB, int NSIDE = 2>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_up2_kernel(const __grid_constant__ TcMaps maps, const TcParams p, const Up2Params u) {
  using Cfg = Up2Cfg<CIN, PLANES, NSLAB, NSIDE>;
  constexpr int COUT = Cfg::COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slab_base = smem;
  uint8_t* side_base = slab_base + Cfg::NBUF * Cfg::SLAB_SET;
  uint8_t* w_base = side_base + Cfg::SIDE_SLOTS * Cfg::SIDE_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + Cfg::W_SLOTS * Cfg::B_BYTES);
  uint64_t* sfull = bars;                          // [NBUF]
  uint64_t* sempty = sfull + Cfg::NBUF;            // [NBUF]
  uint64_t* dfull = sempty + Cfg::NBUF;            // [SIDE_SLOTS]
  uint64_t* dempty = dfull + Cfg::SIDE_SLOTS;      // [SIDE_SLOTS]
  uint64_t* wfull = dempty + Cfg::SIDE_SLOTS;      // [W_SLOTS]
  uint64_t* wempty = wfull + Cfg::W_SLOTS;         // [W_SLOTS]
  uint64_t* tfull = wempty + Cfg::W_SLOTS;         // [2]
  uint64_t* tempty = tfull + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(w_base + Cfg::W_SLOTS * Cfg::B_BYTES + 512);
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.Dt * p.tiles_h * p.tiles_w;      // low-res tiles; 8 classes inside each

  if (threadIdx.x < COUT) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::NBUF; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 1); }
    for (int i = 0; i < Cfg::SIDE_SLOTS; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dempty[i], 1); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
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
  pdl_trigger();                 // persistent single-wave grid: let the next kernel's CTAs take over SMs as ours retire
  if (warp != 0) pdl_wait();     // (the producer warp waits after its loads of static weights, see below)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t sb = 0, pb = 0, sd = 0, pd = 0, sw = 0, pw = 0;
      if (u.wres) {     // resident weights: slot i holds weight tile i for the whole kernel
        mbar_expect_tx(&wfull[0], u.nw * Cfg::B_BYTES);
        for (int i = 0; i < u.nw; ++i) tma_load_2d(w_base + i * Cfg::B_BYTES, &maps.w, &wfull[0], 0, i * Cfg::B_ROWS);
      }
      pdl_wait();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int td = r % p.Dt;
        const int b = r / p.Dt;
        mbar_wait(&sempty[sb], pb ^ 1);
        mbar_expect_tx(&sfull[sb], u.nslab * PLANES * Cfg::SLAB_BYTES);
        for (int sl = 0; sl < u.nslab; ++sl)
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_5d(slab_base + sb * Cfg::SLAB_SET + (sl * PLANES + pl) * Cfg::SLAB_PITCH, &maps.a[0], &sfull[sb], 0,
                        tw * TC_TW, th * TC_TH, td + u.slab_dz[sl], pl * p.B + b);
        if (++sb == Cfg::NBUF) { sb = 0; pb ^= 1; }
        for (int cls = 0; cls < p.ncls; ++cls) {
          if (u.has_side && (cls & 1) == 0) {
            // one box of 128-byte w-PAIR rows serves both x parities (classes cls, cls+1): contiguous full-line
            // requests instead of two element-strided boxes of 64-byte pieces
            mbar_wait(&dempty[sd], pd ^ 1);
            mbar_expect_tx(&dfull[sd], PLANES * Cfg::A_BYTES);
#pragma unroll
            for (int pl = 0; pl < PLANES; ++pl)
              tma_load_5d(side_base + sd * Cfg::SIDE_SLOT + pl * Cfg::A_BYTES, &maps.a[1 + (cls >> 1)], &dfull[sd], 0,
                          tw * TC_TW, th * TC_TH, td, pl * p.B + b);
            if (++sd == Cfg::SIDE_SLOTS) { sd = 0; pd ^= 1; }
          }
          const int nt = u.wres ? 0 : (u.cls_tap0[cls + 1] - u.cls_tap0[cls] + (u.has_side ? 1 : 0));
          for (int j = 0; j < nt; ++j) {
            const int widx = (j < u.cls_tap0[cls + 1] - u.cls_tap0[cls]) ? u.taps[u.cls_tap0[cls] + j].widx : u.side_widx;
            mbar_wait(&wempty[sw], pw ^ 1);
            mbar_expect_tx(&wfull[sw], Cfg::B_BYTES);
            tma_load_2d(w_base + sw * Cfg::B_BYTES, &maps.w, &wfull[sw], 0, widx * Cfg::B_ROWS);
            if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_full = make_idesc(TC_M, PLANES * COUT);
      constexpr uint32_t idesc_half = make_idesc(TC_M, COUT);
      constexpr uint64_t DA = desc_const<Cfg::ROWB>(HB_W * Cfg::ROWB);     // slab views: 8-row groups 10 rows apart
      constexpr uint64_t DD = desc_const<Cfg::ROWB>(8 * Cfg::ROWB);        // dense tiles (side boxes, weights)
      constexpr uint32_t NACC = PLANES * COUT;
      const uint32_t slab_u32 = smem_u32(slab_base), side_u32 = smem_u32(side_base), w_u32 = smem_u32(w_base);
      uint32_t sb = 0, pb = 0, sd = 0, pd = 0, sw = 0, pw = 0, it = 0;
      if (u.wres) { mbar_wait(&wfull[0], 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&sfull[sb], pb);
        tc_fence_after();
        const uint32_t set_u32 = slab_u32 + sb * Cfg::SLAB_SET;
        for (int cls = 0; cls < p.ncls; ++cls, ++it) {
          const uint32_t acc = it & 1;
          mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_addr = tmem_base + acc * NACC;
          const int t0 = u.cls_tap0[cls], t1 = u.cls_tap0[cls + 1];
          for (int t = t0; t < t1; ++t) {
            const Up2Tap tp = u.taps[t];
            if (!u.wres) { mbar_wait(&wfull[sw], pw); tc_fence_after(); }
            const uint32_t a0 = set_u32 + (uint32_t)(tp.slab * PLANES) * Cfg::SLAB_PITCH + (uint32_t)(tp.dy * HB_W + tp.dx) * Cfg::ROWB;
            const uint64_t da0 = DA + (a0 >> 4);
            const uint64_t db0 = DD + ((w_u32 + (u.wres ? (uint32_t)tp.widx : sw) * Cfg::B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k) {
              if (leader && !(p.dbg & 2)) {
                umma_bf16(d_addr, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc_full, (t > t0 || k > 0) ? 1u : 0u);
                if (PLANES == 2)
                  umma_bf16(d_addr + COUT, da0 + (uint64_t)((Cfg::SLAB_PITCH >> 4) + k * 2), db0 + (uint64_t)(k * 2), idesc_half, 1u);
              }
            }
            if (!u.wres) {
              __syncwarp();
              if (leader) umma_commit(&wempty[sw]);
              if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
            }
          }
          if (u.has_side) {
            if ((cls & 1) == 0) mbar_wait(&dfull[sd], pd);      // the pair box of classes (cls, cls+1)
            if (!u.wres) mbar_wait(&wfull[sw], pw);
            tc_fence_after();
            // A = pair rows (SWIZZLE_128B, dense); the odd-x class reads the second voxel of each pair: K-offset +64 B
            constexpr uint64_t DS = desc_const<128>(8 * 128);
            const uint64_t da0 = DS + ((side_u32 + sd * Cfg::SIDE_SLOT + (uint32_t)(cls & 1) * 64u) >> 4);
            const uint64_t db0 = DD + ((w_u32 + (u.wres ? (uint32_t)u.side_widx : sw) * Cfg::B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < 2; ++k) {                        // the side input has 32 channels
              if (leader && !(p.dbg & 2)) {
                umma_bf16(d_addr, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc_full, (t1 > t0 || k > 0) ? 1u : 0u);
                if (PLANES == 2)
                  umma_bf16(d_addr + COUT, da0 + (uint64_t)((Cfg::A_BYTES >> 4) + k * 2), db0 + (uint64_t)(k * 2), idesc_half, 1u);
              }
            }
            __syncwarp();
            if (leader) { if (!u.wres) umma_commit(&wempty[sw]); if (cls & 1) umma_commit(&dempty[sd]); }
            if (!u.wres) { if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; } }
            if (cls & 1) { if (++sd == Cfg::SIDE_SLOTS) { sd = 0; pd ^= 1; } }
          }
          if (leader) umma_commit(&tfull[acc]);
          __syncwarp();
        }
        if (leader) umma_commit(&sempty[sb]);      // all MMAs reading this slab set have been issued
        __syncwarp();
        if (++sb == Cfg::NBUF) { sb = 0; pb ^= 1; }
      }
    }
  } else {
    tc_epilogue<COUT, PLANES>(p, total_tiles, tmem_base, tfull, tempty, s_scale, s_shift, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================================
// ConvTranspose3d 64 -> 32 (+ redir side tap), TWO depth-adjacent low-res tiles per weight fetch
// (Multi_Aggregation.conv3 + redir, cva.py:20-31).
// The up2 kernel above streams 28 weight tiles of 8 KB per 128-voxel tile: 280 KB of weights against 92 KB of slabs and
// 131 KB of side rows -- it sits on the L2->SMEM limit (8.3 TB/s at 133 us) with the tensor pipe 25 % busy.  Here a work
// item is a PAIR of low-res tiles at depths (d, d+1):
//   * three slabs (d, d+1, d+2) feed both tiles (tap slab index s of tile 0 = slab s, of tile 1 = slab s+1), and the
//     slabs are 9 x 17 instead of 10 x 18 voxels (the taps of a transposed conv only reach offsets {0,1});
//   * every streamed weight tile is used by TWO MMAs chains (M = 2 x 128), and the side (redir) weights stay resident;
//   L2->SMEM bytes per tile: 503 KB -> 298 KB.  Four accumulator buffers (2 classes x 2 sub-tiles) in TMEM.
// Class order inside a pair: even x-parity class = taps then side, odd class = side then taps, so the pair-row box of a
// class pair is released early and the next one has >= 4 taps of MMA time to arrive.
// =====================================================================================================
constexpr int DP_W = TC_TW + 1, DP_H = TC_TH + 1;     // slab 9 x 17

template <int PLANES>
struct DeconvPairCfg {
  static constexpr int CIN = 64, COUT = 32, ROWB = 128;
  static constexpr int SLAB_BYTES = DP_W * DP_H * ROWB;
  static constexpr int SLAB_PITCH = (SLAB_BYTES + 1023) / 1024 * 1024;
  static constexpr int SLAB_SET = 3 * PLANES * SLAB_PITCH;
  static constexpr int A_BYTES = TC_M * 128;                       // one plane of a side box: 128 w-pair rows of 128 B
  static constexpr int SIDE_SLOT = PLANES * A_BYTES;
  static constexpr int B_ROWS = PLANES * COUT;
  static constexpr int B_BYTES = B_ROWS * ROWB;
  static constexpr int FIXED = SLAB_SET + 2 * SIDE_SLOT + B_BYTES + 1024 + 512 + 2 * COUT * 4;
  static constexpr int W_FIT = (227 * 1024 - FIXED) / B_BYTES;
  static constexpr int W_SLOTS = W_FIT > 12 ? 12 : W_FIT;
  static constexpr int NACC = PLANES * COUT;
  static constexpr int TMEM_COLS = 4 * NACC;
  static constexpr int SMEM_BYTES = FIXED + W_SLOTS * B_BYTES;
};

template <int PLANES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_deconv_pair_kernel(const __grid_constant__ TcMaps maps, const TcParams p, const Up2Params u) {
  using Cfg = DeconvPairCfg<PLANES>;
  constexpr int COUT = Cfg::COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slab_base = smem;
  uint8_t* side_base = slab_base + Cfg::SLAB_SET;
  uint8_t* wside = side_base + 2 * Cfg::SIDE_SLOT;                 // resident side (redir) weights
  uint8_t* w_base = wside + Cfg::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + Cfg::W_SLOTS * Cfg::B_BYTES);
  uint64_t* sfull = bars;                          // [1]
  uint64_t* sempty = sfull + 1;                    // [1]
  uint64_t* dfull = sempty + 1;                    // [2]
  uint64_t* dempty = dfull + 2;                    // [2]
  uint64_t* wsfull = dempty + 2;                   // [1]
  uint64_t* wfull = wsfull + 1;                    // [W_SLOTS]
  uint64_t* wempty = wfull + Cfg::W_SLOTS;         // [W_SLOTS]
  uint64_t* tfull = wempty + Cfg::W_SLOTS;         // [4]
  uint64_t* tempty = tfull + 4;                    // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);
  float* s_scale = reinterpret_cast<float*>(w_base + Cfg::W_SLOTS * Cfg::B_BYTES + 512);
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_pairs = p.B * p.Dt * p.tiles_h * p.tiles_w;      // p.Dt = depth pairs

  if (threadIdx.x < COUT) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    mbar_init(sfull, 1); mbar_init(sempty, 1); mbar_init(wsfull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dempty[i], 1); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
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
  pdl_trigger();                 // persistent single-wave grid: let the next kernel's CTAs take over SMs as ours retire
  if (warp != 0) pdl_wait();     // (the producer warp waits after its loads of static weights, see below)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(wsfull, Cfg::B_BYTES);
      tma_load_2d(wside, &maps.w, wsfull, 0, u.side_widx * Cfg::B_ROWS);
      pdl_wait();
      uint32_t ps = 0, pd = 0, sw = 0, pw = 0;
      for (int pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
        int r = pair;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int d0 = 2 * (r % p.Dt);
        const int b = r / p.Dt;
        mbar_wait(sempty, ps ^ 1);
        mbar_expect_tx(sfull, 3 * PLANES * Cfg::SLAB_BYTES);
        for (int sl = 0; sl < 3; ++sl)
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_5d(slab_base + (sl * PLANES + pl) * Cfg::SLAB_PITCH, &maps.a[0], sfull, 0, tw * TC_TW, th * TC_TH,
                        d0 + sl, pl * p.B + b);
        ps ^= 1;
        // (an L2 prefetch of this pair's side rows / the next pair's slabs was measured SLOWER: 163 vs 147 us -- the prefetch
        //  requests queue in the same TMA unit in front of the loads the MMA warp is waiting for)
        for (int cls = 0; cls < 8; ++cls) {
          if ((cls & 1) == 0) {
            // pair-row boxes of the two sub-tiles for the x-parity class pair (cls, cls+1): slot = sub-tile
            for (int sub = 0; sub < 2; ++sub) {
              mbar_wait(&dempty[sub], pd ^ 1);
              mbar_expect_tx(&dfull[sub], PLANES * Cfg::A_BYTES);
#pragma unroll
              for (int pl = 0; pl < PLANES; ++pl)
                tma_load_5d(side_base + sub * Cfg::SIDE_SLOT + pl * Cfg::A_BYTES, &maps.a[1 + (cls >> 1)], &dfull[sub], 0,
                            tw * TC_TW, th * TC_TH, d0 + sub, pl * p.B + b);
            }
            pd ^= 1;
          }
          for (int t = u.cls_tap0[cls]; t < u.cls_tap0[cls + 1]; ++t) {
            mbar_wait(&wempty[sw], pw ^ 1);
            mbar_expect_tx(&wfull[sw], Cfg::B_BYTES);
            tma_load_2d(w_base + sw * Cfg::B_BYTES, &maps.w, &wfull[sw], 0, u.taps[t].widx * Cfg::B_ROWS);
            if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_full = make_idesc(TC_M, PLANES * COUT);
    constexpr uint32_t idesc_half = make_idesc(TC_M, COUT);
    constexpr uint64_t DA = desc_const<128>(DP_W * 128);        // slab views: 8-row groups 9 rows apart
    constexpr uint64_t DD = desc_const<128>(8 * 128);           // dense tiles (side boxes, weights)
    constexpr uint32_t NACC = Cfg::NACC;
    const uint32_t slab_u32 = smem_u32(slab_base), side_u32 = smem_u32(side_base), w_u32 = smem_u32(w_base);
    const uint64_t db_side = DD + (smem_u32(wside) >> 4);
    uint32_t ps = 0, pd = 0, sw = 0, pw = 0, it = 0;
    mbar_wait(wsfull, 0);
    tc_fence_after();
    for (int pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
      mbar_wait(sfull, ps);
      tc_fence_after();
      ps ^= 1;
      for (int cls = 0; cls < 8; ++cls, it += 2) {
        const uint32_t acc0 = it & 3, acc1 = (it + 1) & 3;
        mbar_wait(&tempty[acc0], ((it >> 2) & 1) ^ 1);
        mbar_wait(&tempty[acc1], (((it + 1) >> 2) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dacc[2] = {tmem_base + acc0 * NACC, tmem_base + acc1 * NACC};
        const bool odd = (cls & 1) != 0;
        uint32_t started = 0;                                   // 0 until the first MMA of this class has been issued
        auto side_taps = [&]() {
          if (!odd) { mbar_wait(&dfull[0], pd); mbar_wait(&dfull[1], pd); tc_fence_after(); }
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            // A = pair rows (dense SWIZZLE_128B); the odd-x class reads the second voxel of each pair: K-offset +64 B
            const uint64_t da0 = DD + ((side_u32 + (uint32_t)sub * Cfg::SIDE_SLOT + (odd ? 64u : 0u)) >> 4);
#pragma unroll
            for (int k = 0; k < 2; ++k) {                       // the side input has 32 channels
              if (leader && !(p.dbg & 2)) {
                umma_bf16(dacc[sub], da0 + (uint64_t)(k * 2), db_side + (uint64_t)(k * 2), idesc_full, (started || k > 0) ? 1u : 0u);
                if (PLANES == 2)
                  umma_bf16(dacc[sub] + COUT, da0 + (uint64_t)((Cfg::A_BYTES >> 4) + k * 2), db_side + (uint64_t)(k * 2), idesc_half, 1u);
              }
            }
          }
          started = 1;
          if (odd) {                                            // both classes of the pair have read the boxes
            __syncwarp();
            if (leader) { umma_commit(&dempty[0]); umma_commit(&dempty[1]); }
            pd ^= 1;
          }
        };
        if (odd) side_taps();
        for (int t = u.cls_tap0[cls]; t < u.cls_tap0[cls + 1]; ++t) {
          const Up2Tap tp = u.taps[t];
          mbar_wait(&wfull[sw], pw);
          tc_fence_after();
          const uint64_t db0 = DD + ((w_u32 + sw * Cfg::B_BYTES) >> 4);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t a0 = slab_u32 + (uint32_t)((tp.slab + sub) * PLANES) * Cfg::SLAB_PITCH +
                                (uint32_t)(tp.dy * DP_W + tp.dx) * 128u;
            const uint64_t da0 = DA + (a0 >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (leader && !(p.dbg & 2)) {
                umma_bf16(dacc[sub], da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc_full, (started || k > 0) ? 1u : 0u);
                if (PLANES == 2)
                  umma_bf16(dacc[sub] + COUT, da0 + (uint64_t)((Cfg::SLAB_PITCH >> 4) + k * 2), db0 + (uint64_t)(k * 2), idesc_half, 1u);
              }
            }
          }
          started = 1;
          __syncwarp();
          if (leader) umma_commit(&wempty[sw]);
          if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
        }
        if (!odd) side_taps();
        __syncwarp();
        if (leader) { umma_commit(&tfull[acc0]); umma_commit(&tfull[acc1]); }
        __syncwarp();
      }
      if (leader) umma_commit(sempty);              // all MMAs reading the slabs of this pair have been issued
      __syncwarp();
    }
  } else {
    tc_epilogue<COUT, PLANES>(p, total_pairs, tmem_base, tfull, tempty, s_scale, s_shift, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// =====================================================================================================
// Depth-marching 3x3x3 stride-1 conv, Cout = 32: the three depth taps are folded into the GEMM's N.
// A work item = one (h, w) tile column x a chunk of n <= 16 output planes.  The CTA walks the n+2 input planes of
// the chunk; the halo slab of input plane d' is read ONCE per (kh, kw) and multiplied against the weights of all
// three kd at once (B rows [W_kd0; W_kd1; W_kd2], N = 96), landing in the accumulators of output planes
// d'+1, d', d'-1, which sit side by side in TMEM: output plane j of the chunk owns columns 32*(n-1-j), so every
// slab's three targets are one contiguous 96-column window that slides down by 32 columns per plane.
// Per tap the activation tile is read from shared memory once instead of three times (the per-tile kernels are
// SMEM-operand bound at N = 32/64).  Parity precision issues hi.Whi into the main window and hi.Wlo, lo.Whi into a
// second window MARCH_CORR columns further (so a chunk is at most 8 planes).
// The window is zeroed by the epilogue warps before an item starts (every MMA accumulates), each finished plane
// is announced through its own mbarrier, and the epilogue drains it while the MMA warp keeps marching.
// =====================================================================================================
template <int CIN, int PLANES>
struct MarchCfg {
  static constexpr int COUT = 32;
  static constexpr int ROWB = CIN * 2;
  static constexpr int SLAB_BYTES = HB_W * HB_H * ROWB;
  static constexpr int SLAB_PITCH = (SLAB_BYTES + 1023) / 1024 * 1024;
  static constexpr int A_SLOT = PLANES * SLAB_PITCH;
  static constexpr int WP_BYTES = 3 * COUT * ROWB;              // one plane of one (kh,kw): rows [kd][co]
  static constexpr int TAP_BYTES = PLANES * WP_BYTES;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr bool WRES = (9 * TAP_BYTES + 2 * A_SLOT) <= BUDGET;
  static constexpr int A_SLOTS_RES = ((BUDGET - 9 * TAP_BYTES) / A_SLOT) > 4 ? 4 : ((BUDGET - 9 * TAP_BYTES) / A_SLOT);
  static constexpr int A_SLOTS = WRES ? A_SLOTS_RES : 2;
  static constexpr int W_SLOTS_STR = ((BUDGET - 2 * A_SLOT) / TAP_BYTES) > 6 ? 6 : ((BUDGET - 2 * A_SLOT) / TAP_BYTES);
  static constexpr int W_SLOTS = WRES ? 1 : W_SLOTS_STR;
  static constexpr int W_BYTES_TOTAL = WRES ? 9 * TAP_BYTES : W_SLOTS * TAP_BYTES;
  static constexpr int TMEM_COLS = 512;
  static constexpr int MAXN = PLANES == 2 ? 8 : 16;          // parity: main + correction accumulators = 2 x 32 columns per plane
  static constexpr int SMEM_BYTES = A_SLOTS * A_SLOT + W_BYTES_TOTAL + 1024 + 512 + 2 * COUT * 4;
};

template <int CIN, int PLANES, int TAPS = 0>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_march_kernel(const __grid_constant__ TcMaps maps, const TcParams p,
                     const __grid_constant__ typename TapArg<TAPS>::type tw) {
  using Cfg = MarchCfg<CIN, PLANES>;
  constexpr int COUT = Cfg::COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;
  uint8_t* w_base = smem + Cfg::A_SLOTS * Cfg::A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + Cfg::W_BYTES_TOTAL);
  uint64_t* afull = bars;                        // [A_SLOTS]
  uint64_t* aempty = afull + Cfg::A_SLOTS;       // [A_SLOTS]
  uint64_t* wfull = aempty + Cfg::A_SLOTS;       // [W_SLOTS]
  uint64_t* wempty = wfull + Cfg::W_SLOTS;       // [W_SLOTS]
  uint64_t* oready = wempty + Cfg::W_SLOTS;      // [MAXN] output plane j of the current item is complete
  uint64_t* pzero = oready + Cfg::MAXN;          // [MAXN] accumulator plane j drained and cleared for the next item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pzero + Cfg::MAXN);
  float* s_scale = reinterpret_cast<float*>(w_base + Cfg::W_BYTES_TOTAL + 512);
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = p.march_n;
  const int total_items = p.B * p.Dt * p.tiles_h * p.tiles_w;      // p.Dt = number of depth chunks

  if (threadIdx.x < COUT) {
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < Cfg::MAXN; ++i) { mbar_init(&oready[i], 1); mbar_init(&pzero[i], TAPS ? 4 : 8); }
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
  pdl_trigger();                 // persistent single-wave grid: let the next kernel's CTAs take over SMs as ours retire
  if (warp != 0) pdl_wait();     // (the producer warp waits after its loads of static weights, see below)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (Cfg::WRES) {
        mbar_expect_tx(&wfull[0], 9 * Cfg::TAP_BYTES);
        for (int t = 0; t < 9; ++t)
          tma_load_2d(w_base + t * Cfg::TAP_BYTES, &maps.w, &wfull[0], 0, t * PLANES * 3 * COUT);
      }
      pdl_wait();
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int r = item;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int ch = r % p.Dt;
        const int b = r / p.Dt;
        for (int sidx = 0; sidx < n + 2; ++sidx) {
          const int dprime = ch * n - 1 + sidx;
          mbar_wait(&aempty[sa], pa ^ 1);
          mbar_expect_tx(&afull[sa], PLANES * Cfg::SLAB_BYTES);
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_5d(a_base + sa * Cfg::A_SLOT + pl * Cfg::SLAB_PITCH, &maps.a[0], &afull[sa], 0, tw * TC_TW - 1,
                        th * TC_TH - 1, dprime, pl * p.B + b);
          if (++sa == Cfg::A_SLOTS) { sa = 0; pa ^= 1; }
          if (!Cfg::WRES) {
            for (int t = 0; t < 9; ++t) {
              mbar_wait(&wempty[sw], pw ^ 1);
              mbar_expect_tx(&wfull[sw], Cfg::TAP_BYTES);
              tma_load_2d(w_base + sw * Cfg::TAP_BYTES, &maps.w, &wfull[sw], 0, t * PLANES * 3 * COUT);
              if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      const bool leader = elect_one();
      constexpr uint64_t DA = desc_const<Cfg::ROWB>(HB_W * Cfg::ROWB);
      constexpr uint64_t DB = desc_const<Cfg::ROWB>(8 * Cfg::ROWB);
      if (Cfg::WRES) { mbar_wait(&wfull[0], 0); tc_fence_after(); }
      const uint32_t a_u32 = smem_u32(a_base), w_u32 = smem_u32(w_base);
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0, wi = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++wi) {
        const int ch = (item / (p.tiles_w * p.tiles_h)) % p.Dt;
#pragma unroll 1
        for (int sidx = 0; sidx < n + 2; ++sidx) {
          if (sidx < n) {                          // slab sidx is the first to write output plane sidx of this item:
            mbar_wait(&pzero[sidx], wi & 1);       // drained (previous item) and cleared by the epilogue warps
            tc_fence_after();
          }
          const int dprime = ch * n - 1 + sidx;
          // output index inside the chunk fed through kd: j = sidx - kd, must lie in [0, n)
          const int kd_lo = max(0, sidx - n + 1), kd_hi = min(2, sidx);
          const int nkd = kd_hi - kd_lo + 1;
          const uint32_t idesc = make_idesc(TC_M, 32 * nkd);
          const uint32_t d_addr = tmem_base + (uint32_t)(32 * (n - 1 - (sidx - kd_lo)));
          const uint32_t w_off = (uint32_t)(kd_lo * COUT) * Cfg::ROWB;          // skip the kd blocks that fall outside
          const bool live = dprime >= 0 && dprime < p.Do;                       // (OOB planes are all zeros: skip the MMAs)
          mbar_wait(&afull[sa], pa);
          tc_fence_after();
          const uint64_t da_slab = DA + ((a_u32 + sa * Cfg::A_SLOT) >> 4);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              uint32_t wtap;
              if (Cfg::WRES) {
                wtap = w_u32 + (uint32_t)(kh * 3 + kw) * Cfg::TAP_BYTES;
              } else {
                mbar_wait(&wfull[sw], pw);
                tc_fence_after();
                wtap = w_u32 + sw * Cfg::TAP_BYTES;
              }
              const uint64_t db_hi = DB + ((wtap + w_off) >> 4);
              const uint64_t db_lo = DB + ((wtap + Cfg::WP_BYTES + w_off) >> 4);
#pragma unroll
              for (int k = 0; k < CIN / 16; ++k) {
                const uint64_t da = da_slab + (uint64_t)(((kh * HB_W + kw) * Cfg::ROWB + k * 32) >> 4);
                if (leader && live && !(p.dbg & 2)) {
                  umma_bf16(d_addr, da, db_hi + (uint64_t)(k * 2), idesc, 1u);
                  if (PLANES == 2) {     // the two small terms accumulate in their own block (truncation bias, see the header)
                    umma_bf16(d_addr + MARCH_CORR, da, db_lo + (uint64_t)(k * 2), idesc, 1u);
                    umma_bf16(d_addr + MARCH_CORR, da + (uint64_t)(Cfg::SLAB_PITCH >> 4), db_hi + (uint64_t)(k * 2), idesc, 1u);
                  }
                }
              }
              if (!Cfg::WRES) {
                __syncwarp();
                if (leader) umma_commit(&wempty[sw]);
                if (++sw == Cfg::W_SLOTS) { sw = 0; pw ^= 1; }
              }
            }
          }
          __syncwarp();
          if (leader) {
            umma_commit(&aempty[sa]);
            if (sidx >= 2) umma_commit(&oready[sidx - 2]);      // output plane sidx-2 has received all three kd
          }
          __syncwarp();
          if (++sa == Cfg::A_SLOTS) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else {
    if constexpr (TAPS) tc_epilogue_taps<PLANES>(p, tw, total_items, tmem_base, oready, pzero, s_scale, s_shift, warp, lane);
    else tc_epilogue<COUT, PLANES>(p, total_items, tmem_base, oready, pzero, s_scale, s_shift, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// march weight pack: [tap (kh,kw) 9][plane][kd 3][co 32][ci] bf16
__global__ void pack_weight_tc_march_kernel(const float* __restrict__ w, int Ci, __nv_bfloat16* __restrict__ out, int planes) {
  const int total = 27 * 32 * Ci;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Ci, co = (i / Ci) % 32, t27 = i / (Ci * 32);
    const int kd = t27 / 9, khw = t27 % 9;
    const float v = w[((size_t)co * Ci + ci) * 27 + t27];
    uint32_t lo;
    const uint32_t hi = split_bf16(v, lo);
    const size_t row_hi = ((size_t)(khw * planes + 0) * 3 + kd) * 32 + co;
    out[row_hi * Ci + ci] = __ushort_as_bfloat16((unsigned short)hi);
    if (planes == 2) {
      const size_t row_lo = ((size_t)(khw * planes + 1) * 3 + kd) * 32 + co;
      out[row_lo * Ci + ci] = __ushort_as_bfloat16((unsigned short)lo);
    }
  }
}

template <int CIN, int PLANES, int TAPS = 0>
static int launch_tc_march(const TcMaps& maps, const TcParams& p, cudaStream_t st,
                           const typename TapArg<TAPS>::type* tw = nullptr) {
  using Cfg = MarchCfg<CIN, PLANES>;
  static_assert(Cfg::A_SLOTS >= 2 && Cfg::W_SLOTS >= 1, "smem plan");
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "smem plan");
  g_num_sms = dca_num_sms();
  cudaFuncSetAttribute(conv_tc_march_kernel<CIN, PLANES, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  const int total = p.B * p.Dt * p.tiles_h * p.tiles_w;
  const int grid = total < g_num_sms ? total : g_num_sms;
  TcParams q = p; fill_recips(q);
  for (int c = 0; c < 8; ++c) fill_comp(q, c, 27 * (CIN / 16));
  typename TapArg<TAPS>::type none{};
  dca_launch(conv_tc_march_kernel<CIN, PLANES, TAPS>, grid, TC_THREADS, Cfg::SMEM_BYTES, st, maps, q, tw ? *tw : none);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
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
                         size_t sD, size_t sB, int box_w = TC_TW, int box_h = TC_TH, int box_c = 0) {
  if (box_c == 0) box_c = C;   // box_c > C: the extra channels are out of bounds -> zero filled
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NB};
  cuuint64_t strides[4] = {sW * 2, sH * 2, sD * 2, sB * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = (box_c * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  return enc(m, (DCA_F16_PLANES ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 5, const_cast<void*>(base), dims, strides, box, es,
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
  return enc(m, (DCA_F16_PLANES ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_use_halo = 1;
static int g_use_s2slab = 1;
static int g_dbg = 0;

template <int CIN, int COUT, int PLANES>
static int launch_tc(const TcMaps& maps, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<CIN, COUT, PLANES>;
  static_assert(Cfg::STAGES >= 2, "pipeline too shallow");
  g_num_sms = dca_num_sms();
  cudaFuncSetAttribute(conv_tc_kernel<CIN, COUT, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       Cfg::SMEM_BYTES);
  const int total = p.B * p.Dt * p.tiles_h * p.tiles_w * p.ncls;
  const int grid = total < g_num_sms ? total : g_num_sms;
  TcParams q = p; fill_recips(q);
  for (int c = 0; c < 8; ++c)
    fill_comp(q, c, (c < q.ncls ? (int)q.cls_tap0[c + 1] - (int)q.cls_tap0[c] : 0) * (CIN / 16));
  dca_launch(conv_tc_kernel<CIN, COUT, PLANES>, grid, TC_THREADS, Cfg::SMEM_BYTES, st, maps, q);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

template <int PLANES>
static int launch_tc_s2slab(const TcMaps& maps, const TcParams& p, cudaStream_t st) {
  using Cfg = S2Cfg<PLANES>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "smem plan");
  g_num_sms = dca_num_sms();
  cudaFuncSetAttribute(conv_tc_s2slab_kernel<PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  const int total = p.B * p.Dt * p.tiles_h * p.tiles_w;
  const int grid = total < g_num_sms ? total : g_num_sms;
  TcParams q = p; fill_recips(q);
  for (int c = 0; c < 8; ++c) fill_comp(q, c, 27 * 2);
  dca_launch(conv_tc_s2slab_kernel<PLANES>, grid, TC_THREADS, Cfg::SMEM_BYTES, st, maps, q);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

template <int CIN, int COUT, int PLANES, int TAPS = 0>
static int launch_tc_halo(const TcMaps& maps, const TcParams& p, cudaStream_t st,
                          const typename TapArg<TAPS>::type* tw = nullptr) {
  using Cfg = HaloCfg<CIN, COUT, PLANES>;
  static_assert(Cfg::A_SLOTS >= 2 && Cfg::W_SLOTS >= 1, "smem plan");
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "smem plan");
  g_num_sms = dca_num_sms();
  cudaFuncSetAttribute(conv_tc_halo_kernel<CIN, COUT, PLANES, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       Cfg::SMEM_BYTES);
  const int total = p.B * p.Dt * p.tiles_h * p.tiles_w;
  const int grid = total < g_num_sms ? total : g_num_sms;
  TcParams q = p; fill_recips(q);
  for (int c = 0; c < 8; ++c) fill_comp(q, c, q.nslab * 9 * (CIN / 16));
  typename TapArg<TAPS>::type none{};
  dca_launch(conv_tc_halo_kernel<CIN, COUT, PLANES, TAPS>, grid, TC_THREADS, Cfg::SMEM_BYTES, st, maps, q, tw ? *tw : none);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}


// ------------------------------------------------------------------ 2-D 3x3 convs (PropgationNet_4x.conv, gwcnet_dca_g.py:112-115)
// weight pack for the 2-D path: [Cout chunk of T][channel slab of T][tap 9][plane][T][T] 16-bit (zero padded); T = 64, or 32
// for the 32-channel layers of the 1/2-resolution stem
__global__ void pack_weight_tc2d_kernel(const float* __restrict__ w, int Co, int Ci, __nv_bfloat16* __restrict__ out,
                                        int planes, int nchunk, int nslab, int T) {
  const int TT = T * T, total = nchunk * nslab * 9 * TT;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % T, co = (i / T) % T, t = (i / TT) % 9, sl = (i / (TT * 9)) % nslab, j = i / (TT * 9 * nslab);
    const int gco = j * T + co, gci = sl * T + ci;
    const float v = (gco < Co && gci < Ci) ? w[((size_t)gco * Ci + gci) * 9 + t] : 0.f;
    uint32_t lo;
    const uint32_t hi = split_bf16(v, lo);
    const size_t tapbase = ((size_t)(j * nslab + sl) * 9 + t) * planes;
    out[((tapbase + 0) * T + co) * T + ci] = __ushort_as_bfloat16((unsigned short)hi);
    if (planes == 2) out[((tapbase + 1) * T + co) * T + ci] = __ushort_as_bfloat16((unsigned short)lo);
  }
}

template <int CIN, int PLANES, int NSLAB, int NSIDE = 2>
static int launch_up2(const TcMaps& maps, const TcParams& p, const Up2Params& u, cudaStream_t st) {
  using Cfg = Up2Cfg<CIN, PLANES, NSLAB, NSIDE>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "smem plan");
  g_num_sms = dca_num_sms();
  cudaFuncSetAttribute(conv_tc_up2_kernel<CIN, PLANES, NSLAB, NSIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  const int total = p.B * p.Dt * p.tiles_h * p.tiles_w;
  const int grid = total < g_num_sms ? total : g_num_sms;
  TcParams q = p; fill_recips(q);
  for (int c = 0; c < 8; ++c)
    fill_comp(q, c, (c < q.ncls ? (int)u.cls_tap0[c + 1] - (int)u.cls_tap0[c] : 0) * (CIN / 16) + (u.has_side ? 2 : 0));
  dca_launch(conv_tc_up2_kernel<CIN, PLANES, NSLAB, NSIDE>, grid, TC_THREADS, Cfg::SMEM_BYTES, st, maps, q, u);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

// =====================================================================================================
// AvgPool3d(3, stride 2, pad 1, count_include_pad) of a 32-channel cost tensor (cva.downsample, cva.py:39).
// The thread-per-output kernel re-reads every input element 27/8 times through L1/L2 (ncu: 428 MB L2->SM for a 184 MB
// tensor, 6.6 TB/s = the fabric limit, 65 us).  Here a CTA owns an (8 w x 4 h) OUTPUT tile and marches along depth: every
// input plane of its (17 x 9) footprint is TMA-loaded exactly once (out-of-tensor coordinates are zero filled = the
// pool's zero padding), reduced 3x3 -> 1 in registers, and the three plane sums of an output depth are combined on the
// fly (an odd input plane is the last plane of output d and the first of d+1).  L2->SM traffic 1.2x the tensor.
// thread = (output voxel, 8-channel chunk); 2 slots of (hi, lo) boxes; several CTAs per SM hide the load latency.
// =====================================================================================================
constexpr int AP_TW = 8, AP_TH = 4, AP_BW = 2 * AP_TW + 1, AP_BH = 2 * AP_TH + 1;
constexpr int AP_BOX = AP_BW * AP_BH * 64;          // one plane (hi or lo) of one input plane's footprint, 32 ch x 2 B rows
constexpr int AP_PITCH = (AP_BOX + 127) / 128 * 128; // TMA destinations are 128-byte aligned

template <int PLANES>
__global__ void __launch_bounds__(AP_TW * AP_TH * 4)
avgpool3d_march_kernel(const __grid_constant__ CUtensorMap xmap, __nv_bfloat16* __restrict__ y, int B, int Do, int Ho,
                       int Wo, int tiles_w, int tiles_h, int nsplit) {
  extern __shared__ uint8_t ap_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ap_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full[2];
  const int tid = threadIdx.x;
  int r = blockIdx.x;
  const int tw = r % tiles_w; r /= tiles_w;
  const int th = r % tiles_h; r /= tiles_h;
  const int sp = r % nsplit;
  const int b = r / nsplit;
  const int dper = (Do + nsplit - 1) / nsplit;
  const int da = sp * dper, db = min(Do, da + dper);
  if (da >= db) return;
  const int nplanes = 2 * (db - da) + 1;           // input planes 2*da - 1 .. 2*db - 1
  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&xmap);
  }
  __syncthreads();
  auto issue = [&](int i) {
    uint8_t* dst = smem + (size_t)(i & 1) * PLANES * AP_PITCH;
    mbar_expect_tx(&full[i & 1], PLANES * AP_BOX);
#pragma unroll
    for (int pl = 0; pl < PLANES; ++pl)
      tma_load_5d(dst + pl * AP_PITCH, &xmap, &full[i & 1], 0, 2 * tw * AP_TW - 1, 2 * th * AP_TH - 1, 2 * da - 1 + i, pl * B + b);
  };
  pdl_wait();
  if (tid == 0) { issue(0); if (nplanes > 1) issue(1); }
  const int c8 = tid & 3, vox = tid >> 2;
  const int ow = vox % AP_TW, oh = vox / AP_TW;
  const int gw = tw * AP_TW + ow, gh = th * AP_TH + oh;
  const bool valid = gw < Wo && gh < Ho;
  const size_t yplane = (size_t)B * Do * Ho * Wo * 32;
  float2 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
  for (int i = 0; i < nplanes; ++i) {
    mbar_wait(&full[i & 1], (uint32_t)((i >> 1) & 1));
    const uint32_t src = smem_u32(smem) + (uint32_t)(i & 1) * PLANES * AP_PITCH;
    float2 rs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) rs[k] = make_float2(0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int off = ((2 * oh + dy) * AP_BW + (2 * ow + dx)) * 64 + c8 * 16;
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) {
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + pl * AP_PITCH + off));
          float f[8];
          unpack8(v, f);
#pragma unroll
          for (int k = 0; k < 4; ++k) rs[k] = __fadd2_rn(rs[k], make_float2(f[2 * k], f[2 * k + 1]));
        }
      }
    // WAR on the slot across proxies: the refill below is an ASYNC-proxy write, which is not ordered behind shared loads
    // that were issued but have not returned yet, and ptxas sinks the last (register-only) adds below a plain BAR.SYNC.
    // About one launch in 200 at KITTI size then had one voxel summed from a half-refilled slot -- always the last lanes
    // of a warp, whose loads are served last.  The barrier therefore takes a predicate computed from every plane sum:
    // "arrived at the barrier" now means "all my loads of this slot have returned".
    float chk = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) chk += rs[k].x + rs[k].y;
    (void)__syncthreads_or(chk != chk);             // every thread has read this slot
    if (tid == 0 && i + 2 < nplanes) issue(i + 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = __fadd2_rn(acc[k], rs[k]);
    if ((i & 1) == 0 && i >= 2) {                   // plane 2d+1 closes output depth d = da + i/2 - 1 ...
      if (valid) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[2 * k] = acc[k].x * (1.0f / 27.0f); o[2 * k + 1] = acc[k].y * (1.0f / 27.0f); }
        const int od = da + (i >> 1) - 1;
        store8<PLANES>(y, yplane, ((((size_t)b * Do + od) * Ho + gh) * Wo + gw) * 32 + c8 * 8, o);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = rs[k];   // ... and opens d + 1
    }
  }
}

extern "C" int dca_avgpool3d_simple(const void* x, void* y, int planes, int B, int C, int Di, int Hi, int Wi, void* stream);

template <int PLANES>
static int launch_deconv_pair(const TcMaps& maps, const TcParams& p, const Up2Params& u, cudaStream_t st) {
  using Cfg = DeconvPairCfg<PLANES>;
  static_assert(Cfg::W_SLOTS >= 3, "weight ring too shallow");
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "smem plan");
  static_assert(Cfg::TMEM_COLS <= 512 && (Cfg::TMEM_COLS & (Cfg::TMEM_COLS - 1)) == 0, "TMEM plan");
  g_num_sms = dca_num_sms();
  cudaFuncSetAttribute(conv_tc_deconv_pair_kernel<PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  const int total = p.B * p.Dt * p.tiles_h * p.tiles_w;
  const int grid = total < g_num_sms ? total : g_num_sms;
  TcParams q = p; fill_recips(q);
  for (int c = 0; c < 8; ++c) fill_comp(q, c, ((int)u.cls_tap0[c + 1] - (int)u.cls_tap0[c]) * 4 + 2);
  dca_launch(conv_tc_deconv_pair_kernel<PLANES>, grid, TC_THREADS, Cfg::SMEM_BYTES, st, maps, q, u);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

static int g_use_deconv_pair = 1;
static int g_use_pool_march = 1;
static int g_up2_side_slots = 2;   // (4 was measured: 91.2 us either way -- the stream is not ring-depth bound)
static int g_pool_ctas_per_sm = 3;   // measured at KITTI (r2f kernel: ld.shared + predicate barrier): 2: 38.7 us, 3: 38.0, 4: 39.1, 5-6: 47.2, 8: 40.4, 12: 40.9 (thread-per-output kernel: 66 us)

}  // namespace dca

using namespace dca;

// 1 (default): transposed conv 64 -> 32 with a side input runs the two-tiles-per-weight-fetch kernel; 0: the up2 kernel
extern "C" int dca_tc_set_deconv_pair(int on) { g_use_deconv_pair = on ? 1 : 0; return DCA_OK; }
// side-box ring depth of dca_up2_tc kind 2 (2 or 4; timing experiments)
extern "C" int dca_tc_set_up2_side_slots(int n) { g_up2_side_slots = (n == 4) ? 4 : 2; return DCA_OK; }

// kind 0: y = act(scale * (conv_transpose3d_k3s2(x) + side 1x1x1) + shift) + res_post
//         x [P][B][Dl][Hl][Wl][Cin], w_tc = 27 taps (+ tap 27 = side weights, Cin-padded) of [planes][32][Cin]
// kind 1: y = act(scale * (trilinear_x2(x) + side 1x1x1) + shift) + res_post
//         x = BORDER-REPLICATED low-res tensor [P][B][Dl+2][Hl+2][Wl+2][32]; w_tc = 5 taps:
//         {27,9,3,1}/64 * I and the side weights.
// side [P][B][2Dl][2Hl][2Wl][side_c] (optional for kind 0), Cout = 32.
extern "C" int dca_up2_tc(int kind, const void* x, int planes, const void* side, int side_c, const void* w_tc,
                          const float* scale, const float* shift, const void* res_post, int planes_res, void* y, int act,
                          int B, int Cin, int Dl, int Hl, int Wl, void* stream) {
  if (!x || !w_tc || !y || B <= 0 || planes < 1 || planes > 2 || Dl <= 0 || Hl <= 0 || Wl <= 0) return DCA_ERR_ARG;
  if (kind < 0 || kind > 2 || (kind >= 1 && (!side || Cin != 32)) || (Cin != 32 && Cin != 64)) return DCA_ERR_UNSUPPORTED;
  if (side && side_c != 32) return DCA_ERR_UNSUPPORTED;       // the side input is read as 128-byte w-pair rows
  cudaStream_t st = (cudaStream_t)stream;
  const int P = planes, Cout = 32;
  // kind 2: x is already at the OUTPUT depth (Dl = output depth); only H and W double
  const int Do = (kind == 2) ? Dl : 2 * Dl, Ho = 2 * Hl, Wo = 2 * Wl;
  TcMaps maps;
  TcParams p;
  Up2Params u;
  memset(&p, 0, sizeof(p));
  memset(&u, 0, sizeof(u));
  p.B = B; p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  p.scale = scale; p.shift = shift;
  p.res_post = (const __nv_bfloat16*)res_post;
  p.res_plane = (size_t)B * Do * Ho * Wo * Cout; p.planes_res = planes_res;
  p.y = (__nv_bfloat16*)y; p.y_plane = p.res_plane; p.planes_out = P; p.act = act;
  p.dbg = g_dbg;
  p.ldc = Cout; p.cout_valid = Cout;
  p.Dt = Dl; p.Ht = Hl; p.Wt = Wl; p.out_stride = 2; p.ncls = (kind == 2) ? 4 : 8; p.cls_inner = 1;
  if (kind == 2) p.out_stride_d = 1;
  p.tiles_w = (Wl + TC_TW - 1) / TC_TW; p.tiles_h = (Hl + TC_TH - 1) / TC_TH;
  // main input: halo box 10 x 18 whose origin is the tile origin (kinds 1/2: the +1 of the padding cancels the -1 halo)
  const int pad = (kind >= 1) ? 2 : 0;
  const int Dx = Dl + (kind == 1 ? 2 : 0), Hx = Hl + pad, Wx = Wl + pad;
  if (!make_act_map(&maps.a[0], x, Cin, Wx, Hx, Dx, P * B, (size_t)Cin, (size_t)Wx * Cin, (size_t)Hx * Wx * Cin,
                    (size_t)Dx * Hx * Wx * Cin, HB_W, HB_H))
    return DCA_ERR_LAUNCH;
  for (int i = 1; i < 9; ++i) maps.a[i] = maps.a[0];
  if (side) {
    // w-PAIR views of the side input, one per (z, y) parity: dims (64 = 2 voxels x 32 channels, Wl, Hl, Dl, planes*B)
    const size_t qH = (size_t)Wo * side_c, qD = (size_t)Ho * Wo * side_c, qB = (size_t)Do * Ho * Wo * side_c;
    for (int pq = 0; pq < p.ncls / 2; ++pq) {
      const int pz = (kind == 2) ? 0 : (pq >> 1) & 1, py = pq & 1;
      const __nv_bfloat16* base = (const __nv_bfloat16*)side + pz * qD + py * qH;
      if (!make_act_map(&maps.a[1 + pq], base, 64, Wl, Hl, Dl, P * B, (size_t)64, 2 * qH, (kind == 2 ? 1 : 2) * qD, qB, TC_TW,
                        TC_TH, 64))
        return DCA_ERR_LAUNCH;
    }
    u.has_side = 1;
  }
  int ntaps = 0;
  if (kind == 0) {
    u.nslab = 2; u.slab_dz[0] = 0; u.slab_dz[1] = 1;
    u.side_widx = 27;
    for (int pc = 0; pc < 8; ++pc) {
      const int par[3] = {(pc >> 2) & 1, (pc >> 1) & 1, pc & 1};
      u.cls_tap0[pc] = (unsigned char)ntaps;
      for (int a = 0; a < 3; ++a) p.cls_off[pc][a] = (signed char)par[a];
      for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
        const int k[3] = {kd, kh, kw};
        int off[3]; bool ok = true;
        for (int a = 0; a < 3; ++a) { const int n = par[a] + 1 - k[a]; if (n & 1) ok = false; off[a] = n / 2; }
        if (!ok) continue;
        Up2Tap& t = u.taps[ntaps++];
        t.slab = (unsigned char)off[0]; t.dy = (unsigned char)off[1]; t.dx = (unsigned char)off[2];
        t.widx = (unsigned char)((kd * 3 + kh) * 3 + kw);
      }
    }
  } else {
    u.nslab = 3; u.slab_dz[0] = 0; u.slab_dz[1] = 1; u.slab_dz[2] = 2;
    u.side_widx = 4; u.wres = 1; u.nw = 5;
    for (int pc = 0; pc < 8; ++pc) {
      const int par[3] = {(pc >> 2) & 1, (pc >> 1) & 1, pc & 1};
      u.cls_tap0[pc] = (unsigned char)ntaps;
      for (int a = 0; a < 3; ++a) p.cls_off[pc][a] = (signed char)par[a];
      // per axis: even output 2i: (in[i-1], .25), (in[i], .75); odd 2i+1: (in[i], .75), (in[i+1], .25);
      // in padded coordinates in[j] sits at j+1, so the two source offsets are (0,1) for even, (1,2) for odd
      for (int m = 0; m < 8; ++m) {
        int n75 = 0, off[3];
        for (int a = 0; a < 3; ++a) {
          const int second = (m >> (2 - a)) & 1;                 // 0 = first source of the axis, 1 = second
          off[a] = par[a] + second;
          const bool is75 = par[a] ? (second == 0) : (second == 1);
          n75 += is75 ? 1 : 0;
        }
        Up2Tap& t = u.taps[ntaps++];
        t.slab = (unsigned char)off[0]; t.dy = (unsigned char)off[1]; t.dx = (unsigned char)off[2];
        t.widx = (unsigned char)(3 - n75);                       // 0: 27/64, 1: 9/64, 2: 3/64, 3: 1/64
      }
    }
  }
  if (kind == 2) {
    // bilinear x2 in (h, w) of a tensor already interpolated along depth: 4 classes x 4 taps, weights {9,3,1}/16 * I
    ntaps = 0;
    memset(u.taps, 0, sizeof(u.taps));
    u.nslab = 1; u.slab_dz[0] = 0;
    u.side_widx = 3; u.wres = 1; u.nw = 4;
    for (int pc = 0; pc < 4; ++pc) {
      const int par[2] = {(pc >> 1) & 1, pc & 1};
      u.cls_tap0[pc] = (unsigned char)ntaps;
      p.cls_off[pc][0] = 0; p.cls_off[pc][1] = (signed char)par[0]; p.cls_off[pc][2] = (signed char)par[1];
      for (int m = 0; m < 4; ++m) {
        int n75 = 0, off[2];
        for (int a = 0; a < 2; ++a) {
          const int second = (m >> (1 - a)) & 1;
          off[a] = par[a] + second;
          const bool is75 = par[a] ? (second == 0) : (second == 1);
          n75 += is75 ? 1 : 0;
        }
        Up2Tap& t = u.taps[ntaps++];
        t.slab = 0; t.dy = (unsigned char)off[0]; t.dx = (unsigned char)off[1];
        t.widx = (unsigned char)(2 - n75);                       // 0: 9/16, 1: 3/16, 2: 1/16
      }
    }
    for (int pc = 4; pc <= 8; ++pc) u.cls_tap0[pc] = (unsigned char)ntaps;
  } else {
    u.cls_tap0[8] = (unsigned char)ntaps;
  }
  const int wt = (kind == 0) ? (side ? 28 : 27) : (kind == 1 ? 5 : 4);
  if (!make_w_map(&maps.w, w_tc, Cin, wt * P * Cout, P * Cout)) return DCA_ERR_LAUNCH;
  if (kind == 0 && Cin == 64 && side && g_use_deconv_pair) {
    // two depth-adjacent tiles per weight fetch over 9 x 17 slabs (conv_tc_deconv_pair_kernel)
    if (!make_act_map(&maps.a[0], x, Cin, Wx, Hx, Dx, P * B, (size_t)Cin, (size_t)Wx * Cin, (size_t)Hx * Wx * Cin,
                      (size_t)Dx * Hx * Wx * Cin, DP_W, DP_H))
      return DCA_ERR_LAUNCH;
    p.Dt = (Dl + 1) / 2; p.pair = 1;
    return P == 2 ? launch_deconv_pair<2>(maps, p, u, st) : launch_deconv_pair<1>(maps, p, u, st);
  }
  if (kind == 1) return P == 2 ? launch_up2<32, 2, 3>(maps, p, u, st) : launch_up2<32, 1, 3>(maps, p, u, st);
  if (kind == 2) {      // one slab per tile: the smem goes into a deeper ring of side boxes (the `cost` stream from DRAM)
    if (g_up2_side_slots == 4) return P == 2 ? launch_up2<32, 2, 1, 4>(maps, p, u, st) : launch_up2<32, 1, 1, 4>(maps, p, u, st);
    return P == 2 ? launch_up2<32, 2, 2>(maps, p, u, st) : launch_up2<32, 1, 2>(maps, p, u, st);
  }
  if (Cin == 64) return P == 2 ? launch_up2<64, 2, 2>(maps, p, u, st) : launch_up2<64, 1, 2>(maps, p, u, st);
  return P == 2 ? launch_up2<32, 2, 2>(maps, p, u, st) : launch_up2<32, 1, 2>(maps, p, u, st);
}

// 1 = halo-slab main loop for k3 s1 (default), 0 = one TMA box per tap (v1)
// kappa of the accumulator-truncation compensation (default 1.56e-8 per MMA step; 0 = off)
extern "C" int dca_tc_set_trunc_comp(float kappa) { g_trunc_kappa = kappa; return DCA_OK; }

extern "C" int dca_tc_set_halo(int on) { g_use_halo = on & 1; g_use_s2slab = (on & 2) ? 0 : 1; return DCA_OK; }
// accumulator interleave (1,2,4) and separate lo block (0/1) of the halo kernel
// timing probes of the tcgen05 kernels: (flags >> 4) & 1 skips the epilogue math + stores, & 2 the MMAs of the halo kernel
// (the first argument is reserved and must be 1)
extern "C" int dca_tc_set_tuning(int reserved, int flags) {
  if (reserved != 1) return DCA_ERR_ARG;
  g_dbg = (flags >> 4) & 7;
  return DCA_OK;
}

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
                             const float* shift, const void* res_pre, const void* res_post, int planes_res,
                             const void* up, int planes_up, const void* side, int side_c, void* y, int planes_out,
                             int act, int B, int Cin, int Cout, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                             void* stream) {
  // side (optional, transposed conv only): cost planes [P][B][Do][Ho][Wo][side_c] whose 1x1x1 conv is accumulated into
  // the same GEMM; its weights are tap 27 of w_tc ([planes][Cout][Cin], zero padded beyond side_c).
  if (side && (mode != 2 || side_c <= 0 || side_c > Cin || (side_c % 8) != 0)) return DCA_ERR_ARG;
  if (up && (Cout != 32 || (Do & 1) || (Ho & 1) || (Wo & 1) || planes_up < 1 || planes_up > 2)) return DCA_ERR_ARG;
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
  p.dbg = g_dbg;
  p.up = (const __nv_bfloat16*)up; p.planes_up = planes_up;
  p.nslab = 3; p.slab_dz[0] = -1; p.slab_dz[1] = 0; p.slab_dz[2] = 1;
  p.ldc = Cout; p.cout_valid = Cout;
  const size_t sW = Cin, sH = (size_t)Wi * Cin, sD = (size_t)Hi * Wi * Cin, sB = (size_t)Di * Hi * Wi * Cin;
  const int ntaps_total = (mode == 3) ? 1 : (side ? 28 : 27);
  if (!make_w_map(&maps.w, w_tc, Cin, ntaps_total * P * Cout, P * Cout)) return DCA_ERR_LAUNCH;

  p.ncls = 1;
  auto run = [&]() -> int {
    p.tiles_w = (p.Wt + TC_TW - 1) / TC_TW;
    p.tiles_h = (p.Ht + TC_TH - 1) / TC_TH;
    if (p.ncls == 1) {
      p.cls_tap0[0] = 0; p.cls_tap0[1] = (unsigned char)p.ntaps;
      for (int a = 0; a < 3; ++a) p.cls_off[0][a] = (signed char)p.out_off[a];
    }
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

  if (mode == 0 && g_use_halo && !(Cin == 32 && Cout == 64)) {
    if (!make_act_map(&maps.a[0], x, Cin, Wi, Hi, Di, P * B, sW, sH, sD, sB, HB_W, HB_H)) return DCA_ERR_LAUNCH;
    for (int i = 1; i < 8; ++i) maps.a[i] = maps.a[0];
    p.Dt = Do; p.Ht = Ho; p.Wt = Wo; p.out_stride = 1; p.ntaps = 27; p.ncls = 1;
    p.cls_tap0[0] = 0; p.cls_tap0[1] = 27;
    p.tiles_w = (p.Wt + TC_TW - 1) / TC_TW;
    p.tiles_h = (p.Ht + TC_TH - 1) / TC_TH;
#define DCA_TCH_CASE(CI, CO)                                                       \
  if (Cin == CI && Cout == CO)                                                     \
    return P == 2 ? launch_tc_halo<CI, CO, 2>(maps, p, st) : launch_tc_halo<CI, CO, 1>(maps, p, st);
    DCA_TCH_CASE(32, 32)
    DCA_TCH_CASE(64, 32)
    DCA_TCH_CASE(64, 64)
#undef DCA_TCH_CASE
    return DCA_ERR_UNSUPPORTED;
  }
  const long long nvox = (long long)B * Di * Hi * Wi;
  if (mode == 3 && (nvox % 8) == 0 && nvox < (1ll << 31)) {
    // 1x1x1 conv: the voxel order is irrelevant, so tile the flat voxel index (8 x nvox/8 view): each tile is 128
    // consecutive voxels = one contiguous 8/16 KB burst per plane instead of sixteen 512-byte rows
    const int rows = (int)(nvox / 8);
    if (!make_act_map(&maps.a[0], x, Cin, 8, rows, 1, P, (size_t)Cin, (size_t)8 * Cin, (size_t)nvox * Cin,
                      (size_t)nvox * Cin))
      return DCA_ERR_LAUNCH;
    for (int i = 1; i < 8; ++i) maps.a[i] = maps.a[0];
    p.linear = 1; p.rB = B; p.rD = Do; p.rH = Ho; p.rW = Wo;
    p.B = 1; p.Do = 1; p.Ho = rows; p.Wo = 8;
    p.Dt = 1; p.Ht = rows; p.Wt = 8; p.out_stride = 1;
    p.ntaps = 1; p.tap_map[0] = 0; p.tap_w[0] = 0;
    return run();
  }
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
  if (mode == 1 && g_use_s2slab && Cin == 32 && Cout == 64 && (Wi % 2) == 0) {
    // halo slabs over the (w-pair x 64 channel) view of the input: dims (64, W/2, H, D, planes*B)
    if (!make_act_map(&maps.a[0], x, 64, Wi / 2, Hi, Di, P * B, (size_t)64, (size_t)Wi * Cin, sD, sB, TC_TW + 1, 2 * TC_TH + 1))
      return DCA_ERR_LAUNCH;
    for (int i = 1; i < 9; ++i) maps.a[i] = maps.a[0];
    p.Dt = Do; p.Ht = Ho; p.Wt = Wo; p.out_stride = 1; p.ntaps = 27; p.ncls = 1;
    p.cls_tap0[0] = 0; p.cls_tap0[1] = 27;
    p.tiles_w = (p.Wt + TC_TW - 1) / TC_TW;
    p.tiles_h = (p.Ht + TC_TH - 1) / TC_TH;
    return P == 2 ? launch_tc_s2slab<2>(maps, p, st) : launch_tc_s2slab<1>(maps, p, st);
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
  // mode 2: transposed conv (o = 2i - 1 + k): the 8 output parity classes partition the 27 taps; one launch
  if (!make_act_map(&maps.a[0], x, Cin, Wi, Hi, Di, P * B, sW, sH, sD, sB)) return DCA_ERR_LAUNCH;
  for (int i = 1; i < 9; ++i) maps.a[i] = maps.a[0];
  if (side) {   // parity views of the side input at the OUTPUT resolution: class c reads side[2t + parity(c)]
    const size_t qW = side_c, qH = (size_t)Wo * side_c, qD = (size_t)Ho * Wo * side_c, qB = (size_t)Do * Ho * Wo * side_c;
    for (int pc = 0; pc < 8; ++pc) {
      const int pz = (pc >> 2) & 1, py = (pc >> 1) & 1, px = pc & 1;
      const __nv_bfloat16* base = (const __nv_bfloat16*)side + pz * qD + py * qH + px * qW;
      if (!make_act_map(&maps.a[1 + pc], base, side_c, Wi, Hi, Di, P * B, 2 * qW, 2 * qH, 2 * qD, qB, TC_TW, TC_TH, Cin))
        return DCA_ERR_LAUNCH;
    }
  }
  p.Dt = Di; p.Ht = Hi; p.Wt = Wi; p.out_stride = 2;
  p.ncls = 8; p.ntaps = 0;
  for (int pc = 0; pc < 8; ++pc) {
    const int par[3] = {(pc >> 2) & 1, (pc >> 1) & 1, pc & 1};
    p.cls_tap0[pc] = (unsigned char)p.ntaps;
    for (int a = 0; a < 3; ++a) p.cls_off[pc][a] = (signed char)par[a];
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
      const int k[3] = {kd, kh, kw};
      int off[3]; bool ok = true;
      for (int a = 0; a < 3; ++a) { const int n = par[a] + 1 - k[a]; if (n & 1) ok = false; off[a] = n / 2; }
      if (!ok) continue;
      const int t = p.ntaps++;
      p.tap_off[t][0] = (signed char)off[0]; p.tap_off[t][1] = (signed char)off[1]; p.tap_off[t][2] = (signed char)off[2];
      p.tap_map[t] = 0; p.tap_w[t] = (signed char)((kd * 3 + kh) * 3 + kw);
    }
    if (side) {
      const int t = p.ntaps++;
      p.tap_off[t][0] = p.tap_off[t][1] = p.tap_off[t][2] = 0;
      p.tap_map[t] = (signed char)(1 + pc); p.tap_w[t] = 27;
    }
  }
  p.cls_tap0[8] = (unsigned char)p.ntaps;
  return run();
}

// channel tile of the 2-D family: 64 (Cin a multiple of 64, up to 320) or 32 (Cin == 32: the 1/2-resolution stem)
static inline int tc2d_tile(int Ci) { return (Ci > 0 && Ci % 64 == 0 && Ci <= 320) ? 64 : (Ci == 32 ? 32 : 0); }

extern "C" long long dca_pack_weights_tc2d_bytes(int Co, int Ci, int planes) {
  const int T = tc2d_tile(Ci);
  if (Co <= 0 || !T || planes < 1 || planes > 2) return 0;
  return (long long)((Co + T - 1) / T) * (Ci / T) * 9 * planes * T * T * 2;
}

extern "C" int dca_pack_weights_tc2d(const float* w, int Co, int Ci, void* out, int planes, void* stream) {
  if (!w || !out || dca_pack_weights_tc2d_bytes(Co, Ci, planes) == 0) return DCA_ERR_ARG;
  const int T = tc2d_tile(Ci);
  const int nchunk = (Co + T - 1) / T, nslab = Ci / T, total = nchunk * nslab * 9 * T * T;
  pack_weight_tc2d_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Co, Ci, (__nv_bfloat16*)out, planes,
                                                                                 nchunk, nslab, T);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

// y = act_post(act(scale * conv2d_3x3(x, dilation dil) + shift) + res):
//   x cost planes [P][B][1][H][W][Cin], Cin a multiple of 64 up to 320 (one 64-channel slab per accumulation pass);
//   y = cost planes [P][B][1][H][W][Cout] (Cout % 64 == 0) or, with out_f32, fp32 channels-last [B][H][W][Cout];
//   res (optional) = cost planes shaped like y, added after `act`; act_post is applied after that add.
// Runs the halo-slab tcgen05 kernel once per 64-channel output chunk.  dil == 2 (feature_extraction.layer4,
// gwcnet_dca_g.py:24): a dilation-2 3x3 conv is four independent dilation-1 convs on the (row parity, column parity)
// sub-images, so the same kernel runs on four strided tensor-map views (H and W even) and the epilogue scatters with
// out_stride 2 -- no gather, no second kernel.
template <int T>
static int conv2d_tc_run_t(const void* const* xs, const int* cs, int nsrc, int planes, const void* w_tc2d,
                           const float* scale, const float* shift, const void* res, int act_post, void* y, int out_f32,
                           int act, int B, int Cin, int Cout, int H, int W, int dil, cudaStream_t st) {
  const int P = planes, nslab = Cin / T, nchunk = (Cout + T - 1) / T;
  const int Hs = H / dil, Ws = W / dil;                       // sub-image extent (dil == 1: the image itself)
  // dilation 2 with ONE source: the four parity sub-images ride the tile-space depth index of a single launch (par2d);
  // with several sources (never needed so far) they are four launches
  const bool batched = dil == 2 && nsrc == 1;
  const int npass = batched ? 1 : dil * dil;
  for (int pass = 0; pass < npass; ++pass) {
    TcMaps maps;
    const int npar = batched ? 4 : 1;
    for (int q = 0; q < npar; ++q) {
      const int par = batched ? q : pass;
      const int pa = par / dil, pb = par % dil;               // row / column parity of this sub-image
      for (int i = 0; i < nsrc; ++i) {
        const int C = cs[i];
        const __nv_bfloat16* xb = (const __nv_bfloat16*)xs[i] + ((size_t)pa * W + pb) * C;
        if (!make_act_map(&maps.a[batched ? q : i], xb, C, Ws, Hs, 1, P * B, (size_t)dil * C, (size_t)dil * W * C,
                          (size_t)H * W * C, (size_t)H * W * C, HB_W, HB_H, T))
          return DCA_ERR_LAUNCH;
      }
    }
    for (int i = batched ? 4 : nsrc; i < 9; ++i) maps.a[i] = maps.a[0];
    const int pa = pass / dil, pb = pass % dil;
    for (int j = 0; j < nchunk; ++j) {
      TcParams p;
      memset(&p, 0, sizeof(p));
      p.B = B; p.Do = 1; p.Ho = H; p.Wo = W;
      p.scale = scale ? scale + j * T : nullptr; p.shift = shift ? shift + j * T : nullptr;
      p.y = (__nv_bfloat16*)y; p.y_plane = (size_t)B * H * W * Cout; p.res_plane = p.y_plane; p.planes_res = res ? P : 1;
      p.res_post = (const __nv_bfloat16*)res; p.act_post = res ? act_post : 0;
      p.planes_out = P; p.act = act; p.dbg = g_dbg;
      p.ldc = Cout; p.co_base = j * T; p.cout_valid = Cout; p.out_f32 = out_f32;
      p.nslab = nslab;
      for (int i = 0, sl = 0; i < nsrc; ++i)
        for (int c0 = 0; c0 < cs[i]; c0 += T, ++sl) { p.slab_map[sl] = batched ? 0 : i; p.slab_c0[sl] = c0; p.slab_dz[sl] = 0; }
      p.par2d = batched ? 1 : 0;
      p.Dt = batched ? 4 : 1; p.Ht = Hs; p.Wt = Ws; p.out_stride = dil; p.out_stride_d = 1; p.ntaps = 9 * nslab; p.ncls = 1;
      p.cls_off[0][0] = 0;
      p.cls_off[0][1] = (signed char)(batched ? 0 : pa); p.cls_off[0][2] = (signed char)(batched ? 0 : pb);
      p.cls_tap0[0] = 0; p.cls_tap0[1] = (unsigned char)(9 * nslab);
      p.tiles_w = (Ws + TC_TW - 1) / TC_TW; p.tiles_h = (Hs + TC_TH - 1) / TC_TH;
      const __nv_bfloat16* wj = (const __nv_bfloat16*)w_tc2d + (size_t)j * nslab * 9 * P * T * T;
      if (!make_w_map(&maps.w, wj, T, nslab * 9 * P * T, P * T)) return DCA_ERR_LAUNCH;
      const int rc = P == 2 ? launch_tc_halo<T, T, 2>(maps, p, st) : launch_tc_halo<T, T, 1>(maps, p, st);
      if (rc != DCA_OK) return rc;
    }
  }
  return DCA_OK;
}

static int conv2d_tc_run(const void* const* xs, const int* cs, int nsrc, int planes, const void* w_tc2d,
                         const float* scale, const float* shift, const void* res, int act_post, void* y, int out_f32,
                         int act, int B, int Cout, int H, int W, int dil, cudaStream_t st) {
  if (!w_tc2d || !y || B <= 0 || planes < 1 || planes > 2 || H <= 0 || W <= 0 || nsrc < 1 || nsrc > 3) return DCA_ERR_ARG;
  int Cin = 0;
  for (int i = 0; i < nsrc; ++i) {
    if (!xs[i]) return DCA_ERR_ARG;
    Cin += cs[i];
  }
  const int T = tc2d_tile(Cin);
  if (!T) return DCA_ERR_UNSUPPORTED;
  for (int i = 0; i < nsrc; ++i)
    if (cs[i] <= 0 || (cs[i] % T) != 0) return DCA_ERR_UNSUPPORTED;
  if (Cout <= 0 || (!out_f32 && (Cout % T) != 0)) return DCA_ERR_UNSUPPORTED;
  if (dil != 1 && !(dil == 2 && (H % 2) == 0 && (W % 2) == 0)) return DCA_ERR_UNSUPPORTED;
  if (res && out_f32) return DCA_ERR_UNSUPPORTED;
  return T == 64 ? conv2d_tc_run_t<64>(xs, cs, nsrc, planes, w_tc2d, scale, shift, res, act_post, y, out_f32, act, B, Cin,
                                       Cout, H, W, dil, st)
                 : conv2d_tc_run_t<32>(xs, cs, nsrc, planes, w_tc2d, scale, shift, res, act_post, y, out_f32, act, B, Cin,
                                       Cout, H, W, dil, st);
}

extern "C" int dca_conv2d_tc_ex(const void* x, int planes, const void* w_tc2d, const float* scale, const float* shift,
                                const void* res, int act_post, void* y, int out_f32, int act, int B, int Cin, int Cout,
                                int H, int W, int dil, void* stream) {
  const void* xs[1] = {x};
  const int cs[1] = {Cin};
  return conv2d_tc_run(xs, cs, 1, planes, w_tc2d, scale, shift, res, act_post, y, out_f32, act, B, Cout, H, W, dil,
                       (cudaStream_t)stream);
}

// The same conv over the CHANNEL CONCATENATION of up to three plane tensors (each a multiple of 64 channels, 320 in
// total at most) without materialising it: every 64-channel slab of the accumulation is read through its own tensor map.
// feature_extraction.lastconv reads cat(l2, l3, l4) (gwcnet_dca_g.py:60-65) this way.
extern "C" int dca_conv2d_tc_cat(const void* x0, int C0, const void* x1, int C1, const void* x2, int C2, int planes,
                                 const void* w_tc2d, const float* scale, const float* shift, void* y, int out_f32, int act,
                                 int B, int Cout, int H, int W, void* stream) {
  const void* xs[3] = {x0, x1, x2};
  const int cs[3] = {C0, C1, C2};
  const int nsrc = x2 ? 3 : (x1 ? 2 : 1);
  return conv2d_tc_run(xs, cs, nsrc, planes, w_tc2d, scale, shift, nullptr, 0, y, out_f32, act, B, Cout, H, W, 1,
                       (cudaStream_t)stream);
}

// y = act(scale * conv2d_3x3(x) + shift) (PropgationNet_4x.conv): dca_conv2d_tc_ex without residual and dilation
extern "C" int dca_conv2d_tc(const void* x, int planes, const void* w_tc2d, const float* scale, const float* shift,
                             void* y, int out_f32, int act, int B, int Cin, int Cout, int H, int W, void* stream) {
  return dca_conv2d_tc_ex(x, planes, w_tc2d, scale, shift, nullptr, 0, y, out_f32, act, B, Cin, Cout, H, W, 1, stream);
}

// 1x1x1 conv 32 -> ntap (<= 32) "per-tap partial products" of a 3x3x3 conv to ONE channel (cva.classify.2 cva.py:53,
// classif3.2 gwcnet_dca_g.py:168):  P[tap][v] = sum_c x[v][c] * w[tap][c], fp32, tap-major [ntap][B*D*H*W].
// dca_tap_gather3d then sums P over the 27 shifted taps.  w_tc = dca_pack_weights_tc of a [32][32][1] weight whose
// rows >= ntap are zero.
extern "C" int dca_conv1_taps_tc(const void* x, int planes, const void* w_tc, float* P, int ntap, int B, int D, int H,
                                 int W, void* stream) {
  if (!x || !w_tc || !P || planes < 1 || planes > 2 || ntap <= 0 || ntap > 32 || B <= 0) return DCA_ERR_ARG;
  const long long nvox = (long long)B * D * H * W;
  if ((nvox % 8) != 0 || nvox / 8 >= (1ll << 31)) return DCA_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int Pn = planes, Cin = 32, Cout = 32;
  TcMaps maps;
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int rows = (int)(nvox / 8);
  if (!make_act_map(&maps.a[0], x, Cin, 8, rows, 1, Pn, (size_t)Cin, (size_t)8 * Cin, (size_t)nvox * Cin, (size_t)nvox * Cin))
    return DCA_ERR_LAUNCH;
  for (int i = 1; i < 9; ++i) maps.a[i] = maps.a[0];
  if (!make_w_map(&maps.w, w_tc, Cin, Pn * Cout, Pn * Cout)) return DCA_ERR_LAUNCH;
  p.B = 1; p.Do = 1; p.Ho = rows; p.Wo = 8;
  p.y = (__nv_bfloat16*)P; p.planes_out = Pn; p.act = 0; p.dbg = g_dbg;
  p.ldc = Cout; p.cout_valid = ntap; p.out_f32 = 2;
  p.nslab = 3;
  p.Dt = 1; p.Ht = rows; p.Wt = 8; p.out_stride = 1; p.ncls = 1;
  p.ntaps = 1; p.tap_map[0] = 0; p.tap_w[0] = 0;
  p.cls_tap0[0] = 0; p.cls_tap0[1] = 1;
  p.tiles_w = 1; p.tiles_h = (rows + TC_TH - 1) / TC_TH;
  return Pn == 2 ? launch_tc<32, 32, 2>(maps, p, st) : launch_tc<32, 32, 1>(maps, p, st);
}

extern "C" long long dca_pack_weights_tc_march_bytes(int Ci, int planes) {
  if ((Ci != 32 && Ci != 64) || planes < 1 || planes > 2) return 0;
  return (long long)27 * planes * 32 * Ci * 2;
}

extern "C" int dca_pack_weights_tc_march(const float* w, int Ci, void* out, int planes, void* stream) {
  if (!w || !out || dca_pack_weights_tc_march_bytes(Ci, planes) == 0) return DCA_ERR_ARG;
  const int total = 27 * 32 * Ci;
  pack_weight_tc_march_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Ci, (__nv_bfloat16*)out, planes);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

// Conv3d k3 s1 p1, Cout = 32, depth-marching kernel (kd folded into N).  Same epilogue contract as dca_conv3d_tc:
//   y = act(scale * conv(x, w) + shift + res_pre) + res_post ;  w_march from dca_pack_weights_tc_march.
// Planes per work item of the depth-marching kernel.  An item of n planes costs n + 2 slab steps (one per input plane
// of its halo'd depth range) and the persistent grid runs ceil(items / SMs) items per CTA, so n trades the halo overhead
// against the fill of the last wave: KITTI 1/4 res (234 tile columns x 48 planes): n = 8 (1404 items, 10 x 10 steps);
// 1/8 res (60 columns x 24 planes): n = 6 (240 items, 2 x 8 = 16 steps) instead of n = 8 (180 items, 2 x 10 = 20).
static int g_march_n = 0;          // > 0: forced (timing experiments)
extern "C" int dca_tc_set_march_n(int n) { g_march_n = n > 0 ? n : 0; return DCA_OK; }
static int choose_march_n(int B, int D, int H, int W, int nmax) {
  if (g_march_n > 0) return g_march_n < nmax ? (g_march_n < D ? g_march_n : D) : (nmax < D ? nmax : D);
  const long long cols = (long long)B * ((H + TC_TH - 1) / TC_TH) * ((W + TC_TW - 1) / TC_TW);
  const int sms = dca_num_sms();
  int best = D < nmax ? D : nmax;
  long long best_cost = -1;
  for (int n = (D < nmax ? D : nmax); n >= 2; --n) {
    const long long items = cols * ((D + n - 1) / n);
    const long long cost = ((items + sms - 1) / sms) * (n + 2);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = n; }
  }
  return best;
}

extern "C" int dca_conv3d_tc_march(const void* x, int planes, const void* w_march, const float* scale, const float* shift,
                                   const void* res_pre, const void* res_post, int planes_res, void* y, int act, int B,
                                   int Cin, int D, int H, int W, void* stream) {
  if (!x || !w_march || !y || B <= 0 || planes < 1 || planes > 2 || D <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  if (Cin != 32 && Cin != 64) return DCA_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = planes, Cout = 32;
  TcMaps maps;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.Do = D; p.Ho = H; p.Wo = W;
  p.scale = scale; p.shift = shift;
  p.res_pre = (const __nv_bfloat16*)res_pre; p.res_post = (const __nv_bfloat16*)res_post;
  p.res_plane = (size_t)B * D * H * W * Cout; p.planes_res = planes_res;
  p.y = (__nv_bfloat16*)y; p.y_plane = p.res_plane; p.planes_out = P; p.act = act;
  p.dbg = g_dbg;
  p.ldc = Cout; p.cout_valid = Cout;
  const int nmax = P == 2 ? 8 : 16;            // planes per work item (TMEM: 512 columns)
  const int n = choose_march_n(B, D, H, W, nmax);
  p.march_n = n;
  p.Dt = (D + n - 1) / n;                      // depth chunks
  p.Ht = H; p.Wt = W; p.out_stride = 1; p.ncls = 1;
  p.tiles_w = (W + TC_TW - 1) / TC_TW; p.tiles_h = (H + TC_TH - 1) / TC_TH;
  if (!make_act_map(&maps.a[0], x, Cin, W, H, D, P * B, (size_t)Cin, (size_t)W * Cin, (size_t)H * W * Cin,
                    (size_t)D * H * W * Cin, HB_W, HB_H))
    return DCA_ERR_LAUNCH;
  for (int i = 1; i < 9; ++i) maps.a[i] = maps.a[0];
  if (!make_w_map(&maps.w, w_march, Cin, 9 * P * 3 * Cout, P * 3 * Cout)) return DCA_ERR_LAUNCH;
  if (Cin == 32) return P == 2 ? launch_tc_march<32, 2>(maps, p, st) : launch_tc_march<32, 1>(maps, p, st);
  return P == 2 ? launch_tc_march<64, 2>(maps, p, st) : launch_tc_march<64, 1>(maps, p, st);
}

// Conv3d k3 s1 p1 (Cin in {32,64} -> 32) + BN + activation, immediately followed by Conv3d(32 -> 1, k3, p1, no bias):
//   P[tap][v] = sum_c w27[tap][c] * act(scale * conv(x, w)[v][c] + shift)      fp32, tap-major [27][B*D*H*W]
// (classif3 gwcnet_dca_g.py:166-168, cva.classify cva.py:51-53).  The 32-channel intermediate is never stored.
// use_march = 1: depth-marching kernel (w = dca_pack_weights_tc_march), else the halo-slab kernel (w = dca_pack_weights_tc).
// w27_host: HOST pointer to [27][32] fp32 (becomes launch parameters).  dca_tap_gather_* turn P into logits / disparity.
extern "C" int dca_conv3d_tc_taps27(const void* x, int planes, const void* w, int use_march, const float* scale,
                                    const float* shift, const float* w27_host, float* P, int act, int B, int Cin, int D,
                                    int H, int W, void* stream) {
  if (!x || !w || !w27_host || !P || B <= 0 || planes < 1 || planes > 2 || D <= 0 || H <= 0 || W <= 0) return DCA_ERR_ARG;
  if (Cin != 32 && Cin != 64) return DCA_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int Pn = planes, Cout = 32;
  TcMaps maps;
  TcParams p;
  TapW tw;
  memcpy(tw.w, w27_host, sizeof(tw.w));
  memset(&p, 0, sizeof(p));
  p.B = B; p.Do = D; p.Ho = H; p.Wo = W;
  p.scale = scale; p.shift = shift;
  p.y = (__nv_bfloat16*)P; p.planes_out = Pn; p.act = act; p.dbg = g_dbg;
  p.ldc = Cout; p.cout_valid = Cout;
  p.Ht = H; p.Wt = W; p.out_stride = 1; p.ncls = 1;
  p.tiles_w = (W + TC_TW - 1) / TC_TW; p.tiles_h = (H + TC_TH - 1) / TC_TH;
  if (!make_act_map(&maps.a[0], x, Cin, W, H, D, Pn * B, (size_t)Cin, (size_t)W * Cin, (size_t)H * W * Cin,
                    (size_t)D * H * W * Cin, HB_W, HB_H))
    return DCA_ERR_LAUNCH;
  for (int i = 1; i < 9; ++i) maps.a[i] = maps.a[0];
  if (use_march) {
    const int nmax = Pn == 2 ? 8 : 16;
    const int n = choose_march_n(B, D, H, W, nmax);
    p.march_n = n;
    p.Dt = (D + n - 1) / n;
    if (!make_w_map(&maps.w, w, Cin, 9 * Pn * 3 * Cout, Pn * 3 * Cout)) return DCA_ERR_LAUNCH;
    if (Cin == 32) return Pn == 2 ? launch_tc_march<32, 2, 1>(maps, p, st, &tw) : launch_tc_march<32, 1, 1>(maps, p, st, &tw);
    return Pn == 2 ? launch_tc_march<64, 2, 1>(maps, p, st, &tw) : launch_tc_march<64, 1, 1>(maps, p, st, &tw);
  }
  p.nslab = 3; p.slab_dz[0] = -1; p.slab_dz[1] = 0; p.slab_dz[2] = 1;
  p.Dt = D; p.ntaps = 27; p.cls_tap0[0] = 0; p.cls_tap0[1] = 27;
  if (!make_w_map(&maps.w, w, Cin, 27 * Pn * Cout, Pn * Cout)) return DCA_ERR_LAUNCH;
  if (Cin == 32) return Pn == 2 ? launch_tc_halo<32, 32, 2, 1>(maps, p, st, &tw) : launch_tc_halo<32, 32, 1, 1>(maps, p, st, &tw);
  return Pn == 2 ? launch_tc_halo<64, 32, 2, 1>(maps, p, st, &tw) : launch_tc_halo<64, 32, 1, 1>(maps, p, st, &tw);
}

// AvgPool3d k3 s2 p1 (count_include_pad) of cost planes [planes][B][Di][Hi][Wi][C] -> [planes][B][ceil/2 ...][C].
// C == 32: TMA-staged depth-marching kernel (every input plane read once); otherwise the thread-per-output kernel.
extern "C" int dca_avgpool3d(const void* x, void* y, int planes, int B, int C, int Di, int Hi, int Wi, void* stream) {
  if (!x || !y || C % 8 != 0 || B <= 0 || Di <= 0 || Hi <= 0 || Wi <= 0 || planes < 1 || planes > 2) return DCA_ERR_ARG;
  if (C != 32 || !g_use_pool_march) return dca_avgpool3d_simple(x, y, planes, B, C, Di, Hi, Wi, stream);
  const int Do = (Di + 1) / 2, Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  CUtensorMap xmap;
  {
    EncodeTiledFn enc = get_encode();
    if (!enc) return DCA_ERR_LAUNCH;
    cuuint64_t dims[5] = {32, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)planes * B};
    cuuint64_t strides[4] = {64, (cuuint64_t)Wi * 64, (cuuint64_t)Hi * Wi * 64, (cuuint64_t)Di * Hi * Wi * 64};
    cuuint32_t box[5] = {32, AP_BW, AP_BH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (enc(&xmap, (DCA_F16_PLANES ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 5,
            const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return dca_avgpool3d_simple(x, y, planes, B, C, Di, Hi, Wi, stream);     // (tensors smaller than one box)
  }
  g_num_sms = dca_num_sms();
  const int tiles_w = (Wo + AP_TW - 1) / AP_TW, tiles_h = (Ho + AP_TH - 1) / AP_TH;
  const long long cols = (long long)B * tiles_w * tiles_h;
  int nsplit = (int)(((long long)g_pool_ctas_per_sm * g_num_sms + cols - 1) / cols);   // CTAs per SM so that loads and sums overlap
  if (nsplit < 1) nsplit = 1;
  if (nsplit > Do) nsplit = Do;
  const long long grid = cols * nsplit;
  if (grid > 0x7fffffffLL) return DCA_ERR_UNSUPPORTED;
  const int smem = 2 * planes * AP_PITCH + 128;
  cudaStream_t st = (cudaStream_t)stream;
  if (planes == 2) {
    cudaFuncSetAttribute(avgpool3d_march_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dca_launch(avgpool3d_march_kernel<2>, (int)grid, AP_TW * AP_TH * 4, smem, st, xmap, (__nv_bfloat16*)y, B, Do, Ho, Wo,
               tiles_w, tiles_h, nsplit);
  } else {
    cudaFuncSetAttribute(avgpool3d_march_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dca_launch(avgpool3d_march_kernel<1>, (int)grid, AP_TW * AP_TH * 4, smem, st, xmap, (__nv_bfloat16*)y, B, Do, Ho, Wo,
               tiles_w, tiles_h, nsplit);
  }
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

// 1 (default): dca_avgpool3d uses the depth-marching TMA kernel for C == 32; 0: always the thread-per-output kernel
// on > 1 additionally sets the number of CTAs per SM the depth range is split for (default 4)
extern "C" int dca_pool_set_march(int on) { g_use_pool_march = on ? 1 : 0; if (on > 1) g_pool_ctas_per_sm = on; return DCA_OK; }
