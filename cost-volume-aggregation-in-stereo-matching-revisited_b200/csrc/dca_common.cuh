// Shared device helpers for the DCANet hot-path kernels (sm_100a).
//
// Cost tensors ("cost planes") live in HBM as channels-last bf16:  [plane][B][D][H][W][C]
//   plane 0 = hi = bf16(x),  plane 1 = lo = bf16(x - hi)   (parity mode, ~16 bit significand)
//   planes == 1 keeps only hi                               (fast mode)
// C is a multiple of 8, so one voxel-chunk of 8 channels is one 16-byte vector per plane.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

#define DCA_OK 0
#define DCA_ERR_ARG (-1)
#define DCA_ERR_LAUNCH (-2)
#define DCA_ERR_UNSUPPORTED (-3)

#define DCA_RETURN_IF_LAUNCH_FAILED()                      \
  do {                                                     \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return DCA_ERR_LAUNCH;         \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// Every kernel of the forward is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it may start while the
// previous kernel of the stream is still draining, runs its prologue (mbarrier init, TMEM allocation, tensor-map
// prefetch, loads of STATIC data such as resident weights) and then executes griddepcontrol.wait, which returns once the
// previous grid has completed and its memory is visible.  Persistent single-wave kernels call
// griddepcontrol.launch_dependents at their very start, so the successor's CTAs take over SMs as this kernel's CTAs
// retire.  A kernel must not read data produced by, nor write data read by, an earlier kernel before pdl_wait().
// dca_set_pdl(0) launches everything fully serialised (same results).
extern "C" int dca_pdl_enabled(void);
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// SM count of the CURRENT device (cached per device ordinal; grids of the persistent kernels are sized from it)
static inline int dca_num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

template <typename... Exp, typename... Act>
static inline cudaError_t dca_launch(void (*kern)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Act&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = dca_pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Exp>(args)...);
}

namespace dca {

// ---- 16-bit plane format --------------------------------------------------------------------------------
// DCA_F16_PLANES = 1 (default): planes are IEEE fp16.  hi = fp16(x) carries 11 significand bits and lo = fp16(x - hi)
// another 11, so hi + lo reproduces x to ~2^-23 (fp32 level) for |x| in [6e-5, 65504] -- the range of every
// BN-normalised activation and weight of this network (measured: |x| <= 30).  Same tensor-pipe
// rate as bf16 (kind::f16).  Conversions saturate to +-65504 instead of producing inf.
// DCA_F16_PLANES = 0: bf16 planes (8 + 8 bits, ~2^-17): 8 more exponent bits for un-normalised checkpoints, at 60x the
// representation error; KITTI-size parity then exceeds 0.05 px at a few pixels per image (measured 0.066 px).
// The storage type in signatures stays `__nv_bfloat16` = "16-bit plane element"; only these helpers interpret the bits.
#ifndef DCA_F16_PLANES
#define DCA_F16_PLANES 1
#endif

__device__ __forceinline__ float2 h16x2_to_f2(uint32_t w) {            // packed pair -> (low half, high half) as fp32
#if DCA_F16_PLANES
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
#else
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
#endif
}
__device__ __forceinline__ uint32_t f2_to_h16x2(float a, float b) {    // a -> low half, b -> high half, round to nearest
#if DCA_F16_PLANES
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
#else
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
#endif
}
__device__ __forceinline__ float h16_bits_to_float(uint32_t bits16) {
#if DCA_F16_PLANES
  return __half2float(__ushort_as_half((unsigned short)bits16));
#else
  return __uint_as_float(bits16 << 16);
#endif
}
__device__ __forceinline__ uint32_t f2h16_bits(float x) {              // round-to-nearest-even, as torch does
  return f2_to_h16x2(x, 0.f) & 0xffffu;
}
// (a, b) -> packed hi pair and packed lo pair (= the 16-bit roundings of the remainders)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = f2_to_h16x2(a, b);
  const float2 hf = h16x2_to_f2(hi);
  lo = f2_to_h16x2(a - hf.x, b - hf.y);
}

// unpack 8 plane elements (one uint4) into 8 floats
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const float2 a = h16x2_to_f2(v.x), b = h16x2_to_f2(v.y), c = h16x2_to_f2(v.z), d = h16x2_to_f2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// split x into hi and lo (the 16-bit rounding of the remainder); returns hi bits, writes lo bits
__device__ __forceinline__ uint32_t split_bf16(float x, uint32_t& lo_bits) {
  const uint32_t h = f2h16_bits(x);
  lo_bits = f2h16_bits(x - h16_bits_to_float(h));
  return h;
}

// load 8 consecutive channels of one voxel (element offset `off`, multiple of 8) as fp32
template <int PLANES>
__device__ __forceinline__ void load8(const __nv_bfloat16* __restrict__ base, size_t plane_stride, size_t off,
                                      float* f) {
  uint4 h = *reinterpret_cast<const uint4*>(base + off);
  unpack8(h, f);
  if (PLANES == 2) {
    uint4 l = *reinterpret_cast<const uint4*>(base + plane_stride + off);
    float g[8];
    unpack8(l, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += g[i];
  }
}

__device__ __forceinline__ void load8_rt(const __nv_bfloat16* __restrict__ base, size_t plane_stride, int planes,
                                         size_t off, float* f) {
  if (planes == 2) load8<2>(base, plane_stride, off, f);
  else load8<1>(base, plane_stride, off, f);
}

// store 8 consecutive channels
template <int PLANES>
__device__ __forceinline__ void store8(__nv_bfloat16* __restrict__ base, size_t plane_stride, size_t off,
                                       const float* f) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(f[2 * i], f[2 * i + 1], h[i], l[i]);
  *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (PLANES == 2) *reinterpret_cast<uint4*>(base + plane_stride + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void store8_rt(__nv_bfloat16* __restrict__ base, size_t plane_stride, int planes,
                                          size_t off, const float* f) {
  if (planes == 2) store8<2>(base, plane_stride, off, f);
  else store8<1>(base, plane_stride, off, f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : 0.1f * v;
  return v;
}

}  // namespace dca
