// Shared device helpers for the DCANet hot-path kernels (sm_100a).
//
// Cost tensors ("cost planes") live in HBM as channels-last bf16:  [plane][B][D][H][W][C]
//   plane 0 = hi = bf16(x),  plane 1 = lo = bf16(x - hi)   (parity mode, ~16 bit significand)
//   planes == 1 keeps only hi                               (fast mode)
// C is a multiple of 8, so one voxel-chunk of 8 channels is one 16-byte vector per plane.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DCA_OK 0
#define DCA_ERR_ARG (-1)
#define DCA_ERR_LAUNCH (-2)
#define DCA_ERR_UNSUPPORTED (-3)

#define DCA_RETURN_IF_LAUNCH_FAILED()                      \
  do {                                                     \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return DCA_ERR_LAUNCH;         \
  } while (0)

namespace dca {

__device__ __forceinline__ float bf16_bits_to_float(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

// unpack 8 bf16 (one uint4) into 8 floats
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

__device__ __forceinline__ uint32_t f2bf_bits(float x) {  // round-to-nearest-even, as torch does
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
}

// split x into hi (bf16) and lo (bf16 of the remainder); returns hi bits, writes lo bits
__device__ __forceinline__ uint32_t split_bf16(float x, uint32_t& lo_bits) {
  uint32_t h = f2bf_bits(x);
  float r = x - __uint_as_float(h << 16);
  lo_bits = f2bf_bits(r);
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
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = split_bf16(f[i], l[i]);
  uint4 hv = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
  *reinterpret_cast<uint4*>(base + off) = hv;
  if (PLANES == 2) {
    uint4 lv = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
    *reinterpret_cast<uint4*>(base + plane_stride + off) = lv;
  }
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
