// Halo-row exchange of the H-sharded single-pair mode over NVLink peer memory (no NCCL, no host round trip).
//
// A tensor is seen as [outer][rows][inner bytes]; a rank owns rows [h, rows-h) and keeps h halo rows at both ends, of
// which only the `live` rows next to the owned ones are ever read for an owned result (hshard.py) and travel.
// One exchange = two launches on the rank's compute stream:
//
//   halo_push_kernel         reads the first / last `live` OWNED rows and stores them, packed, straight into the UPPER /
//                            LOWER neighbour's staging slot (peer pointers into the neighbours' symmetric buffers: the
//                            stores travel over NVLink), then releases them: __threadfence_system + one system-scope
//                            atomic increment per CTA of the neighbour's arrival counter.
//   halo_wait_unpack_kernel  acquires the own arrival counters (ld.acquire.sys, bounded spin), then copies the staged
//                            rows into the halo rows of the local tensor.
//
// Both neighbours push before they wait (stream order), so there is no cycle.  The staging slots are double buffered
// by exchange parity on the host side: a neighbour can only overwrite slot p after it has seen this rank's NEXT push,
// which this rank issues after its unpack of slot p (same stream).  The arrival counters only ever grow
// (exchange number x PUSH_CTAS), so they need no reset.  A wait that exceeds the time-out (default 10 s,
// dca_halo_set_timeout_ms) sets *err and falls through instead of hanging the device; the host checks the error word
// after every forward (hshard.PeerHalo) and raises.
//
// STATUS: green on 2 B200 (tests/test_gpu_hshard.py, round 2); bench lines under profiles/r2_hshard_*.
#include "dca_common.cuh"

namespace dca {

constexpr int PUSH_CTAS = 32;         // CTAs per direction; the arrival counter grows by this much per exchange
constexpr int HALO_THREADS = 256;

struct HaloMsg {
  const char* src;                    // nullptr: no neighbour on this side
  char* dst;
  long long outer;                    // slices
  long long src_pitch, dst_pitch;     // bytes between consecutive slices at the source / destination
  long long chunk;                    // bytes per slice = live * inner (multiple of 16)
};

__device__ __forceinline__ void halo_copy(const HaloMsg& m) {
  const long long vec_per = m.chunk >> 4;
  const long long total = m.outer * vec_per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / vec_per;
    const long long off = (i - o * vec_per) << 4;
    const uint4 v = *reinterpret_cast<const uint4*>(m.src + o * m.src_pitch + off);
    *reinterpret_cast<uint4*>(m.dst + o * m.dst_pitch + off) = v;
  }
}

__global__ void __launch_bounds__(HALO_THREADS)
halo_push_kernel(const HaloMsg up, const HaloMsg down, unsigned long long* flag_up, unsigned long long* flag_down) {
  const HaloMsg& m = blockIdx.y ? down : up;
  unsigned long long* flag = blockIdx.y ? flag_down : flag_up;
  if (m.src == nullptr) return;
  halo_copy(m);                        // dst is peer memory
  __threadfence_system();              // every thread: its stores are visible system-wide before the signal
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd_system(flag, 1ULL);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(HALO_THREADS)
halo_wait_unpack_kernel(const HaloMsg top, const HaloMsg bottom, const unsigned long long* flag_top,
                        const unsigned long long* flag_bottom, unsigned long long target, int* err,
                        unsigned long long timeout_ns) {
  const HaloMsg& m = blockIdx.y ? bottom : top;
  const unsigned long long* flag = blockIdx.y ? flag_bottom : flag_top;
  if (m.src == nullptr) return;
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < target) {
      __nanosleep(100);
      if (global_ns() - t0 > timeout_ns) { atomicExch(err, 1); break; }
    }
  }
  __syncthreads();
  halo_copy(m);                        // src is the local staging slot the neighbour stored into
}

// Both halves in ONE launch: CTA (x, y) pushes its share of the rows for direction y into the neighbour's staging slot,
// releases them, then acquires the own arrival counter of direction y and unpacks its share of the staged rows.  No CTA
// waits for another CTA of its own grid (only for the neighbours' grids, which run on other GPUs), and every push
// precedes the wait of the same CTA, so two neighbouring ranks cannot block each other.
__global__ void __launch_bounds__(HALO_THREADS)
halo_exchange_kernel(const HaloMsg push_up, const HaloMsg push_down, unsigned long long* pflag_up,
                     unsigned long long* pflag_down, const HaloMsg top, const HaloMsg bottom,
                     const unsigned long long* flag_top, const unsigned long long* flag_bottom, unsigned long long target,
                     int* err, unsigned long long timeout_ns) {
  pdl_wait();                          // the rows come from the previous kernel of the stream
  {
    const HaloMsg& m = blockIdx.y ? push_down : push_up;
    unsigned long long* flag = blockIdx.y ? pflag_down : pflag_up;
    if (m.src != nullptr) {
      halo_copy(m);
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd_system(flag, 1ULL);
    }
  }
  const HaloMsg& m = blockIdx.y ? bottom : top;
  const unsigned long long* flag = blockIdx.y ? flag_bottom : flag_top;
  if (m.src == nullptr) return;
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < target) {
      __nanosleep(100);
      if (global_ns() - t0 > timeout_ns) { atomicExch(err, 1); break; }
    }
  }
  __syncthreads();
  halo_copy(m);
}

}  // namespace dca

using namespace dca;

static unsigned long long g_halo_timeout_ns = 10000000000ULL;

extern "C" int dca_halo_push_ctas() { return PUSH_CTAS; }

// time-out of the arrival wait (default 10 s): first-call skew between ranks (lazy module loads, weight packing) can
// reach seconds; after it the error word is set and the forward is invalid
extern "C" int dca_halo_set_timeout_ms(int ms) {
  if (ms <= 0) return DCA_ERR_ARG;
  g_halo_timeout_ns = (unsigned long long)ms * 1000000ULL;
  return DCA_OK;
}

// The `live` owned rows next to each cut travel: rows [h, h+live) of `t` go to the upper neighbour's staging slot, rows
// [rows-h-live, rows-h) to the lower neighbour's.  peer_*_stage / peer_*_flag: peer-mapped device addresses
// (0 = no neighbour on that side).
extern "C" int dca_halo_push(const void* t, long long outer, long long rows, long long inner_bytes, int h, int live,
                             void* peer_up_stage, void* peer_down_stage, void* peer_up_flag, void* peer_down_flag,
                             void* stream) {
  if (!t || outer <= 0 || h <= 0 || live <= 0 || live > h || rows < 2LL * h + live || inner_bytes <= 0) return DCA_ERR_ARG;
  const long long chunk = (long long)live * inner_bytes;
  if ((chunk & 15) || ((rows * inner_bytes) & 15) || (inner_bytes & 15) || ((uintptr_t)t & 15) ||
      ((uintptr_t)peer_up_stage & 15) || ((uintptr_t)peer_down_stage & 15))
    return DCA_ERR_UNSUPPORTED;
  if ((peer_up_stage && !peer_up_flag) || (peer_down_stage && !peer_down_flag)) return DCA_ERR_ARG;
  if (!peer_up_stage && !peer_down_stage) return DCA_OK;
  const char* base = (const char*)t;
  HaloMsg up{peer_up_stage ? base + (long long)h * inner_bytes : nullptr, (char*)peer_up_stage, outer,
             rows * inner_bytes, chunk, chunk};
  HaloMsg down{peer_down_stage ? base + (rows - h - (long long)live) * inner_bytes : nullptr, (char*)peer_down_stage, outer,
               rows * inner_bytes, chunk, chunk};
  halo_push_kernel<<<dim3(PUSH_CTAS, 2), HALO_THREADS, 0, (cudaStream_t)stream>>>(
      up, down, (unsigned long long*)peer_up_flag, (unsigned long long*)peer_down_flag);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

// Waits until the own arrival counters reach `target`, then fills the `live` halo rows next to the owned ones: rows
// [h-live, h) of `t` from `stage_top`, rows [rows-h, rows-h+live) from `stage_bottom` (0 = image border on that side:
// nothing is waited for or written).
extern "C" int dca_halo_wait_unpack(void* t, long long outer, long long rows, long long inner_bytes, int h, int live,
                                    const void* stage_top, const void* stage_bottom, const void* flag_top,
                                    const void* flag_bottom, unsigned long long target, void* err, void* stream) {
  if (!t || !err || outer <= 0 || h <= 0 || live <= 0 || live > h || rows < 2LL * h + live || inner_bytes <= 0)
    return DCA_ERR_ARG;
  const long long chunk = (long long)live * inner_bytes;
  if ((chunk & 15) || ((rows * inner_bytes) & 15) || (inner_bytes & 15) || ((uintptr_t)t & 15) ||
      ((uintptr_t)stage_top & 15) || ((uintptr_t)stage_bottom & 15))
    return DCA_ERR_UNSUPPORTED;
  if ((stage_top && !flag_top) || (stage_bottom && !flag_bottom)) return DCA_ERR_ARG;
  if (!stage_top && !stage_bottom) return DCA_OK;
  char* base = (char*)t;
  HaloMsg top{(const char*)stage_top, base + (long long)(h - live) * inner_bytes, outer, chunk, rows * inner_bytes, chunk};
  HaloMsg bottom{(const char*)stage_bottom, base + (rows - (long long)h) * inner_bytes, outer, chunk,
                 rows * inner_bytes, chunk};
  halo_wait_unpack_kernel<<<dim3(PUSH_CTAS, 2), HALO_THREADS, 0, (cudaStream_t)stream>>>(
      top, bottom, (const unsigned long long*)flag_top, (const unsigned long long*)flag_bottom, target, (int*)err,
      g_halo_timeout_ns);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}

// dca_halo_push + dca_halo_wait_unpack as ONE launch (same arguments, same semantics): the default of hshard.PeerHalo.
extern "C" int dca_halo_exchange(void* t, long long outer, long long rows, long long inner_bytes, int h, int live,
                                 void* peer_up_stage, void* peer_down_stage, void* peer_up_flag, void* peer_down_flag,
                                 const void* stage_top, const void* stage_bottom, const void* flag_top,
                                 const void* flag_bottom, unsigned long long target, void* err, void* stream) {
  if (!t || !err || outer <= 0 || h <= 0 || live <= 0 || live > h || rows < 2LL * h + live || inner_bytes <= 0)
    return DCA_ERR_ARG;
  const long long chunk = (long long)live * inner_bytes;
  if ((chunk & 15) || ((rows * inner_bytes) & 15) || (inner_bytes & 15) || ((uintptr_t)t & 15) ||
      ((uintptr_t)peer_up_stage & 15) || ((uintptr_t)peer_down_stage & 15) || ((uintptr_t)stage_top & 15) ||
      ((uintptr_t)stage_bottom & 15))
    return DCA_ERR_UNSUPPORTED;
  if ((peer_up_stage && !peer_up_flag) || (peer_down_stage && !peer_down_flag) || (stage_top && !flag_top) ||
      (stage_bottom && !flag_bottom))
    return DCA_ERR_ARG;
  if (!peer_up_stage && !peer_down_stage && !stage_top && !stage_bottom) return DCA_OK;
  char* base = (char*)t;
  HaloMsg up{peer_up_stage ? base + (long long)h * inner_bytes : nullptr, (char*)peer_up_stage, outer, rows * inner_bytes,
             chunk, chunk};
  HaloMsg down{peer_down_stage ? base + (rows - h - (long long)live) * inner_bytes : nullptr, (char*)peer_down_stage, outer,
               rows * inner_bytes, chunk, chunk};
  HaloMsg top{(const char*)stage_top, base + (long long)(h - live) * inner_bytes, outer, chunk, rows * inner_bytes, chunk};
  HaloMsg bottom{(const char*)stage_bottom, base + (rows - (long long)h) * inner_bytes, outer, chunk, rows * inner_bytes,
                 chunk};
  dca_launch(halo_exchange_kernel, dim3(PUSH_CTAS, 2), HALO_THREADS, 0, (cudaStream_t)stream, up, down,
             (unsigned long long*)peer_up_flag, (unsigned long long*)peer_down_flag, top, bottom,
             (const unsigned long long*)flag_top, (const unsigned long long*)flag_bottom, target, (int*)err,
             g_halo_timeout_ns);
  DCA_RETURN_IF_LAUNCH_FAILED();
  return DCA_OK;
}
