// Shared device/host helpers for the pcm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcm_b200.h"

namespace pcm {

// ---- error plumbing: the C ABI never aborts; it returns a status and keeps a message ---------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PCM_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      pcm::set_error(__VA_ARGS__);              \
      return PCM_ERR_INVALID;                   \
    }                                           \
  } while (0)

// ---- programmatic dependent launch (default on; PCM_PDL=0 off) ------------------------------------------------------
// Every kernel of the library goes through pcm::launch and follows the dependent-launch protocol: it executes
// `griddepcontrol.wait` (PCM_PDL_ENTRY / pdl_wait) before it touches global memory and before any thread exits, so that
// "this grid completed" always implies "its predecessors completed".  Unless PCM_PDL=0 the launches carry the
// programmatic-stream-serialization attribute (the edges of the captured CUDA graph become programmatic) and the next
// kernel of the stream may begin launching before this one has drained.
// Measured on B200 (flagship step, CUDA graph replay):
//   round 1 (120 dependent launches, one priority level): classic launches 2.11-2.16 ms | attribute, trigger at grid
//   completion 2.11 ms | attribute + early trigger (`griddepcontrol.launch_dependents` at kernel entry, PCM_PDL_EARLY=1)
//   2.31 ms — dependents parked in griddepcontrol.wait are released later than a fresh launch would start;
//   round 2 (108 launches, critical chain on a high-priority stream): classic 1.666 ms | attribute 1.621 ms (-2.7 %: the
//   ~85 kernels of the chain each start ~0.5 us earlier) — so the attribute is now the DEFAULT (PCM_PDL=0 turns it off),
//   the early trigger stays off.  tests/test_cpu_boundary.py checks that every kernel follows the protocol (no global
//   memory access before griddepcontrol.wait, no launch that bypasses pcm::launch).
#ifndef PCM_PDL_EARLY
#define PCM_PDL_EARLY 0          // 1: kernels also issue griddepcontrol.launch_dependents at entry
#endif
__device__ __forceinline__ void pdl_launch_dependents() {
#if PCM_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define PCM_PDL_ENTRY()            \
  do {                             \
    pcm::pdl_launch_dependents();  \
    pcm::pdl_wait();               \
  } while (0)

bool pdl_enabled();

// Dropout epoch: ONE device cell owned by the library (allocated on first use, zero).  Every kernel that draws a
// counter-based dropout mask mixes it into its seed, and pcm_dropout_epoch_advance bumps it with a one-thread kernel —
// inside a captured CUDA graph the masks therefore change on every replay although the scalar seeds are frozen in the
// graph, while forward and backward of one step (same epoch) still regenerate identical masks.
unsigned long long* dropout_epoch_cell();
__device__ __forceinline__ unsigned long long mix_epoch(unsigned long long seed, const unsigned long long* epoch) {
  return seed + 0xD1342543DE82EF95ull * __ldg(epoch);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

// the same with a thread-block cluster of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

// dtype dispatch for activation storage: 0 = f32, 1 = bf16
#define PCM_DISPATCH_DTYPE(dtype, T, ...)                          \
  do {                                                             \
    if ((dtype) == PCM_F32) { using T = float; __VA_ARGS__; }      \
    else if ((dtype) == PCM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { pcm::set_error("bad dtype %d", (int)(dtype)); return PCM_ERR_INVALID; } \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- 8-channel vector access (all channel counts are padded to a multiple of 8) --------------
__device__ __forceinline__ void load8(const float* __restrict__ p, float v[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* __restrict__ p, float v[8]) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float v[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
// value as the storage type would round it (so saved/recomputed quantities agree bit-for-bit)
template <typename T> __device__ __forceinline__ float round_to(float x) { return to_f(from_f<T>(x)); }

// Counter-based uniform in [0,1) for the dropout masks: element `idx` of the stream `seed` (backward regenerates the
// same mask from the same pair).  32-bit arithmetic — a Weyl step on the index, then the lowbias32 finalizer (two
// multiplies, three xor-shifts): the 64-bit splitmix finalizer used before cost ~3x the integer instructions and
// dominated the attention kernels with p > 0 (forward 39 -> 67 us, backward 54 -> 86 us per layer).
__device__ __forceinline__ float dropout_uniform(unsigned long long seed, unsigned long long idx) {
  uint32_t x = (uint32_t)idx * 0x9E3779B1u + (uint32_t)seed;
  x ^= ((uint32_t)(idx >> 32) + (uint32_t)(seed >> 32)) * 0x85EBCA77u;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// ---- reductions ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; result valid in thread 0 (smem: >= 32 entries of T)
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
  if (wid == 0) v = warp_sum(v);
  return v;
}

}  // namespace pcm
