// fc_common.cuh — launch plumbing shared by the .cu files (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fc_math.cuh"

namespace fc {

constexpr int kThreads = 256;           // 8 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kSmemTargetBytes = 40 * 1024;   // staged parameter tile per CTA (v1 kernels)
constexpr int kSmemMaxBytes = 200 * 1024;

struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};

inline const DeviceInfo& device_info() {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo info;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cached_dev = dev;
  }
  return info;
}

inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// How the lanes of a warp are laid over the D_t features of a row (see DESIGN.md "lane mapping"):
// seg = min(32, next_pow2(D_t)) lanes cover one row; a warp covers 32/seg rows per pass.
struct LaneMap {
  int seg;            // lanes per row segment (power of two)
  int rows_per_warp;  // 32 / seg
};

inline LaneMap lane_map(int d_t) {
  LaneMap m;
  m.seg = d_t >= 32 ? 32 : next_pow2(d_t);
  m.rows_per_warp = 32 / m.seg;
  return m;
}

// Rows per staged tile so that the parameter tile is about kSmemTargetBytes and every warp has work.
inline int tile_rows(int row_floats, const LaneMap& m, int64_t B) {
  const int64_t row_bytes = (int64_t)row_floats * 4;
  const int r0 = kWarps * m.rows_per_warp;
  int64_t mult = kSmemTargetBytes / (r0 * row_bytes);
  if (mult < 1) mult = 1;
  int64_t rows = r0 * mult;
  while (rows > 1 && rows * row_bytes > kSmemMaxBytes) rows >>= 1;
  if (rows > B) rows = B > 0 ? B : 1;
  return (int)rows;
}

__device__ __forceinline__ float seg_reduce_sum(float v, int seg) {
  // xor-shuffle sum over aligned segments of `seg` lanes (seg is a power of two <= 32)
  if (seg == 32) {  // the common case (D_t >= 32), fully unrolled
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }
  for (int off = seg >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Cooperative global -> shared copy of `n` floats (16-byte vectors when both sides allow it).
__device__ __forceinline__ void stage_in(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int64_t n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = __ldcs(s4 + i);
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcs(src + i);
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcs(src + i);
  }
}

__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int64_t n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) __stcs(d4 + i, s4[i]);
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) __stcs(dst + i, src[i]);
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) __stcs(dst + i, src[i]);
  }
}

#define FC_CHECK_LAUNCH()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return FC_ERR_CUDA;   \
  } while (0)

}  // namespace fc
