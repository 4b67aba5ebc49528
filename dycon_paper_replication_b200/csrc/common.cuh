// Shared host/device helpers for the DyCON B200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "dycon_b200.h"

namespace dycon {

// ---- thread-local error string (dycon_last_error) ------------------------------------------
char* error_buffer();
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);

#define DYCON_REQUIRE(cond, code, ...)                       \
  do {                                                       \
    if (!(cond)) return ::dycon::fail((code), __VA_ARGS__);  \
  } while (0)

#define DYCON_CUDA(expr)                                               \
  do {                                                                 \
    cudaError_t _e = (expr);                                           \
    if (_e != cudaSuccess) return ::dycon::cuda_fail(_e, #expr);       \
  } while (0)

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline cudaStream_t as_stream(dycon_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // SMs of the current device (cached per device)
void count_launches(int n);  // feeds dycon_launch_count()
uint64_t launches();

constexpr int kMaxPartials = 4096;  // per-block partial slots in a reduction workspace

// Workspace layout shared by the two-stage reductions: [0] ticket counter (u32), then
// kMaxPartials * kLanes doubles of per-block partials starting at byte 16.
struct ReduceWorkspace {
  unsigned int* ticket;
  double* partials;
};
inline ReduceWorkspace carve_reduce_workspace(void* ws) {
  return {reinterpret_cast<unsigned int*>(ws),
          reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 16)};
}

#ifdef __CUDACC__
// ---- device-side reductions ----------------------------------------------------------------
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

// Block-wide sum of kLanes values per thread.  Fixed order -> bit-reproducible.  Result valid
// in thread 0.  `scratch` holds kLanes * 32 doubles.
template <int kLanes>
__device__ __forceinline__ void block_sum(double (&v)[kLanes], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < kLanes; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();  // scratch may still be read by a previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kLanes; ++k) scratch[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < kLanes; ++k) {
      double x = lane < nwarp ? scratch[k * 32 + lane] : 0.0;
      v[k] = warp_sum(x);
    }
  }
}

// Second stage: the block that takes the last ticket sums every block's partials in index
// order and returns true (in all of its threads) with the totals in thread 0's `total`.
// The ticket is reset so the workspace is left zeroed (re-entrant across launches).
// `last_flag`: one int of shared memory (kernels whose dynamic shared memory is 1024-byte aligned and nearly
// full pass a slot of it: even one byte of static shared memory would cost them a whole kilobyte).
template <int kLanes>
__device__ __forceinline__ bool grid_sum_last_block(double (&v)[kLanes], double (&total)[kLanes],
                                                    unsigned int* ticket, double* partials,
                                                    unsigned int nblocks, unsigned int block_id,
                                                    double* scratch, int* last_flag) {
  block_sum<kLanes>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kLanes; ++k) partials[(size_t)block_id * kLanes + k] = v[k];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    *last_flag = (t == nblocks - 1);
  }
  __syncthreads();
  if (!*last_flag) return false;
  __threadfence();
  double acc[kLanes];
#pragma unroll
  for (int k = 0; k < kLanes; ++k) acc[k] = 0.0;
  for (unsigned int b = threadIdx.x; b < nblocks; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < kLanes; ++k) acc[k] += __ldcg(&partials[(size_t)b * kLanes + k]);
  }
  block_sum<kLanes>(acc, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kLanes; ++k) total[k] = acc[k];
    *ticket = 0u;
  }
  return true;
}
template <int kLanes>
__device__ __forceinline__ bool grid_sum_last_block(double (&v)[kLanes], double (&total)[kLanes],
                                                    unsigned int* ticket, double* partials,
                                                    unsigned int nblocks, unsigned int block_id,
                                                    double* scratch) {
  __shared__ int is_last;
  return grid_sum_last_block<kLanes>(v, total, ticket, partials, nblocks, block_id, scratch, &is_last);
}
#endif  // __CUDACC__

}  // namespace dycon
