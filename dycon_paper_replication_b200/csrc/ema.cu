// K3 -- multi-tensor mean-teacher EMA (reference: code/train_DyCON_BraTS19.py:155-164).
//
// The reference walks the parameter list in Python and launches mul_ + add_ per tensor
// (96 launches for the 48 tensors of unet_3D, most of them a few hundred bytes).  Here the
// pointer table travels in the kernel's parameter space (no host->device copy, capturable in a
// CUDA graph) and one launch covers every tensor: the tensors are cut into 1024-element
// chunks, a block owns a contiguous range of chunks and finds its first tensor by binary search.  HBM-bound: 12 B/param.
//
// Rounding order matches ema.mul_(alpha).add_(p, alpha=1-alpha):  t = rn(ema*alpha), then
// ema = fma(1-alpha, p, t) (ATen's CUDA add-with-alpha functor contracts a + alpha*b to one FMA).
#include "common.cuh"

namespace dycon {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 1024;        // elements per chunk: 256 threads x one float4
constexpr int kMaxTensors = 320;    // per launch; 320*(8+8+8+4) B = 8960 B of kernel parameters

struct EmaTable {
  float* ema[kMaxTensors];
  const float* param[kMaxTensors];
  long long numel[kMaxTensors];
  int chunk_start[kMaxTensors + 1];
  int n;
};

__device__ __forceinline__ float ema_one(float e, float p, float alpha, float oma) {
  return fmaf(oma, p, __fmul_rn(e, alpha));
}

// Persistent: every block owns an equal, contiguous range of chunks (one resident wave, no tail wave),
// and keeps two chunks (4 x 16 B per thread) in flight.
__global__ void __launch_bounds__(kThreads)
ema_multi_kernel(const __grid_constant__ EmaTable tab, float alpha, float oma, int total_chunks) {
  const int per = total_chunks / (int)gridDim.x, rem = total_chunks % (int)gridDim.x;
  const int bx = (int)blockIdx.x;
  const int c_begin = bx * per + (bx < rem ? bx : rem), c_end = c_begin + per + (bx < rem ? 1 : 0);
  if (c_begin >= c_end) return;
  int k = 0;   // tensor of chunk c_begin: largest k with chunk_start[k] <= c_begin
  {
    int lo = 0, hi = tab.n;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tab.chunk_start[mid] <= c_begin) lo = mid; else hi = mid;
    }
    k = lo;
  }
  for (int c = c_begin; c < c_end; c += 2) {
    float* e[2];
    const float* p[2];
    long long left[2];
    bool vec[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int cu = c + u;
      if (cu < c_end) {
        while (tab.chunk_start[k + 1] <= cu) ++k;
        const long long base = (long long)(cu - tab.chunk_start[k]) * kChunk;
        e[u] = tab.ema[k] + base;
        p[u] = tab.param[k] + base;
        left[u] = tab.numel[k] - base;
        vec[u] = left[u] >= kChunk && ((reinterpret_cast<uintptr_t>(e[u]) | reinterpret_cast<uintptr_t>(p[u])) & 15) == 0;
      } else {
        e[u] = nullptr; p[u] = nullptr; left[u] = 0; vec[u] = false;
      }
    }
    float4 ev[2], pv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (vec[u]) {
        ev[u] = *reinterpret_cast<const float4*>(e[u] + threadIdx.x * 4);
        pv[u] = __ldcs(reinterpret_cast<const float4*>(p[u] + threadIdx.x * 4));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (vec[u]) {
        float4 o;
        o.x = ema_one(ev[u].x, pv[u].x, alpha, oma);
        o.y = ema_one(ev[u].y, pv[u].y, alpha, oma);
        o.z = ema_one(ev[u].z, pv[u].z, alpha, oma);
        o.w = ema_one(ev[u].w, pv[u].w, alpha, oma);
        *reinterpret_cast<float4*>(e[u] + threadIdx.x * 4) = o;
      } else {
        const long long n = left[u] < kChunk ? left[u] : kChunk;
        for (long long i = threadIdx.x; i < n; i += kThreads) e[u][i] = ema_one(e[u][i], p[u][i], alpha, oma);
      }
    }
  }
}

}  // namespace
}  // namespace dycon

using namespace dycon;

extern "C" int dycon_ema_multi(float* const* ema_ptrs, const float* const* param_ptrs, const int64_t* numels,
                               int n_tensors, float alpha, float one_minus_alpha, dycon_stream_t stream) {
  DYCON_REQUIRE(n_tensors >= 0, DYCON_ERR_ARG, "EMA: n_tensors=%d", n_tensors);
  if (n_tensors == 0) return DYCON_OK;
  DYCON_REQUIRE(ema_ptrs && param_ptrs && numels, DYCON_ERR_ARG, "EMA: NULL table");
  for (int k = 0; k < n_tensors; ++k) {
    DYCON_REQUIRE(numels[k] >= 0, DYCON_ERR_ARG, "EMA: numels[%d] < 0", k);
    DYCON_REQUIRE(numels[k] == 0 || (ema_ptrs[k] && param_ptrs[k]), DYCON_ERR_ARG, "EMA: NULL tensor %d", k);
    DYCON_REQUIRE(aligned(ema_ptrs[k], 4) && aligned(param_ptrs[k], 4), DYCON_ERR_ARG, "EMA: tensor %d misaligned", k);
  }
  cudaStream_t st = as_stream(stream);
  int k = 0;
  while (k < n_tensors) {
    EmaTable tab;
    tab.n = 0;
    long long blocks = 0;   // chunks
    while (k < n_tensors && tab.n < kMaxTensors) {
      const long long nb = (numels[k] + kChunk - 1) / kChunk;
      if (nb == 0) { ++k; continue; }
      if (blocks + nb > 0x7fffffffLL / 2) break;
      tab.ema[tab.n] = ema_ptrs[k];
      tab.param[tab.n] = param_ptrs[k];
      tab.numel[tab.n] = numels[k];
      tab.chunk_start[tab.n] = (int)blocks;
      blocks += nb;
      ++tab.n;
      ++k;
    }
    if (tab.n == 0) {
      DYCON_REQUIRE(k >= n_tensors, DYCON_ERR_UNSUPPORTED, "EMA: tensor %d too large for one launch", k);
      break;
    }
    tab.chunk_start[tab.n] = (int)blocks;
    // one resident wave (occupancy-derived blocks per SM), or fewer blocks when the work is small
    static const int per_sm = [] {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ema_multi_kernel, kThreads, 0) != cudaSuccess || n < 1) n = 4;
      return n;
    }();
    long long grid = (long long)sm_count() * per_sm;
    if (grid > blocks) grid = blocks;
    ema_multi_kernel<<<(unsigned)grid, kThreads, 0, st>>>(tab, alpha, one_minus_alpha, (int)blocks);
    DYCON_CUDA(cudaGetLastError());
    count_launches(1);
  }
  return DYCON_OK;
}
