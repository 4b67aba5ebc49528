// K3 -- multi-tensor mean-teacher EMA (reference: code/train_DyCON_BraTS19.py:155-164).
//
// The reference walks the parameter list in Python and launches mul_ + add_ per tensor
// (96 launches for the 48 tensors of unet_3D, most of them a few hundred bytes).  Here the
// pointer table travels in the kernel's parameter space (no host->device copy, capturable in a
// CUDA graph) and one launch covers every tensor: a block owns a 4096-element chunk of one
// tensor, found by binary search over the per-tensor block offsets.  HBM-bound: 12 B/param.
//
// Rounding order matches ema.mul_(alpha).add_(p, alpha=1-alpha):  t = rn(ema*alpha), then
// ema = fma(1-alpha, p, t) (ATen's CUDA add-with-alpha functor contracts a + alpha*b to one FMA).
#include "common.cuh"

namespace dycon {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;        // elements per block: 256 threads x 4 float4
constexpr int kMaxTensors = 320;    // per launch; 320*(8+8+8+4) B = 8960 B of kernel parameters

struct EmaTable {
  float* ema[kMaxTensors];
  const float* param[kMaxTensors];
  long long numel[kMaxTensors];
  int block_start[kMaxTensors + 1];
  int n;
};

__global__ void __launch_bounds__(kThreads)
ema_multi_kernel(const __grid_constant__ EmaTable tab, float alpha, float oma) {
  // block -> tensor: largest k with block_start[k] <= blockIdx.x
  int lo = 0, hi = tab.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.block_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  float* __restrict__ e = tab.ema[lo];
  const float* __restrict__ p = tab.param[lo];
  const long long n = tab.numel[lo];
  const long long base = (long long)(blockIdx.x - tab.block_start[lo]) * kChunk;
  const long long end = base + kChunk < n ? base + kChunk : n;
  const bool vec = ((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(p)) & 15) == 0;
  if (vec && end - base == kChunk) {
    float4 ev[4], pv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = base + (long long)(k * kThreads + threadIdx.x) * 4;
      ev[k] = *reinterpret_cast<const float4*>(e + i);
      pv[k] = __ldcs(reinterpret_cast<const float4*>(p + i));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = base + (long long)(k * kThreads + threadIdx.x) * 4;
      float4 o;
      o.x = fmaf(oma, pv[k].x, __fmul_rn(ev[k].x, alpha));
      o.y = fmaf(oma, pv[k].y, __fmul_rn(ev[k].y, alpha));
      o.z = fmaf(oma, pv[k].z, __fmul_rn(ev[k].z, alpha));
      o.w = fmaf(oma, pv[k].w, __fmul_rn(ev[k].w, alpha));
      *reinterpret_cast<float4*>(e + i) = o;
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += kThreads) e[i] = fmaf(oma, p[i], __fmul_rn(e[i], alpha));
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
    long long blocks = 0;
    while (k < n_tensors && tab.n < kMaxTensors) {
      const long long nb = (numels[k] + kChunk - 1) / kChunk;
      if (nb == 0) { ++k; continue; }
      if (blocks + nb > 0x7fffffffLL / 2) break;
      tab.ema[tab.n] = ema_ptrs[k];
      tab.param[tab.n] = param_ptrs[k];
      tab.numel[tab.n] = numels[k];
      tab.block_start[tab.n] = (int)blocks;
      blocks += nb;
      ++tab.n;
      ++k;
    }
    if (tab.n == 0) {
      DYCON_REQUIRE(k >= n_tensors, DYCON_ERR_UNSUPPORTED, "EMA: tensor %d too large for one launch", k);
      break;
    }
    tab.block_start[tab.n] = (int)blocks;
    ema_multi_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(tab, alpha, one_minus_alpha);
    DYCON_CUDA(cudaGetLastError());
    count_launches(1);
  }
  return DYCON_OK;
}
