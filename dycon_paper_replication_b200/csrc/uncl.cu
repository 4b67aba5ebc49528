// K1 -- fused UnCL forward / backward (reference: code/utils/dycon_losses.py:94-118).
//
// Per voxel v (channels c):  ps = softmax(s), pt = softmax(t), H = -sum p*log(p + 1e-6),
//   L_v = sum_c (ps-pt)^2 / (exp(beta*Hs) + exp(beta*Ht)) + beta*(Hs + Ht),   loss = mean_v L_v.
// The reference's (B,B,...) broadcast at :116 only replicates terms, so mean_v L_v is its value.
//
// HBM-bound: one coalesced float4 pass over the four logit streams (C == 2).  The forward also
// writes the *unit* gradient u_v = dL_v/ds[:,1,v] (4 B/voxel; ds[:,0] = -ds[:,1] for C == 2), so
// the backward is a 4-byte read + 8-byte write instead of a 16-byte re-read:
//   fwd 16 B read + 4 B write, bwd 4 B read + 8 B write = 32 B/voxel  (recompute design: 40 B/voxel).
// Sums are two-stage and fixed-order (bit-reproducible); the last block to finish reduces the
// per-block partials, so there is no second launch and no float atomics.
#include "common.cuh"

namespace dycon {
namespace {

constexpr float kEps = 1e-6f;  // dycon_losses.py:95 -- NOT a numerical no-op, keep inside the log
constexpr int kThreads = 256;

// ---- C == 2 fast path ------------------------------------------------------------------------
__device__ __forceinline__ void softmax2(float x0, float x1, float& p0, float& p1) {
  const float d = x1 - x0;
  const float e = __expf(-fabsf(d));            // (0, 1]
  const float hi = __fdividef(1.f, 1.f + e);    // probability of the larger logit
  const float lo = e * hi;
  p1 = d >= 0.f ? hi : lo;
  p0 = d >= 0.f ? lo : hi;
}

// Returns L_v and the unit gradient dL_v/ds1 (SURVEY.md section 0.1 closed form).
__device__ __forceinline__ void voxel2(float s0, float s1, float t0, float t1, float beta, float& L,
                                       float& u) {
  float ps0, ps1, pt0, pt1;
  softmax2(s0, s1, ps0, ps1);
  softmax2(t0, t1, pt0, pt1);
  const float ls0 = __logf(ps0 + kEps), ls1 = __logf(ps1 + kEps);
  const float lt0 = __logf(pt0 + kEps), lt1 = __logf(pt1 + kEps);
  const float hs = -(ps0 * ls0 + ps1 * ls1);
  const float ht = -(pt0 * lt0 + pt1 * lt1);
  const float es = __expf(beta * hs), et = __expf(beta * ht);
  const float w = __fdividef(1.f, es + et);
  const float d0 = ps0 - pt0, d1 = ps1 - pt1;
  const float q = d0 * d0 + d1 * d1;
  L = fmaf(q, w, beta * (hs + ht));
  const float dl_dh = beta - q * beta * es * w * w;
  const float g0 = 2.f * d0 * w - dl_dh * (ls0 + __fdividef(ps0, ps0 + kEps));
  const float g1 = 2.f * d1 * w - dl_dh * (ls1 + __fdividef(ps1, ps1 + kEps));
  u = ps0 * ps1 * (g1 - g0);  // ps1*(g1 - ps0*g0 - ps1*g1) with ps0 + ps1 = 1
}

template <int kVec>
struct Pack;
template <>
struct Pack<4> {
  using type = float4;
  static __device__ __forceinline__ float4 load(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
};
template <>
struct Pack<1> {
  using type = float;
  static __device__ __forceinline__ float load(const float* p) { return __ldcs(p); }
};
__device__ __forceinline__ float elem(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ float elem(const float& v, int) { return v; }
__device__ __forceinline__ void set_elem(float4& v, int k, float x) {
  if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x;
}
__device__ __forceinline__ void set_elem(float& v, int, float x) { v = x; }

template <int kVec>
__global__ void __launch_bounds__(kThreads)
uncl_fwd_c2_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t B, int64_t V, float beta,
                   double inv_count, float* __restrict__ stash, unsigned int* ticket, double* partials,
                   double* __restrict__ sum_out, float* __restrict__ loss_out) {
  using P = Pack<kVec>;
  using vec_t = typename P::type;
  __shared__ double scratch[32];
  const int64_t nvec = V / kVec;
  float acc = 0.f;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    const float* s0 = s + (2 * b) * V;
    const float* s1 = s0 + V;
    const float* t0 = t + (2 * b) * V;
    const float* t1 = t0 + V;
    float* st = stash + b * V;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kThreads) {
      const vec_t a0 = P::load(s0 + i * kVec), a1 = P::load(s1 + i * kVec);
      const vec_t b0 = P::load(t0 + i * kVec), b1 = P::load(t1 + i * kVec);
      vec_t uo;
      float lsum = 0.f;
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        float L, u;
        voxel2(elem(a0, k), elem(a1, k), elem(b0, k), elem(b1, k), beta, L, u);
        lsum += L;
        set_elem(uo, k, u);
      }
      *reinterpret_cast<vec_t*>(st + i * kVec) = uo;  // default policy: re-read by the backward from L2
      acc += lsum;
    }
  }
  double v[1] = {(double)acc}, total[1];
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (grid_sum_last_block<1>(v, total, ticket, partials, nblocks, bid, scratch) && threadIdx.x == 0) {
    *sum_out = total[0];
    if (loss_out) *loss_out = (float)(total[0] * inv_count);
  }
}

template <int kVec>
__global__ void __launch_bounds__(kThreads)
uncl_bwd_c2_kernel(const float* __restrict__ stash, int64_t B, int64_t V, float inv_count,
                   const float* __restrict__ grad_out, float* __restrict__ grad_s) {
  using vec_t = typename Pack<kVec>::type;
  const float scale = __ldg(grad_out) * inv_count;
  const int64_t nvec = V / kVec;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    const float* st = stash + b * V;
    float* g0 = grad_s + (2 * b) * V;
    float* g1 = g0 + V;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kThreads) {
      const vec_t u = *reinterpret_cast<const vec_t*>(st + i * kVec);
      vec_t p, n;
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        const float x = elem(u, k) * scale;
        set_elem(p, k, x);
        set_elem(n, k, -x);
      }
      *reinterpret_cast<vec_t*>(g0 + i * kVec) = n;
      *reinterpret_cast<vec_t*>(g1 + i * kVec) = p;
    }
  }
}

// ---- generic C (scalar per voxel, channels re-read through L1; libm-accurate math) -----------
template <bool kBackward>
__device__ __forceinline__ float voxel_generic(const float* __restrict__ sp, const float* __restrict__ tp, int64_t V,
                                               int C, float beta, float scale, float* __restrict__ gp) {
  float ms = -INFINITY, mt = -INFINITY;
  for (int c = 0; c < C; ++c) {
    ms = fmaxf(ms, sp[c * V]);
    mt = fmaxf(mt, tp[c * V]);
  }
  float zs = 0.f, zt = 0.f;
  for (int c = 0; c < C; ++c) {
    zs += expf(sp[c * V] - ms);
    zt += expf(tp[c * V] - mt);
  }
  const float rs = 1.f / zs, rt = 1.f / zt;
  float hs = 0.f, ht = 0.f, q = 0.f;
  for (int c = 0; c < C; ++c) {
    const float ps = expf(sp[c * V] - ms) * rs, pt = expf(tp[c * V] - mt) * rt;
    hs -= ps * logf(ps + kEps);
    ht -= pt * logf(pt + kEps);
    q += (ps - pt) * (ps - pt);
  }
  const float es = expf(beta * hs), et = expf(beta * ht);
  const float w = 1.f / (es + et);
  const float L = q * w + beta * (hs + ht);
  if (kBackward) {
    const float dl_dh = beta - q * beta * es * w * w;
    float spg = 0.f;
    for (int c = 0; c < C; ++c) {
      const float ps = expf(sp[c * V] - ms) * rs, pt = expf(tp[c * V] - mt) * rt;
      const float g = 2.f * (ps - pt) * w - dl_dh * (logf(ps + kEps) + ps / (ps + kEps));
      spg += ps * g;
    }
    for (int c = 0; c < C; ++c) {
      const float ps = expf(sp[c * V] - ms) * rs, pt = expf(tp[c * V] - mt) * rt;
      const float g = 2.f * (ps - pt) * w - dl_dh * (logf(ps + kEps) + ps / (ps + kEps));
      gp[c * V] = scale * ps * (g - spg);
    }
  }
  return L;
}

__global__ void __launch_bounds__(kThreads)
uncl_fwd_generic_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t B, int C, int64_t V,
                        float beta, double inv_count, unsigned int* ticket, double* partials,
                        double* __restrict__ sum_out, float* __restrict__ loss_out) {
  __shared__ double scratch[32];
  float acc = 0.f;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < V; i += (int64_t)gridDim.x * kThreads) {
      acc += voxel_generic<false>(s + b * C * V + i, t + b * C * V + i, V, C, beta, 0.f, nullptr);
    }
  }
  double v[1] = {(double)acc}, total[1];
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (grid_sum_last_block<1>(v, total, ticket, partials, nblocks, bid, scratch) && threadIdx.x == 0) {
    *sum_out = total[0];
    if (loss_out) *loss_out = (float)(total[0] * inv_count);
  }
}

__global__ void __launch_bounds__(kThreads)
uncl_bwd_generic_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t B, int C, int64_t V,
                        float beta, float inv_count, const float* __restrict__ grad_out, float* __restrict__ grad_s) {
  const float scale = __ldg(grad_out) * inv_count;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < V; i += (int64_t)gridDim.x * kThreads) {
      voxel_generic<true>(s + b * C * V + i, t + b * C * V + i, V, C, beta, scale, grad_s + b * C * V + i);
    }
  }
}

// grid.x blocks per sample, grid.y sample lanes; ~8 resident CTAs per SM, at most kMaxPartials blocks.
dim3 pick_grid(int64_t B, int64_t work_items_per_sample) {
  const int64_t target = (int64_t)sm_count() * 8;
  int64_t gy = B < 1024 ? B : 1024;
  int64_t gx = (work_items_per_sample + kThreads - 1) / kThreads;
  int64_t cap = target / gy;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  while (gx * gy > kMaxPartials) {
    if (gx > 1) --gx; else --gy;
  }
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

int check_common(int64_t B, int C, int64_t V) {
  DYCON_REQUIRE(B > 0 && V > 0, DYCON_ERR_ARG, "UnCL: B=%lld V=%lld must be positive", (long long)B, (long long)V);
  DYCON_REQUIRE(C >= 2 && C <= 1024, DYCON_ERR_UNSUPPORTED, "UnCL: C=%d outside [2, 1024]", C);
  return DYCON_OK;
}

}  // namespace
}  // namespace dycon

using namespace dycon;

extern "C" {

size_t dycon_uncl_workspace_bytes(void) { return 16 + sizeof(double) * kMaxPartials; }

int dycon_uncl_fwd(const float* s, const float* t, int64_t B, int C, int64_t V, float beta, double inv_count,
                   float* stash, double* sum_out, float* loss_out, void* workspace, size_t workspace_bytes,
                   dycon_stream_t stream) {
  if (int rc = check_common(B, C, V)) return rc;
  DYCON_REQUIRE(s && t && sum_out && workspace, DYCON_ERR_ARG, "UnCL fwd: NULL s/t/sum_out/workspace");
  DYCON_REQUIRE(aligned(s, 4) && aligned(t, 4) && aligned(workspace, 16) && aligned(sum_out, 8), DYCON_ERR_ARG,
                "UnCL fwd: misaligned pointer");
  DYCON_REQUIRE(workspace_bytes >= dycon_uncl_workspace_bytes(), DYCON_ERR_WORKSPACE,
                "UnCL fwd: workspace %zu < %zu bytes", workspace_bytes, dycon_uncl_workspace_bytes());
  ReduceWorkspace ws = carve_reduce_workspace(workspace);
  cudaStream_t st = as_stream(stream);
  if (C == 2) {
    DYCON_REQUIRE(stash && aligned(stash, 4), DYCON_ERR_ARG, "UnCL fwd: C == 2 needs a stash of B*V floats");
    const bool vec = (V % 4 == 0) && aligned(s, 16) && aligned(t, 16) && aligned(stash, 16);
    if (vec) {
      dim3 grid = pick_grid(B, V / 4);
      uncl_fwd_c2_kernel<4><<<grid, kThreads, 0, st>>>(s, t, B, V, beta, inv_count, stash, ws.ticket, ws.partials,
                                                        sum_out, loss_out);
    } else {
      dim3 grid = pick_grid(B, V);
      uncl_fwd_c2_kernel<1><<<grid, kThreads, 0, st>>>(s, t, B, V, beta, inv_count, stash, ws.ticket, ws.partials,
                                                        sum_out, loss_out);
    }
  } else {
    dim3 grid = pick_grid(B, V);
    uncl_fwd_generic_kernel<<<grid, kThreads, 0, st>>>(s, t, B, C, V, beta, inv_count, ws.ticket, ws.partials,
                                                       sum_out, loss_out);
  }
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

int dycon_uncl_bwd(const float* s, const float* t, const float* stash, int64_t B, int C, int64_t V, float beta,
                   double inv_count, const float* grad_out, float* grad_s, dycon_stream_t stream) {
  if (int rc = check_common(B, C, V)) return rc;
  DYCON_REQUIRE(grad_out && grad_s, DYCON_ERR_ARG, "UnCL bwd: NULL grad_out/grad_s");
  cudaStream_t st = as_stream(stream);
  if (C == 2) {
    DYCON_REQUIRE(stash, DYCON_ERR_ARG, "UnCL bwd: C == 2 needs the stash written by the forward");
    const bool vec = (V % 4 == 0) && aligned(stash, 16) && aligned(grad_s, 16);
    if (vec) {
      uncl_bwd_c2_kernel<4><<<pick_grid(B, V / 4), kThreads, 0, st>>>(stash, B, V, (float)inv_count, grad_out, grad_s);
    } else {
      uncl_bwd_c2_kernel<1><<<pick_grid(B, V), kThreads, 0, st>>>(stash, B, V, (float)inv_count, grad_out, grad_s);
    }
  } else {
    DYCON_REQUIRE(s && t, DYCON_ERR_ARG, "UnCL bwd: C != 2 recomputes from s/t (NULL given)");
    uncl_bwd_generic_kernel<<<pick_grid(B, V), kThreads, 0, st>>>(s, t, B, C, V, beta, (float)inv_count, grad_out,
                                                                  grad_s);
  }
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // extern "C"
