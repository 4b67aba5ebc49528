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
// MUFU budget (the XU pipe issues 16 lanes/clk/SM, so it -- not HBM -- bounds this kernel unless
// kept to ~12 ops/voxel): flush-to-zero approximations straight from PTX (no denormal fix-up
// code), and the two-class algebra below instead of a generic softmax/log chain.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

// Two-class softmax + entropy terms of one voxel.  With a = |x1-x0|, e = exp(-a), y = eps(1+e):
//   p_hi = 1/(1+e), p_lo = e p_hi,
//   ln(p_hi+eps) = log1p(y) - ln(1+e)      (log1p(y) = y - y^2/2 exactly in fp32, y <= 2e-6)
//   ln(p_lo+eps) = ln(e + y) - ln(1+e)
struct TwoClass {
  float p_hi, p_lo, ln_hi, ln_lo, e, x;   // x = e + y (reused by the gradient: p_lo/(p_lo+eps) = e/x)
  float H;
  bool hi_is_1;
};
__device__ __forceinline__ TwoClass two_class(float x0, float x1) {
  TwoClass o;
  const float d = x1 - x0, a = fabsf(d);
  o.hi_is_1 = d >= 0.f;
  o.e = ex2_approx(-a * kLog2e) + (a - a);          // (a - a): +-inf logits become NaN like torch.softmax
  const float ope = 1.f + o.e;
  o.p_hi = rcp_approx(ope);
  o.p_lo = o.e * o.p_hi;
  const float y = kEps * ope;
  const float l1 = kLn2 * lg2_approx(ope);
  o.ln_hi = fmaf(y, fmaf(-0.5f, y, 1.f), -l1);
  o.x = o.e + y;
  o.ln_lo = fmaf(kLn2, lg2_approx(o.x), -l1);
  o.H = -(o.p_hi * o.ln_hi + o.p_lo * o.ln_lo);
  return o;
}

// Returns L_v and the unit gradient dL_v/ds1 (SURVEY.md section 0.1 closed form).
//   u = ps0 ps1 (g1 - g0),  g_c = 2 (ps_c - pt_c) W - dL/dHs (ln(ps_c+eps) + ps_c/(ps_c+eps))
__device__ __forceinline__ void voxel2(float s0, float s1, float t0, float t1, float beta, float& L,
                                       float& u) {
  const TwoClass S = two_class(s0, s1), T = two_class(t0, t1);
  const float es = ex2_approx(beta * kLog2e * S.H), et = ex2_approx(beta * kLog2e * T.H);
  const float w = rcp_approx(es + et);
  const float ps1 = S.hi_is_1 ? S.p_hi : S.p_lo, ps0 = S.hi_is_1 ? S.p_lo : S.p_hi;
  const float pt1 = T.hi_is_1 ? T.p_hi : T.p_lo, pt0 = T.hi_is_1 ? T.p_lo : T.p_hi;
  const float d0 = ps0 - pt0, d1 = ps1 - pt1;
  const float q = d0 * d0 + d1 * d1;
  L = fmaf(q, w, beta * (S.H + T.H));
  const float dl_dh = beta - q * beta * es * w * w;
  // (ln_hi - ln_lo) + (r_hi - r_lo),  r_hi = 1 - y(1-y),  r_lo = e / (e + y)
  const float y = S.x - S.e;
  const float r_hi = fmaf(-y, 1.f - y, 1.f), r_lo = S.e * rcp_approx(S.x);
  const float hi_minus_lo = (S.ln_hi - S.ln_lo) + (r_hi - r_lo);
  const float g1_minus_g0 = 2.f * (d1 - d0) * w - dl_dh * (S.hi_is_1 ? hi_minus_lo : -hi_minus_lo);
  u = S.p_hi * S.p_lo * g1_minus_g0;
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

// Block + grid reduction tail of the C == 2 forward.  Warps 1.. leave their partial in shared memory and
// retire at once (bar.arrive); only warp 0 waits for them, publishes the block partial and takes the
// ticket, so no warp idles behind the L2 round trip of the atomic.  Fixed order -> bit-reproducible.
__device__ __forceinline__ void uncl_finish(float acc, unsigned int* ticket, double* partials, double inv_count,
                                            double* __restrict__ sum_out, float* __restrict__ loss_out) {
  __shared__ float warp_part[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float w = warp_sum(acc);
  if (lane == 0) warp_part[warp] = w;
  if (warp != 0) {
    __threadfence_block();
    asm volatile("bar.arrive 1, %0;" ::"n"(kThreads) : "memory");
    return;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
  double blk = 0.0;
#pragma unroll
  for (int k = 0; k < kThreads / 32; ++k) blk += (double)warp_part[k];
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  unsigned int t = 0;
  if (lane == 0) {
    partials[bid] = blk;
    __threadfence();
    t = atomicAdd(ticket, 1u);
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t != nblocks - 1) return;
  __threadfence();
  double tot = 0.0;
  for (unsigned int k = lane; k < nblocks; k += 32) tot += __ldcg(&partials[k]);
  tot = warp_sum(tot);
  if (lane == 0) {
    *sum_out = tot;
    if (loss_out) *loss_out = (float)(tot * inv_count);
    *ticket = 0u;
  }
}

// Grid: (blocks per sample, sample lanes).  The loads of iteration k+1 are issued before the arithmetic of
// iteration k (register double buffer): with ~80 instructions and 12 MUFU per voxel the kernel sits between
// the XU pipe and HBM, and a warp that waits for its loads at the top of every iteration leaves both idle.
template <int kVec>
__global__ void __launch_bounds__(kThreads, 3)
uncl_fwd_c2_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t B, int64_t V, float beta,
                   double inv_count, float* __restrict__ stash, unsigned int* ticket, double* partials,
                   double* __restrict__ sum_out, float* __restrict__ loss_out) {
  using P = Pack<kVec>;
  using vec_t = typename P::type;
  const int64_t nvec = V / kVec;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  float acc = 0.f;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    const float* s0 = s + (2 * b) * V;
    const float* s1 = s0 + V;
    const float* t0 = t + (2 * b) * V;
    const float* t1 = t0 + V;
    float* st = stash + b * V;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= nvec) continue;
    vec_t a0 = P::load(s0 + i * kVec), a1 = P::load(s1 + i * kVec);
    vec_t b0 = P::load(t0 + i * kVec), b1 = P::load(t1 + i * kVec);
    while (true) {
      const int64_t nx = i + stride;
      const bool more = nx < nvec;
      vec_t na0, na1, nb0, nb1;
      if (more) {
        na0 = P::load(s0 + nx * kVec); na1 = P::load(s1 + nx * kVec);
        nb0 = P::load(t0 + nx * kVec); nb1 = P::load(t1 + nx * kVec);
      }
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
      if (!more) break;
      a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
      i = nx;
    }
  }
  uncl_finish(acc, ticket, partials, inv_count, sum_out, loss_out);
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
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += 2 * stride) {   // two loads in flight
      const int64_t i2 = i + stride;
      const bool two = i2 < nvec;
      const vec_t u = *reinterpret_cast<const vec_t*>(st + i * kVec);
      vec_t u2 = u;
      if (two) u2 = *reinterpret_cast<const vec_t*>(st + i2 * kVec);
      vec_t p, n, p2, n2;
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        const float x = elem(u, k) * scale, x2 = elem(u2, k) * scale;
        set_elem(p, k, x);
        set_elem(n, k, -x);
        set_elem(p2, k, x2);
        set_elem(n2, k, -x2);
      }
      __stcs(reinterpret_cast<vec_t*>(g0 + i * kVec), n);
      __stcs(reinterpret_cast<vec_t*>(g1 + i * kVec), p);
      if (two) {
        __stcs(reinterpret_cast<vec_t*>(g0 + i2 * kVec), n2);
        __stcs(reinterpret_cast<vec_t*>(g1 + i2 * kVec), p2);
      }
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

// One resident wave: grid.x blocks per sample x grid.y sample lanes = SMs x (CTAs that fit per SM), so
// the grid-stride loops see no second wave and no tail; at most kMaxPartials blocks.
template <typename Kernel>
int resident_ctas(Kernel kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 4;
  return per_sm * sm_count();
}

dim3 pick_grid(int64_t B, int64_t work_items_per_sample, int resident) {
  int64_t gy = B < 1024 ? B : 1024;
  int64_t gx = (work_items_per_sample + kThreads - 1) / kThreads;
  int64_t cap = resident / gy;
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
      static const int res = resident_ctas(uncl_fwd_c2_kernel<4>);
      dim3 grid = pick_grid(B, V / 4, res);
      uncl_fwd_c2_kernel<4><<<grid, kThreads, 0, st>>>(s, t, B, V, beta, inv_count, stash, ws.ticket, ws.partials,
                                                        sum_out, loss_out);
    } else {
      static const int res = resident_ctas(uncl_fwd_c2_kernel<1>);
      dim3 grid = pick_grid(B, V, res);
      uncl_fwd_c2_kernel<1><<<grid, kThreads, 0, st>>>(s, t, B, V, beta, inv_count, stash, ws.ticket, ws.partials,
                                                        sum_out, loss_out);
    }
  } else {
    static const int res = resident_ctas(uncl_fwd_generic_kernel);
    dim3 grid = pick_grid(B, V, res);
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
      static const int res = resident_ctas(uncl_bwd_c2_kernel<4>);
      uncl_bwd_c2_kernel<4><<<pick_grid(B, V / 4, res), kThreads, 0, st>>>(stash, B, V, (float)inv_count, grad_out, grad_s);
    } else {
      static const int res = resident_ctas(uncl_bwd_c2_kernel<1>);
      uncl_bwd_c2_kernel<1><<<pick_grid(B, V, res), kThreads, 0, st>>>(stash, B, V, (float)inv_count, grad_out, grad_s);
    }
  } else {
    DYCON_REQUIRE(s && t, DYCON_ERR_ARG, "UnCL bwd: C != 2 recomputes from s/t (NULL given)");
    static const int res = resident_ctas(uncl_bwd_generic_kernel);
    uncl_bwd_generic_kernel<<<pick_grid(B, V, res), kThreads, 0, st>>>(s, t, B, C, V, beta, (float)inv_count, grad_out,
                                                                  grad_s);
  }
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // extern "C"
