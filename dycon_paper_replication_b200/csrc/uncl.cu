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
#include "exchange.cuh"
#include "tc_common.cuh"

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
// Two-class algebra, in log2 units (entropies H2 = H / ln 2: the ln 2 factors are applied once per voxel or
// once per thread instead of once per logarithm).  Per distribution, with a = |x1 - x0|, e = exp(-a):
//   p_hi = 1/(1+e), p_lo = e p_hi, y = eps (1+e), x = e + y,
//   log2(p_hi + eps) = yh - l1,  yh = log2(e) (y - y^2/2)   (log1p(y) exactly in fp32, y <= 2e-6)
//   log2(p_lo + eps) = lx - l1,  lx = log2(x),  l1 = log2(1+e)
//   H2 = l1 - p_hi yh - p_lo lx                               (p_hi + p_lo = 1)
struct TwoClass {
  float p_hi, p_lo, e, x, y;
  float yh, lx;       // log2 terms (see above)
  float H2;
  bool hi_is_1;
};
// Stage 1 (before the reciprocal): e, 1 + e and x.  Stage 2 gets 1/(1 + e) from the caller, which takes the
// reciprocals of several quantities with ONE MUFU.RCP of their product (MUFU issues 16 lanes/clk/SM and is,
// with the issue slots, what bounds this kernel).
__device__ __forceinline__ void two_class_begin(float x0, float x1, TwoClass& o, float& ope) {
  const float d = x1 - x0, a = fabsf(d);
  o.hi_is_1 = d >= 0.f;
  o.e = ex2_approx(-a * kLog2e) + (a - a);          // (a - a): +-inf logits become NaN like torch.softmax
  ope = 1.f + o.e;
  o.y = kEps * ope;
  o.x = o.e + o.y;
}
__device__ __forceinline__ void two_class_finish(TwoClass& o, float ope, float r_ope) {
  o.p_hi = r_ope;
  o.p_lo = o.e * r_ope;
  const float l1 = lg2_approx(ope);
  o.lx = lg2_approx(o.x);
  o.yh = (kLog2e * o.y) * fmaf(-0.5f, o.y, 1.f);
  o.H2 = fmaf(-o.p_lo, o.lx, fmaf(-o.p_hi, o.yh, l1));
}

// Returns the two parts of L_v = q W + beta ln2 (Hs2 + Ht2) separately (the caller accumulates them apart and
// applies beta ln2 once) and the unit gradient dL_v/ds1 (SURVEY.md section 0.1 closed form):
//   u = ps0 ps1 (g1 - g0),  g_c = 2 (ps_c - pt_c) W - dL/dHs (ln(ps_c+eps) + ps_c/(ps_c+eps))
// Two classes: ps0 - pt0 = -(ps1 - pt1) = -d, so q = 2 d^2 and 2 (d1 - d0) W = 4 d W.
__device__ __forceinline__ void voxel2(float s0, float s1, float t0, float t1, float beta, float& qw, float& h2,
                                       float& u) {
  TwoClass S, T;
  float as, at;                                     // 1 + e of the student / teacher
  two_class_begin(s0, s1, S, as);
  two_class_begin(t0, t1, T, at);
  // 1/as, 1/at, 1/S.x from one reciprocal (as, at in [1, 2], S.x in [1e-6, 2]: the product cannot over/underflow)
  const float p12 = as * at;
  const float r = rcp_approx(p12 * S.x);
  const float r_x = r * p12, r12 = r * S.x;
  two_class_finish(S, as, r12 * at);
  two_class_finish(T, at, r12 * as);
  // exp(beta H) = 2^(beta log2(e) ln2 H2) = 2^(beta H2)
  const float es = ex2_approx(beta * S.H2), et = ex2_approx(beta * T.H2);
  const float w = rcp_approx(es + et);
  const float d = (S.hi_is_1 ? S.p_hi : S.p_lo) - (T.hi_is_1 ? T.p_hi : T.p_lo);   // ps1 - pt1
  const float dw = d * w;
  qw = 2.f * d * dw;
  h2 = S.H2 + T.H2;
  // dL/dHs = beta (1 - q es W^2)
  const float dl_dh = fmaf(-beta * (2.f * dw * dw), es, beta);
  // (ln_hi - ln_lo) + (r_hi - r_lo),  r_hi = p_hi/(p_hi+eps) = 1 - y(1-y),  r_lo = e / (e + y)
  const float r_hi = fmaf(-S.y, 1.f - S.y, 1.f), r_lo = S.e * r_x;
  const float hi_minus_lo = fmaf(kLn2, S.yh - S.lx, r_hi - r_lo);
  const float g1_minus_g0 = fmaf(-dl_dh, S.hi_is_1 ? hi_minus_lo : -hi_minus_lo, 4.f * dw);
  u = S.p_hi * S.p_lo * g1_minus_g0;
}

// ---- the same for TWO voxels in packed fp32 (FFMA2 / FMUL2 / FADD2: one issue slot for both; sm_100) --------------
// The C == 2 forward sits between the XU pipe (10 MUFU per voxel) and the issue slots (~80 instructions per voxel): the
// packed form halves the fp32 share of the latter.  Same operations in the same order as voxel2 -- every packed
// instruction is the IEEE operation on both halves -- so the results are bit-identical to the scalar form.
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 sel2(bool cx, bool cy, float2 a, float2 b) { return make_float2(cx ? a.x : b.x, cy ? a.y : b.y); }
struct TwoClass2 {
  float2 p_hi, p_lo, e, x, y, yh, lx, H2;
  bool hi_x, hi_y;      // hi_is_1 of the two voxels
};
__device__ __forceinline__ void two_class_begin2(float2 x0, float2 x1, TwoClass2& o, float2& ope) {
  const float2 d = __fadd2_rn(x1, neg2(x0));
  const float2 a = make_float2(fabsf(d.x), fabsf(d.y));
  o.hi_x = d.x >= 0.f;
  o.hi_y = d.y >= 0.f;
  const float2 t = __fmul2_rn(neg2(a), f2(kLog2e));
  o.e = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), __fadd2_rn(a, neg2(a)));   // (a - a): +-inf logits -> NaN
  ope = __fadd2_rn(f2(1.f), o.e);
  o.y = __fmul2_rn(f2(kEps), ope);
  o.x = __fadd2_rn(o.e, o.y);
}
__device__ __forceinline__ void two_class_finish2(TwoClass2& o, float2 ope, float2 r_ope) {
  o.p_hi = r_ope;
  o.p_lo = __fmul2_rn(o.e, r_ope);
  const float2 l1 = make_float2(lg2_approx(ope.x), lg2_approx(ope.y));
  o.lx = make_float2(lg2_approx(o.x.x), lg2_approx(o.x.y));
  o.yh = __fmul2_rn(__fmul2_rn(f2(kLog2e), o.y), __ffma2_rn(f2(-0.5f), o.y, f2(1.f)));
  o.H2 = __ffma2_rn(neg2(o.p_lo), o.lx, __ffma2_rn(neg2(o.p_hi), o.yh, l1));
}
__device__ __forceinline__ void voxel2x2(float2 s0, float2 s1, float2 t0, float2 t1, float beta, float2& qw, float2& h2,
                                         float2& u) {
  TwoClass2 S, T;
  float2 as, at;
  two_class_begin2(s0, s1, S, as);
  two_class_begin2(t0, t1, T, at);
  const float2 p12 = __fmul2_rn(as, at);
  const float2 pr = __fmul2_rn(p12, S.x);
  const float2 r = make_float2(rcp_approx(pr.x), rcp_approx(pr.y));
  const float2 r_x = __fmul2_rn(r, p12), r12 = __fmul2_rn(r, S.x);
  two_class_finish2(S, as, __fmul2_rn(r12, at));
  two_class_finish2(T, at, __fmul2_rn(r12, as));
  const float2 bs = __fmul2_rn(f2(beta), S.H2), bt = __fmul2_rn(f2(beta), T.H2);
  const float2 es = make_float2(ex2_approx(bs.x), ex2_approx(bs.y)), et = make_float2(ex2_approx(bt.x), ex2_approx(bt.y));
  const float2 ws = __fadd2_rn(es, et);
  const float2 w = make_float2(rcp_approx(ws.x), rcp_approx(ws.y));
  const float2 d = __fadd2_rn(sel2(S.hi_x, S.hi_y, S.p_hi, S.p_lo), neg2(sel2(T.hi_x, T.hi_y, T.p_hi, T.p_lo)));   // ps1 - pt1
  const float2 dw = __fmul2_rn(d, w);
  qw = __fmul2_rn(__fmul2_rn(f2(2.f), d), dw);
  h2 = __fadd2_rn(S.H2, T.H2);
  const float2 dl_dh = __ffma2_rn(__fmul2_rn(f2(-beta), __fmul2_rn(__fmul2_rn(f2(2.f), dw), dw)), es, f2(beta));
  const float2 r_hi = __ffma2_rn(neg2(S.y), __fadd2_rn(f2(1.f), neg2(S.y)), f2(1.f)), r_lo = __fmul2_rn(S.e, r_x);
  const float2 hml = __ffma2_rn(f2(kLn2), __fadd2_rn(S.yh, neg2(S.lx)), __fadd2_rn(r_hi, neg2(r_lo)));
  const float2 g = __ffma2_rn(neg2(dl_dh), sel2(S.hi_x, S.hi_y, hml, neg2(hml)), __fmul2_rn(f2(4.f), dw));
  u = __fmul2_rn(__fmul2_rn(S.p_hi, S.p_lo), g);
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
// Sharded batch (x.world > 1): the same warp then runs the partial-sum exchange with the peers (exchange.cuh), so
// sum_out / loss_out hold the GLOBAL sum / loss and the step has no extra launch for it.
__device__ __forceinline__ void uncl_finish(float acc, unsigned int* ticket, double* partials, double inv_count,
                                            double* __restrict__ sum_out, float* __restrict__ loss_out,
                                            const ExchangeCtx& x) {
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
  if (lane == 0) *ticket = 0u;
  if (x.world > 1) tot = exchange_warp(x, tot, 1);
  if (lane == 0) {
    *sum_out = tot;
    if (loss_out) *loss_out = (float)(tot * inv_count);
  }
}

// Grid: (blocks per sample, sample lanes).  The loads of iteration k+1 are issued before the arithmetic of
// iteration k (register double buffer): with ~80 instructions and 12 MUFU per voxel the kernel sits between
// the XU pipe and HBM, and a warp that waits for its loads at the top of every iteration leaves both idle.
template <int kVec>
__global__ void __launch_bounds__(kThreads, 3)
uncl_fwd_c2_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t B, int64_t V, float beta,
                   double inv_count, float* __restrict__ stash, unsigned int* ticket, double* partials,
                   double* __restrict__ sum_out, float* __restrict__ loss_out, const __grid_constant__ ExchangeCtx xc) {
  using P = Pack<kVec>;
  using vec_t = typename P::type;
  const int64_t nvec = V / kVec;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  float acc = 0.f, acc_h = 0.f;     // sum of q W, sum of Hs2 + Ht2 (log2 units)
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
      float qsum = 0.f, hsum = 0.f;
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        float qw, h2, u;
        voxel2(elem(a0, k), elem(a1, k), elem(b0, k), elem(b1, k), beta, qw, h2, u);
        qsum += qw;
        hsum += h2;
        set_elem(uo, k, u);
      }
      *reinterpret_cast<vec_t*>(st + i * kVec) = uo;  // default policy: re-read by the backward from L2
      acc += qsum;
      acc_h += hsum;
      if (!more) break;
      a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
      i = nx;
    }
  }
  uncl_finish(fmaf(beta * kLn2, acc_h, acc), ticket, partials, inv_count, sum_out, loss_out, xc);
}

// ---- C == 2, 16-byte aligned: cp.async pipelined forward ------------------------------------------
// A persistent CTA walks chunks of 1024 voxels.  Every thread runs its own 3-stage cp.async pipeline: it
// copies the four float4 it will need two chunks from now (s0, s1, t0, t1) into its private slots of a
// shared-memory ring, waits for the oldest group and computes from shared memory.  Nothing is shared between
// threads, so there are no barriers; 1024 threads x 2 groups x 64 B = 131 KB of loads are in flight per SM no
// matter how the compiler schedules the arithmetic.  (The register-prefetch version had its loads sunk to the
// end of the loop body by ptxas and waited for them a third of the time; a TMA-bulk + mbarrier version was
// limited to ~12 B/clk/SM from HBM by the TMA unit's outstanding-request budget -- profiles/.)
constexpr int kPipeStages = 3;
constexpr int kPipeChunk = 1024;                 // voxels per chunk = 256 threads x float4
constexpr size_t kPipeSmemBytes = (size_t)kPipeStages * 4 * kThreads * sizeof(float4);   // 48 KB: four CTAs per SM

// The logits are read exactly once: fetched with an L2 evict-first policy they do not push out what the step still
// needs from the 126 MB L2 -- FeCL's pair matrices (51 MB, written by the forward, read by the backward GEMM) and the
// stash of this kernel.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16_stream(void* dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

__global__ void __launch_bounds__(kThreads, 4)
uncl_fwd_c2_pipe_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t V, int64_t chunks_per_sample,
                        int64_t total_chunks, float beta, double inv_count, float* __restrict__ stash,
                        unsigned int* ticket, double* partials, double* __restrict__ sum_out,
                        float* __restrict__ loss_out, const __grid_constant__ ExchangeCtx xc) {
  extern __shared__ __align__(16) float4 ring[];          // [kPipeStages][4 streams][kThreads]
  const int tid = threadIdx.x;
  const int step = gridDim.x, cps = (int)chunks_per_sample;
  // (sample, chunk-in-sample) positions advance incrementally: no 64-bit division in the loop
  struct Pos { int b, k; };
  auto advance = [&](Pos& p, int n) {
    p.k += n;
    while (p.k >= cps) { p.k -= cps; ++p.b; }
  };
  const int nb = (int)(total_chunks / chunks_per_sample);
  const uint64_t pol = l2_evict_first_policy();

  auto issue = [&](const Pos& p, int st) {                // this thread's four 16-byte copies of chunk p
    if (p.b < nb) {
      const int64_t v = (int64_t)p.k * kPipeChunk + tid * 4;
      if (v < V) {
        const float* s0 = s + (2 * (int64_t)p.b) * V + v;
        const float* t0 = t + (2 * (int64_t)p.b) * V + v;
        float4* dst = ring + (size_t)st * 4 * kThreads + tid;
        cp_async16_stream(dst, s0, pol);
        cp_async16_stream(dst + kThreads, s0 + V, pol);
        cp_async16_stream(dst + 2 * kThreads, t0, pol);
        cp_async16_stream(dst + 3 * kThreads, t0 + V, pol);
      }
    }
    cp_async_commit();                                    // always: keeps the group count uniform
  };

  Pos cur{0, 0}, nxt{0, 0};
  advance(cur, blockIdx.x);
  nxt = cur;
#pragma unroll
  for (int k = 0; k < kPipeStages - 1; ++k) {
    issue(nxt, k);
    advance(nxt, step);
  }
  float acc = 0.f, acc_h = 0.f;     // sum of q W, sum of Hs2 + Ht2 (log2 units)
  for (int it = 0; cur.b < nb; ++it) {
    issue(nxt, (it + kPipeStages - 1) % kPipeStages);
    advance(nxt, step);
    cp_async_wait<kPipeStages - 1>();                     // the current chunk has landed (in this thread's own slots)
    const int64_t v = (int64_t)cur.k * kPipeChunk + tid * 4;
    if (v < V) {
      const float4* src = ring + (size_t)(it % kPipeStages) * 4 * kThreads + tid;
      const float4 a0 = src[0], a1 = src[kThreads], b0 = src[2 * kThreads], b1 = src[3 * kThreads];
      float4 uo;
      float qsum, hsum;
      {
        float2 qa, ha, ua, qb, hb, ub;
        voxel2x2(make_float2(a0.x, a0.y), make_float2(a1.x, a1.y), make_float2(b0.x, b0.y), make_float2(b1.x, b1.y), beta, qa, ha, ua);
        voxel2x2(make_float2(a0.z, a0.w), make_float2(a1.z, a1.w), make_float2(b0.z, b0.w), make_float2(b1.z, b1.w), beta, qb, hb, ub);
        // (the scalar form's order of the four additions)
        qsum = (((0.f + qa.x) + qa.y) + qb.x) + qb.y;
        hsum = (((0.f + ha.x) + ha.y) + hb.x) + hb.y;
        uo = make_float4(ua.x, ua.y, ub.x, ub.y);
      }
      *reinterpret_cast<float4*>(stash + (int64_t)cur.b * V + v) = uo;   // default policy: re-read by the backward from L2
      acc += qsum;
      acc_h += hsum;
    }
    advance(cur, step);
  }
  cp_async_wait<0>();
  uncl_finish(fmaf(beta * kLn2, acc_h, acc), ticket, partials, inv_count, sum_out, loss_out, xc);
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
      const vec_t u = __ldcs(reinterpret_cast<const vec_t*>(st + i * kVec));       // last use of the stash: evict first
      vec_t u2 = u;
      if (two) u2 = __ldcs(reinterpret_cast<const vec_t*>(st + i2 * kVec));
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

// Occupancy of the pipelined forward (per device: a process may drive several GPUs).
int uncl_pipe_ctas_per_sm(int* out) {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  DYCON_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    int n = 0;
    DYCON_CUDA(cudaFuncSetAttribute(uncl_fwd_c2_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmemBytes));
    DYCON_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, uncl_fwd_c2_pipe_kernel, kThreads, kPipeSmemBytes));
    DYCON_REQUIRE(n > 0, DYCON_ERR_DEVICE, "UnCL fwd: cannot place the pipelined kernel (%zu B of shared memory)", kPipeSmemBytes);
    cached = n;
    cached_dev = dev;
  }
  *out = cached;
  return DYCON_OK;
}

int check_common(int64_t B, int C, int64_t V) {
  DYCON_REQUIRE(B > 0 && V > 0, DYCON_ERR_ARG, "UnCL: B=%lld V=%lld must be positive", (long long)B, (long long)V);
  DYCON_REQUIRE(C >= 2 && C <= 1024, DYCON_ERR_UNSUPPORTED, "UnCL: C=%d outside [2, 1024]", C);
  return DYCON_OK;
}

}  // namespace
}  // namespace dycon

using namespace dycon;

extern "C" size_t dycon_uncl_workspace_bytes(void) { return 16 + sizeof(double) * kMaxPartials; }

namespace {

// x == nullptr or x->world == 1: the plain forward.  Otherwise the C == 2 kernels run the exchange of the partial
// sum in their own tail; the generic-C kernel is followed by the stand-alone exchange launch.
int uncl_fwd_impl(const float* s, const float* t, int64_t B, int C, int64_t V, float beta, double inv_count,
                  float* stash, double* sum_out, float* loss_out, void* workspace, size_t workspace_bytes,
                  const ExchangeCtx& xc, cudaStream_t st) {
  if (int rc = check_common(B, C, V)) return rc;
  DYCON_REQUIRE(s && t && sum_out && workspace, DYCON_ERR_ARG, "UnCL fwd: NULL s/t/sum_out/workspace");
  DYCON_REQUIRE(aligned(s, 4) && aligned(t, 4) && aligned(workspace, 16) && aligned(sum_out, 8), DYCON_ERR_ARG,
                "UnCL fwd: misaligned pointer");
  DYCON_REQUIRE(workspace_bytes >= dycon_uncl_workspace_bytes(), DYCON_ERR_WORKSPACE,
                "UnCL fwd: workspace %zu < %zu bytes", workspace_bytes, dycon_uncl_workspace_bytes());
  ReduceWorkspace ws = carve_reduce_workspace(workspace);
  if (C == 2) {
    DYCON_REQUIRE(stash && aligned(stash, 4), DYCON_ERR_ARG, "UnCL fwd: C == 2 needs a stash of B*V floats");
    const bool vec = (V % 4 == 0) && aligned(s, 16) && aligned(t, 16) && aligned(stash, 16);
    if (vec) {
      int per_sm = 0;
      if (int rc = uncl_pipe_ctas_per_sm(&per_sm)) return rc;
      const int64_t cps = (V + kPipeChunk - 1) / kPipeChunk, total = cps * B;
      int64_t grid = (int64_t)per_sm * sm_count();
      if (grid > total) grid = total;
      if (grid > kMaxPartials) grid = kMaxPartials;
      uncl_fwd_c2_pipe_kernel<<<(unsigned)grid, kThreads, kPipeSmemBytes, st>>>(
          s, t, V, cps, total, beta, inv_count, stash, ws.ticket, ws.partials, sum_out, loss_out, xc);
    } else {
      static const int res = resident_ctas(uncl_fwd_c2_kernel<1>);
      dim3 grid = pick_grid(B, V, res);
      uncl_fwd_c2_kernel<1><<<grid, kThreads, 0, st>>>(s, t, B, V, beta, inv_count, stash, ws.ticket, ws.partials,
                                                        sum_out, loss_out, xc);
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

}  // namespace

extern "C" {

int dycon_uncl_fwd(const float* s, const float* t, int64_t B, int C, int64_t V, float beta, double inv_count,
                   float* stash, double* sum_out, float* loss_out, void* workspace, size_t workspace_bytes,
                   dycon_stream_t stream) {
  ExchangeCtx xc;
  make_exchange_ctx(&xc, nullptr, 0, 1, nullptr, DYCON_CHANNEL_UNCL, 0.0);
  return uncl_fwd_impl(s, t, B, C, V, beta, inv_count, stash, sum_out, loss_out, workspace, workspace_bytes, xc,
                       as_stream(stream));
}

int dycon_uncl_fwd_sharded(const float* s, const float* t, int64_t B, int C, int64_t V, float beta, double inv_count,
                           float* stash, double* sum_out, float* loss_out, void* workspace, size_t workspace_bytes,
                           void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                           double timeout_s, dycon_stream_t stream) {
  ExchangeCtx xc;
  if (int rc = make_exchange_ctx(&xc, peer_inboxes, rank, world, seq_counters, DYCON_CHANNEL_UNCL, timeout_s)) return rc;
  if (C == 2 || xc.world == 1)
    return uncl_fwd_impl(s, t, B, C, V, beta, inv_count, stash, sum_out, loss_out, workspace, workspace_bytes, xc,
                         as_stream(stream));
  ExchangeCtx none;
  make_exchange_ctx(&none, nullptr, 0, 1, nullptr, DYCON_CHANNEL_UNCL, 0.0);
  if (int rc = uncl_fwd_impl(s, t, B, C, V, beta, inv_count, stash, sum_out, nullptr, workspace, workspace_bytes, none,
                             as_stream(stream)))
    return rc;
  return dycon_exchange_sums(sum_out, 1, sum_out, peer_inboxes, rank, world, seq_counters, DYCON_EXCHANGE_UNCL, inv_count,
                             0.0, loss_out, timeout_s, stream);
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
