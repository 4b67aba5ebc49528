// K2 (16-bit tensor-core modes) -- FeCL forward / backward on tcgen05 + TMEM + TMA (sm_100a).
// Operands are bf16 or fp16 (template parameter; same kind::f16 MMA, same rate, fp32 accumulation).
// fp16's 10-bit mantissa is what keeps the gradient inside the 2e-3 budget on every input.
// Reference: code/utils/dycon_losses.py:150-235; per-pair algebra in fecl_math.cuh / SURVEY.md 0.2.
//
// Data layout in HBM (the `state` buffer, written by the forward, read by the backward):
//   hdr    : 1 KiB; hdr[0] = power-of-two scale applied to H / Gc before the 16-bit conversion
//   Fb, Tb : 16-bit [B][Npad][Dpad], row-major (K-major for the MMAs), Npad = ceil128(N),
//            Dpad = ceil64(D); padding rows / columns are zero.
//   stats  : planes of B*N floats: m (column == row max), n (negative sum), A, kappa, P (positive count),
//            then kMaxSplits planes of P1's per-split partial n.
// No (N,N) tensor is ever written: every similarity tile lives in TMEM and is consumed in place.
//
// Kernels (all: TMA producer warp, MMA issuer warps, 16 epilogue warps in two teams; mbarrier pipelines):
//   pack16_kernel                 fp32 (B,N,D) with any strides -> Fb, Tb; zeroes the statistics
//   fecl_tc_sweep_kernel<0,..,2>  P0: S = F_I F_J^T sub-tiles -> row max m_i            (256-row CTAs)
//   fecl_tc_sweep_kernel<1,..,2>  P1: S sub-tiles -> partial n_i, positive counts P_i   (256-row CTAs)
//   fecl_tc_sweep_kernel<2,..,1>  P2: S | CS sub-tiles -> kappa_i, row loss, A_i, cross sum / count, the loss
//   fecl_tc_bwd_kernel            per 32-column sub-tile: S | CS -> H = G + G^T, Gc (16-bit, written to smem in
//                                 the UMMA K-major swizzle) -> dF_I += H F_J + Gc T_J with F_J / T_J read as
//                                 MN-major operands from the very stage that produced S | CS.
// The four forward kernels are chained with programmatic dependent launch.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "exchange.cuh"
#include "fecl_internal.h"
#include "tc_common.cuh"

namespace dycon {
namespace {

using namespace tc;

// ---- device-clock timeline of ONE CTA (measurement aid, compiled in only with -DDYCON_TIMELINE) --------------
// Lane 0 of the producer warp, of the MMA issuer warps and of one warp per epilogue team stores clock64() at the
// events of every sub-tile into a buffer that tools/timeline.py reads back through dycon_debug_timeline().
#ifdef DYCON_TIMELINE
__device__ unsigned long long g_timeline[2][4][64][8];      // [kernel: 0 sweep (P2), 1 backward][role][sub-tile][event]
#define DYCON_TL(kern, on, role, t, ev)                                                  \
  do {                                                                                   \
    if ((on) && (t) < 64) g_timeline[kern][role][t][ev] = (unsigned long long)clock64(); \
  } while (0)
__device__ __forceinline__ void g_tl_cls(int t, int cls) { if (t < 64) g_timeline[1][0][t][7] = (unsigned long long)cls; }
// Spans of EVERY CTA of the four kernels of the default path on the GPU-wide nanosecond timer (%globaltimer): entry,
// main loop reached, main loop done, exit, SM id, work units -- shows launch skew, programmatic-launch overlap and tails.
__device__ unsigned long long g_span[4][512][6];            // [0 pack, 1 sweep, 2 row kernel, 3 backward GEMM][CTA][event]
__device__ __forceinline__ unsigned long long g_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int g_smid() { unsigned int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned int g_cta() { return (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x; }
#define DYCON_SPAN(kern, on, ev)                                              \
  do {                                                                        \
    if ((on) && g_cta() < 512) g_span[kern][g_cta()][ev] = g_now();           \
  } while (0)
#define DYCON_SPAN_V(kern, on, ev, val)                                       \
  do {                                                                        \
    if ((on) && g_cta() < 512) g_span[kern][g_cta()][ev] = (unsigned long long)(val); \
  } while (0)
#else
#define DYCON_TL(kern, on, role, t, ev) do {} while (0)
#define DYCON_SPAN(kern, on, ev) do {} while (0)
#define DYCON_SPAN_V(kern, on, ev, val) do {} while (0)
__device__ __forceinline__ void g_tl_cls(int, int) {}
#endif

constexpr int kTM = 128;               // rows per CTA (UMMA M)
constexpr uint32_t kChunk128 = 128 * 128;   // bytes of a [128 rows][64 bf16] swizzle chunk
constexpr uint32_t kChunk64 = 64 * 128;     // bytes of a [ 64 rows][64 bf16] swizzle chunk
constexpr uint32_t kChunk32 = 32 * 128;     // bytes of a [ 32 rows][64 bf16] swizzle chunk
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Tests as 1.0f / 0.0f masks (one FSET each, no predicate): with dozens of pairs per lane in flight ptxas otherwise parks
// the predicates in bit masks, two LOP3 per pair to build and two more to read.  NaN labels compare unequal to everything
// (mask_ne is the unordered test), a NaN similarity is never above a threshold.
__device__ __forceinline__ float mask_eq(float a, float b) { float r; asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mask_ne(float a, float b) { float r; asm("set.neu.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mask_gt(float a, float b) { float r; asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
// Focal variants are compile-time (a runtime `if (focal)` / gamma test per pair costs branches in the
// hottest loop): 0 = no focal weight, 1 = gamma == 2 (the scripts' value), 2 = any gamma.
enum { kNoFocal = 0, kFocalG2 = 1, kFocalAny = 2 };

// Positive pair, with t = log2(e_ij) (the exponent argument), e = 2^t and the row's negative sum n:
//   T = e + n, d = e/T, log2 d = t - log2 T.
// fwd returns phi2 with phi(d) = -ln2 * phi2 (the -ln2 is applied once per row) and the A-summand
// phi'(d) d / T;  bwd returns phi'(d) d (1-d).
// (1-d)^(gamma-1) for any gamma.  A row without negatives (n_i = 0: a single-class sample) has d = 1 (or 1 + ulp),
// where the reference evaluates (1-1)^gamma = 0: clamp 1-d to >= 0 and take the logarithm of a positive floor, so
// that gamma > 1 gives 0, gamma == 1 gives 1 (torch.pow(0., 0.) = 1) and nothing turns into NaN.
__device__ __forceinline__ float pow_gm1(float& omd, float gamma) {
  omd = fmaxf(omd, 0.f);
  return ex2_approx((gamma - 1.f) * lg2_approx(fmaxf(omd, 1e-30f)));
}

template <int kFocal>
__device__ __forceinline__ void pos_fwd(float t, float e, float n, float gamma, float& phi2, float& a_term) {
  const float T = e + n;
  const float rT = rcp_approx(T);
  const float tL = t - lg2_approx(T);
  if (kFocal == kNoFocal) {
    phi2 = tL;
    a_term = -rT;
  } else {
    const float d = e * rT;
    float omd = 1.f - d;
    const float w1 = kFocal == kFocalG2 ? omd : pow_gm1(omd, gamma);
    const float w = w1 * omd;
    phi2 = tL * w;
    a_term = fmaf(gamma * kLn2 * w1 * d, tL, -w) * rT;
  }
}
template <int kFocal>
__device__ __forceinline__ float pos_bwd(float t, float e, float n, float gamma) {
  const float T = e + n;
  const float d = e * rcp_approx(T);
  float omd = 1.f - d;
  if (kFocal == kNoFocal) return -omd;
  const float tL = t - lg2_approx(T);
  const float w1 = kFocal == kFocalG2 ? omd : pow_gm1(omd, gamma);
  return w1 * omd * fmaf(gamma * kLn2 * d, tL, -omd);
}

// The loss sweep of the stored-pairs build (kStore): the forward terms of pos_fwd AND the backward term
// px = phi'(d) d (1 - d) of the same pair (it shares every factor with a_term), and -- with a teacher -- the
// reciprocal of the cross term's argument out of the SAME MUFU.RCP:  fac = 1 - cs on a hard negative, else 1, so
// that R = 1/(T fac) yields 1/T = R fac and 1/(1 - cs) = R T.
template <int kFocal>
__device__ __forceinline__ void pos_fwd_x(float t, float e, float n, float gamma, float fac, float& phi2, float& a_term,
                                          float& px, float& rom) {
  const float T = e + n;
  const float R = rcp_approx(T * fac);
  const float rT = R * fac;
  rom = R * T;
  const float tL = t - lg2_approx(T);
  if (kFocal == kNoFocal) {
    phi2 = tL;
    a_term = -rT;
    px = -(n * rT);                       // -(1 - d)
  } else {
    const float d = e * rT;
    float omd = 1.f - d;
    const float w1 = kFocal == kFocalG2 ? omd : pow_gm1(omd, gamma);
    const float w = w1 * omd;
    const float y = fmaf(gamma * kLn2 * d, tL, -omd);
    phi2 = tL * w;
    a_term = w1 * y * rT;
    px = w * y;
  }
}

// ---- rows sorted by label ----------------------------------------------------------------------------
// FeCL is equivariant under a permutation of the rows of a sample, and every pair term depends on the labels only
// through same / different.  The forward therefore packs the rows of each sample SORTED by label (stable): a
// (row block, column sub-tile) is then all-positive, all-negative or -- only where a class boundary crosses it --
// mixed, which the sweeps and the backward turn into three specialised epilogue bodies and into sub-tiles that
// are skipped outright.  fecl_rank_kernel computes, per sample, the sorted position of every row (a counting
// rank: every row is compared with every key of the sample out of shared memory) and, by sorted position, the original row
// (`perm`, used to scatter the gradient back), the label, the row weight and the class bounds
// [cls_lo, cls_hi) = sorted positions that carry the same label.  NaN labels compare unequal to everything,
// themselves included (dycon_losses.py:172): each NaN row is a class of its own.
enum { kClsMixed = 0, kClsSame = 1, kClsDiff = 2 };

struct RankParams {
  const float* labels;       // (B, N), the caller's order
  const float* row_weight;   // (B, N) or nullptr
  int N;
  int* rank;                 // rank[b*N + n] = sorted position of row n
  int* perm;                 // perm[b*N + pos] = n
  int* cls_lo;               // by sorted position
  int* cls_hi;
  float* ys;                 // labels by sorted position
  float* rws;                // row weights by sorted position (written only with a row_weight)
  int pdl;
};

__device__ __forceinline__ uint32_t label_key(float y) {
  const uint32_t u = __float_as_uint(y + 0.f);            // -0 -> +0: they are the same label
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // unsigned order == float order
}

// Grid (ceil(N / 32), B), 256 threads: warp w ranks the four rows 32 blockIdx.x + 4 w + {0..3}; its lanes stride
// over the keys of the sample (shared memory, conflict-free) and every loaded key is compared against the four row
// keys.  An iteration lies entirely in front of, entirely behind or (at most two) across the four rows, which
// decides warp-uniformly whether an equal key counts as "in front" (stable order).  Dynamic shared memory:
// ceil4(N) keys.  (A first version with one THREAD per row ran 15 us at N = 1728: 216 long, lonely warps.)
__global__ void __launch_bounds__(256) fecl_rank_kernel(const RankParams p) {
  extern __shared__ __align__(16) uint32_t rk_keys[];
  const int tid = threadIdx.x, b = blockIdx.y, N = p.N, N4 = (N + 3) & ~3;
  const int warp = tid >> 5, lane = tid & 31;
  if (p.pdl) pdl_trigger();
  const float* lab = p.labels + (size_t)b * N;
  for (int j = tid; j < N4; j += 256) rk_keys[j] = j < N ? label_key(__ldg(lab + j)) : 0xffffffffu;
  __syncthreads();
  const int i0 = blockIdx.x * 32 + warp * 4;
  if (i0 >= N) return;
  uint32_t k[4];
  int lt[4] = {0, 0, 0, 0}, eqb[4] = {0, 0, 0, 0}, eqa[4] = {0, 0, 0, 0};   // keys < k; equal keys in front of / behind row i
#pragma unroll
  for (int r = 0; r < 4; ++r) k[r] = rk_keys[min(i0 + r, N - 1)];
  for (int base = 0; base < N4; base += 32) {
    const int j = base + lane;
    const uint32_t kj = j < N4 ? rk_keys[j] : 0xffffffffu;        // padding never sorts in front of a row
    if (base + 31 < i0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) { lt[r] += kj < k[r]; eqb[r] += kj == k[r]; }
    } else if (base > i0 + 3) {
#pragma unroll
      for (int r = 0; r < 4; ++r) { lt[r] += kj < k[r]; eqa[r] += kj == k[r]; }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const bool e = kj == k[r];
        lt[r] += kj < k[r];
        eqb[r] += e && j < i0 + r;
        eqa[r] += e && j >= i0 + r;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lt[r] += __shfl_xor_sync(0xffffffffu, lt[r], o);
      eqb[r] += __shfl_xor_sync(0xffffffffu, eqb[r], o);
      eqa[r] += __shfl_xor_sync(0xffffffffu, eqa[r], o);
    }
  }
  if (lane < 4 && i0 + lane < N) {
    const int i = i0 + lane;
    const int l = lane == 0 ? lt[0] : lane == 1 ? lt[1] : lane == 2 ? lt[2] : lt[3];
    const int eb = lane == 0 ? eqb[0] : lane == 1 ? eqb[1] : lane == 2 ? eqb[2] : eqb[3];
    const int ea = lane == 0 ? eqa[0] : lane == 1 ? eqa[1] : lane == 2 ? eqa[2] : eqa[3];
    // (the padding keys 0xffffffff only ever equal a NaN key, whose bounds are set below)
    const float y = __ldg(lab + i);
    const int pos = l + eb;
    const size_t o = (size_t)b * N;
    p.rank[o + i] = pos;
    p.perm[o + pos] = i;
    p.ys[o + pos] = y;
    if (p.row_weight) p.rws[o + pos] = __ldg(p.row_weight + o + i);
    const bool nan = y != y;
    p.cls_lo[o + pos] = nan ? pos : l;
    p.cls_hi[o + pos] = nan ? pos + 1 : l + eb + ea;
  }
}

// The label range of the rows of a CTA in sorted positions: their labels occupy [lo, hi); `uniform`: one label.
struct RowClass { int lo, hi, uniform; };
__device__ __forceinline__ RowClass load_row_class(const int* cls_lo, const int* cls_hi, size_t off, int i0, int rows, int N) {
  RowClass rc{0, 0x7fffffff, 0};                                // rows not sorted: every sub-tile is MIXED
  if (cls_lo) {
    const int last = min(i0 + rows, N) - 1;
    const int a = __ldg(cls_lo + off + i0), b = __ldg(cls_lo + off + last);
    rc.lo = a;
    rc.hi = __ldg(cls_hi + off + last);
    rc.uniform = a == b;
  }
  return rc;
}
// Sub-tiles (tc columns each) whose pairs are ALL positive: [ta, tb).  hi <= N, so none of them holds padding.
__device__ __forceinline__ void same_range(const RowClass& rc, int tc, int& ta, int& tb) {
  ta = tb = 0;
  if (rc.uniform) {
    ta = (rc.lo + tc - 1) / tc;
    tb = rc.hi / tc;
    if (tb < ta) tb = ta;
  }
}
__device__ __forceinline__ int tile_class(const RowClass& rc, int t, int tc, int N, int ta, int tb) {
  if (t >= ta && t < tb) return kClsSame;
  const int j0 = t * tc, j1 = min(j0 + tc, N);
  return (j1 <= rc.lo || j0 >= rc.hi) ? kClsDiff : kClsMixed;   // padded columns of a DIFF sub-tile have e = 0, cs = 0
}

// ---- pack: fp32 (B,N,D) with element strides -> bf16 [B][Npad][Dpad], zero padded ---------------
template <bool kBf16> struct Cvt;
template <> struct Cvt<true> {
  using type = __nv_bfloat16;
  static __device__ __forceinline__ type one(float x) { return __float2bfloat16(x); }
  static __device__ __forceinline__ uint32_t two(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct Cvt<false> {
  using type = __half;
  static __device__ __forceinline__ type one(float x) { return __float2half_rn(x); }
  static __device__ __forceinline__ uint32_t two(float a, float b) {   // saturate: fp16 overflows at 65504
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // a -> low half, b -> high half
    return r;
  }
};

// One launch packs the student and (if present) the teacher: fp32 (B,N,D) with element strides ->
// 16-bit [B][Npad][Dpad], zero padded; the CTAs of the first column tile also zero the per-row
// statistics (so no memset node is needed).  A CTA moves a 64 x 64 tile through shared memory: the
// caller's layout has n contiguous (strides (D*N, 1, N)), the kernel layout has d contiguous.
struct PackParams {
  const float* src[2];
  int64_t sb[2], sn[2], sd[2];
  void* dst[2];
  float* stats;        // kNumStats planes of B*N floats, zero-filled here
  int B, N, D, Npad, Dpad;
  int pdl;
  int merge;           // global-negatives mode: the B samples become ONE sample of B*N rows (row b*N + n, no
                       // per-sample padding; the caller zero-fills the tail rows of the state once)
  const int* rank;     // rows sorted by label: row n of sample b goes to row rank[b*N + n] (nullptr: stays at n).
                       // Written by fecl_rank_kernel, the predecessor in the stream: griddepcontrol.wait first.
  const float* scale[2];   // per-row factor (B*N floats, or nullptr): 1/max(|x|, eps) of F.normalize folded into the
                           // conversion, so that the caller need not materialise normalised embeddings (prep.cu)
  unsigned int* gc_flag;   // stored pairs: (Npad/128) x (Npad/64) tile flags per sample, cleared here (or nullptr)
};

template <bool kBf16>
__global__ void __launch_bounds__(256)
pack16_kernel(const PackParams p) {
  using T16 = typename Cvt<kBf16>::type;
  // [64 d][64 n] floats, the column index XORed with 4 (d / 8): a float4 of four consecutive n stays one aligned 16-byte
  // store, and the transposed reads -- a lane takes 8 consecutive d of one n, eight lanes cover the 64 d of that n --
  // spread over all 32 banks (bits 2..4 of the bank come from d / 8, bits 0..1 from n)
  __shared__ __align__(16) float tile[64 * 64];
  auto tix = [](int d, int n) { return d * 64 + (n ^ (((d >> 3) & 7) << 2)); };
  const int which = blockIdx.z / p.B, b = blockIdx.z - which * p.B;
  const int n0 = blockIdx.x * 64, d0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  DYCON_SPAN(0, tid == 0, 0);
  DYCON_SPAN_V(0, tid == 0, 4, g_smid());
  if (p.pdl) pdl_trigger();     // the first sweep may set up (barriers, TMEM) while this grid drains; it waits before reading
  // (ternaries, not p.x[which]: a dynamic index would copy the parameter arrays to local memory)
  const float* s = (which ? p.src[1] : p.src[0]) + (int64_t)b * (which ? p.sb[1] : p.sb[0]);
  const int64_t sn = which ? p.sn[1] : p.sn[0], sd = which ? p.sd[1] : p.sd[0];
  T16* dst = reinterpret_cast<T16*>(which ? p.dst[1] : p.dst[0]);
  const float* rscale = which ? p.scale[1] : p.scale[0];
  if (which == 0 && blockIdx.y == 0 && blockIdx.x == 0 && p.gc_flag) {
    const int nf = (p.Npad >> 7) * (p.Npad >> 6);
    for (int q = tid; q < nf; q += 256) p.gc_flag[(size_t)b * nf + q] = 0u;
  }
  if (which == 0 && blockIdx.y == 0) {
    for (int q = tid; q < kNumStats * 64; q += 256) {
      const int n = n0 + (q & 63);
      if (n < p.N) p.stats[(size_t)(q >> 6) * p.B * p.N + (size_t)b * p.N + n] = 0.f;
    }
  }
  if (sn == 1) {   // lanes along n for the read, along d for the write
    if (((p.N | sd | (which ? p.sb[1] : p.sb[0])) & 3) == 0 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
      // 16-byte loads, all four of a thread issued before the first use (the kernel is latency bound)
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = tid + 256 * k, d = d0 + (idx >> 4), n = n0 + (idx & 15) * 4;
        v[k] = (n < p.N && d < p.D) ? __ldg(reinterpret_cast<const float4*>(s + n + (int64_t)d * sd))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = tid + 256 * k;
        *reinterpret_cast<float4*>(&tile[tix(idx >> 4, (idx & 15) * 4)]) = v[k];
      }
    } else {
      const int tx = tid & 63, ty = tid >> 6;
#pragma unroll 4
      for (int k = ty; k < 64; k += 4) {
        const int n = n0 + tx, d = d0 + k;
        tile[tix(k, tx)] = (n < p.N && d < p.D) ? __ldg(s + n + (int64_t)d * sd) : 0.f;
      }
    }
    __syncthreads();
    if (p.rank && p.pdl) pdl_wait();       // the loads above overlap the rank kernel; its output is needed from here on
    // a thread writes 8 consecutive d of one row n as ONE 16-byte store; the eight lanes of a row fill its 128 bytes
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int u = tid + 256 * k, g = u & 7, r = u >> 3, n = n0 + r;
      if (p.merge ? n < p.N : n < p.Npad) {
        const int nd = (p.rank && n < p.N) ? __ldg(p.rank + (size_t)b * p.N + n) : n;
        const float sc = (rscale && n < p.N) ? __ldg(rscale + (size_t)b * p.N + n) : 1.f;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = tile[tix(8 * g + j, r)] * sc;
        *reinterpret_cast<uint4*>(dst + ((size_t)b * (p.merge ? p.N : p.Npad) + nd) * p.Dpad + d0 + 8 * g) =
            make_uint4(Cvt<kBf16>::two(f[0], f[1]), Cvt<kBf16>::two(f[2], f[3]), Cvt<kBf16>::two(f[4], f[5]),
                       Cvt<kBf16>::two(f[6], f[7]));
      }
    }
  } else {         // any other layout (d contiguous or generic): lanes along d for both
    if (p.rank && p.pdl) pdl_wait();
    const int tx = tid & 63, ty = tid >> 6;
    for (int r = ty; r < 64; r += 4) {
      const int n = n0 + r, d = d0 + tx;
      if (p.merge ? n < p.N : n < p.Npad) {
        const float sc = (rscale && n < p.N) ? __ldg(rscale + (size_t)b * p.N + n) : 1.f;
        const float v = (n < p.N && d < p.D) ? __ldg(s + (int64_t)n * sn + (int64_t)d * sd) * sc : 0.f;
        const int nd = (p.rank && n < p.N) ? __ldg(p.rank + (size_t)b * p.N + n) : n;
        dst[((size_t)b * (p.merge ? p.N : p.Npad) + nd) * p.Dpad + d] = Cvt<kBf16>::one(v);
      }
    }
  }
  DYCON_SPAN(0, tid == 0, 3);
}

// =================================================================================================
//  Forward sweeps
// =================================================================================================
// A CTA owns 128 rows of one sample (A tile resident in shared memory) and walks the sub-tiles of its
// column range.  A sub-tile is one 32 KB ring stage and one N = 64 MMA per K step:
//   modes 0 / 1 and mode 2 without a teacher: 64 columns of F_J            -> S  (64 TMEM columns)
//   mode 2 with a teacher: 32 columns, the F_J and T_J rows interleaved per K chunk -> S | CS (32 + 32)
// TMA is a high-latency path, so the ring is five stages deep.  Agents (they only meet on mbarriers):
//   warp 0 TMA producer | warps 1..3 MMA issuers (sub-tile t is issued by warp 1 + t % 3) |
//   warps 4..11 epilogue team 0 (even sub-tiles) | warps 12..19 epilogue team 1 (odd sub-tiles)
// An epilogue thread owns one row (TMEM lane quadrant = warp % 4) and half of the sub-tile's columns.
constexpr int kSwThreads = 640;
constexpr int kSwTeamThreads = 256;
constexpr int kSwStages = 5;                // ring depth with a 128-row A tile (3 with the 256-row tile of P0 / P1)
constexpr int kSwSlots = 4;                 // TMEM accumulator slots (two per team) of 64 columns per 128 rows
constexpr int kMaxSplits = 8;               // column splits of a row block (P0 / P1); partial n_i go to slots
__device__ __forceinline__ void sw_team_barrier(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kSwTeamThreads) : "memory");
}
__device__ __forceinline__ void sw_epi_barrier() { asm volatile("bar.sync 3, 512;" ::: "memory"); }

struct SweepParams {
  int N, Npad, KC, has_teacher;
  int pdl;             // launched with programmatic stream serialization: griddepcontrol.wait / launch_dependents
  int rb_lo;           // first row block of this launch and the rows it owns, [row_lo, row_hi): the whole sample
  int row_lo, row_hi;  // normally; one rank's share of the merged batch in global-negatives mode
  int splits;          // column splits: grid.y CTAs share a row block, each sweeps 1/splits of the sub-tiles
  int splits1;         // the split count of the P1 launch (P2 adds up that many partial n_i per row)
  float* npart;        // kMaxSplits planes of B*N floats: P1's per-split partial n_i (summed in split order by P2)
  float* apart;        // kMaxSplits planes: P2's per-split partial A_i (summed in split order by the backward)
  FeclScalars sc;
  float c1;            // inv_tau * log2(e)
  float inv_rows;
  double inv_rows_d;
  float hscale;
  float* hdr;
  const float* labels;       // by row of the packed operands (sorted by label when cls_lo != nullptr)
  const float* row_weight;
  const int* cls_lo;         // rows sorted by label: class bounds by row (nullptr: the caller's row order, P1 counts
  const int* cls_hi;         // the positives itself)
  int use_classes;           // specialise / skip sub-tiles by their label class (needs cls_lo)
  float* stat_m;       // zero-filled before the sweeps; split CTAs combine with atomicMax / atomicAdd (<= 2
  float* stat_p;       // contributors per address, so the float sums are order-independent: a+b == b+a)
  float* stat_n;
  float* stat_kappa;
  unsigned int* ticket;
  double* partials;
  double* sums_out;
  float* loss_out;
  ExchangeCtx x;       // sharded batch (x.world > 1): P2's last block exchanges the three sums with the peers
  // stored pairs (kStore): the loss sweep leaves, per pair, what the backward needs -- the backward then is a
  // streaming GEMM with an elementwise fix-up instead of a fourth recomputation of the similarity tiles
  void* pair_x;        // [B][Npad][Npad] 16-bit: kappa_i phi'(d) d (1-d) h_mul on positive pairs (final), 256 e_ij on
                       // negative pairs (the backward multiplies it by -kappa_i A_i h_mul / 256), 0 on the diagonal / padding
  void* pair_gc;       // [B][Npad][Npad] 16-bit: 1/(1 - cs_ij) on hard negatives, else 0 (teacher only)
  unsigned int* gc_flag;   // [B][Npad/128][Npad/64]: set where a (128-row, 64-column) tile of pair_gc is not all zero
};

struct SweepMisc {
  uint64_t a_full[8];      // one per 16 KB K chunk of the (up to two) A tiles: the first MMAs start after 16 + 32 KB
  uint64_t b_full[kSwStages], b_empty[kSwStages];
  uint64_t acc_full[kSwSlots], acc_empty[kSwSlots];
  uint32_t tmem_slot;
  uint32_t pad_;
  alignas(16) float col[2][2][2][64];   // [team][slot][stat: y, m2][column]; padded columns: y = NaN, m2 = +inf
};

// The sub-tiles a CTA walks: compact index u in [0, count) -> sub-tile t.  One contiguous run with at most one
// gap, because the sub-tiles of one label class are contiguous once the rows are sorted.
struct TileMap { int base, gap_at, gap_len, count; };
__device__ __forceinline__ int tile_of(const TileMap& tm, int u) {
  const int t = u + tm.base;
  return t >= tm.gap_at ? t + tm.gap_len : t;
}

// Cross (teacher) term of one pair: -log(1 - cs + 1e-18) over the hard negatives (dycon_losses.py:217-229), ONE
// log per chunk of 16 pairs from the product of the 64 (1 - cs) -- the power of two keeps sixteen factors inside
// fp32's range unless cs > 0.9999 on average; the caller then falls back to one log per pair (cross_chunk_end).
__device__ __forceinline__ void cross_pair(float cs, bool hard, float& cprod, float& cmin, float& nh) {
  const float fac = hard ? fmaf(-64.f, cs, 64.f) : 1.f;
  cmin = fminf(cmin, fac);           // <= 0: cs >= 1 (NaN like the reference's log of a negative number, or 1e-18)
  cprod *= fac;
  nh += hard ? 1.f : 0.f;
}

// kMode 0: row max m_i                                            (P0)
// kMode 1: negative sums n_i (and, for unsorted rows, the positive counts P_i)   (P1; needs every m)
// kMode 2: kappa_i, row loss, A_i and the teacher cross term       (P2; needs every n and P)
// Grid (row blocks, column splits, samples).  Three launches instead of one fused sweep: n_i needs all m_k
// and d_ij needs the complete n_i, i.e. two grid-wide dependencies, and splitting the columns of a row
// block over several CTAs (so that small batches still fill 148 SMs) adds a third.
// kRT = row tiles per CTA.  With kRT = 2 a CTA keeps TWO 128-row A tiles resident and every B sub-tile feeds
// two MMAs (ring of 3 stages, 128 TMEM columns per sub-tile; an epilogue thread then owns one of 256 rows and
// all columns of the sub-tile).  That halves the operand bytes per flop for P0 / P1, and -- what matters for
// the epilogue-bound P2 -- it lets the grid be (row blocks / 2) x (up to 8 column splits): 140 of the 148 SMs
// at the BraTS19 shape instead of the 112 that 128-row blocks x 2 splits reach.
// Label classes (rows sorted by label): P1 only needs the sub-tiles that hold a negative pair and P2 without a
// teacher only those that hold a positive pair -- the others are not even loaded -- and P2's epilogue runs a
// positives-only body on all-positive sub-tiles, a cross-only body on all-negative ones and the general body
// only where a class boundary crosses the sub-tile.  The class is uniform over the team, so nothing diverges.
template <int kMode, bool kBf16, int kFocal, int kRT, bool kStore = false>
__global__ void __launch_bounds__(kSwThreads, 1)
fecl_tc_sweep_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapF,
                     const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapXs,
                     const __grid_constant__ CUtensorMap mapGs, const SweepParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // (broadcast: tells the compiler that the warp index is warp-uniform, so that everything the single-thread
  //  roles derive from it -- sub-tile numbers, smem / TMEM addresses, descriptors -- lives in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int KC = p.KC;
  DYCON_SPAN(1, kMode == 3 && threadIdx.x == 0, 0);
  DYCON_SPAN_V(1, kMode == 3 && threadIdx.x == 0, 4, g_smid());
  // pair tiles leave through the staging buffers: the stored-pairs loss sweep (X | Gc) and the similarity sweep (S | Gc)
  constexpr bool kSt = kStore || kMode == 3;
  // 32 KB of staging tiles for the TMA stores take the place of one ring stage
  constexpr int kStages = kRT == 2 ? (kSt ? 2 : 3) : (kSt ? kSwStages - 1 : kSwStages);
  constexpr uint32_t kStgBytes = kSt ? 32768u : 0u;
  constexpr int kSlotCols = 64 * kRT;
  const uint32_t a_tile = (uint32_t)KC * kChunk128, a_bytes = a_tile * kRT, stage_bytes = (uint32_t)KC * kChunk64;
  uint8_t* const sA = smem;
  uint8_t* const sStage = smem + a_bytes;
  uint8_t* const sStg = sStage + kStages * stage_bytes;      // [team][X | Gc][2 regions][128 rows][16 cols] (kStore)
  SweepMisc& ms = *reinterpret_cast<SweepMisc*>(sStg + kStgBytes);
  const int b = blockIdx.z, i0 = (blockIdx.x + p.rb_lo) * kTM * kRT, split = blockIdx.y;
  const bool teacher_on = (kMode == 2 || kMode == 3) && p.has_teacher;
  const int tcols = teacher_on ? 32 : 64;                  // columns per sub-tile
  // sub-tiles that hold at least one real column (stored pairs: the backward reads 64-column tiles, so the
  // columns up to the next multiple of 64 are swept too -- they come out as zeros)
  const int nt_all = kSt ? (p.N + 63) / 64 * (64 / tcols) : (p.N + tcols - 1) / tcols;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int q = 0; q < 8; ++q) mbar_init(&ms.a_full[q], 1);
    for (int s = 0; s < kSwStages; ++s) mbar_init(&ms.b_full[s], 1), mbar_init(&ms.b_empty[s], 1);
    for (int a = 0; a < kSwSlots; ++a) mbar_init(&ms.acc_full[a], 1), mbar_init(&ms.acc_empty[a], kSwTeamThreads / 32);
    fence_mbar_init();
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapF);
    if (teacher_on) prefetch_tmap(&mapT);
    if (kSt) prefetch_tmap(&mapXs), prefetch_tmap(&mapGs);
    if ((kMode == 0 || kMode == 3) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
      p.hdr[0] = p.hscale;
      p.hdr[1] = kMode == 3 ? 1.f : (float)p.splits;       // how many partial A_i planes the backward has to add up
    }
  }
  if (warp == 1) tmem_alloc(&ms.tmem_slot, kSwSlots * kSlotCols);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ms.tmem_slot, 0);

  double red[3] = {0.0, 0.0, 0.0};

  // Programmatic dependent launch: this grid may have started while the previous sweep (or the pack kernel)
  // still runs.  P0 reads the pack kernel's output, so everybody waits here; P1 / P2 only depend on their
  // predecessor through the row statistics, so the TMA producer and the MMA issuers run ahead (operands and
  // class bounds come from the pack / rank kernels, which are complete once the predecessor has passed its own
  // wait) and only the epilogue warps wait, right before they first touch the statistics.
  if ((kMode == 0 || kMode == 3) && p.pdl) pdl_wait();
  DYCON_SPAN(1, kMode == 3 && threadIdx.x == 0, 1);

  // ---- which sub-tiles this CTA walks (every agent computes the same map) ----
  RowClass rc{0, 0x7fffffff, 0};
  if (kMode != 0 && p.use_classes) rc = load_row_class(p.cls_lo, p.cls_hi, (size_t)b * p.N, i0, kTM * kRT, p.N);
  int ta, tb;
  same_range(rc, tcols, ta, tb);
  TileMap tmap{0, 0x7fffffff, 0, nt_all};
  if (kMode == 1) {                                  // all-positive sub-tiles hold no negative pair
    tmap.gap_at = ta;
    tmap.gap_len = tb - ta;
    tmap.count = nt_all - (tb - ta);
  } else if (kMode == 2 && !teacher_on) {            // all-negative sub-tiles hold no positive pair
    const int td0 = min(rc.lo / tcols, nt_all - 1);
    const int td1 = rc.hi >= p.N ? nt_all : (rc.hi + tcols - 1) / tcols;
    tmap.base = td0;
    tmap.count = max(td1 - td0, 1);
  }
  const int u0 = (int)((long long)split * tmap.count / p.splits), u1 = (int)((long long)(split + 1) * tmap.count / p.splits);
  const int nt = u1 - u0;
  const bool tl_on = (kMode == 2 || kMode == 3) && blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 1 && lane == 0;   // (timeline build only)
  (void)tl_on;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (elect.sync, not lane == 0: with a lane test nvcc wraps EVERY tcgen05.mma / TMA instruction of a single-thread
    //  role into an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall, ~130 issue cycles per MMA -- which is what bound
    //  the sweeps: a sub-tile's 32 MMAs took 4300 cycles to issue against ~1600 to execute)
    if (elect_one()) {
      // first the K chunk the first MMA needs, then the first B sub-tile, then the rest of A
      auto load_a = [&](int h, int c) {
        mbar_expect_tx(&ms.a_full[h * 4 + c], kChunk128);
        tma_load_2d(sA + h * a_tile + c * kChunk128, &mapA, c * 64, b * p.Npad + i0 + h * kTM, &ms.a_full[h * 4 + c]);
      };
      if (nt > 0) load_a(0, 0);
      for (int t = 0; t < nt; ++t) {
        const int s = t % kStages, row = b * p.Npad + tile_of(tmap, u0 + t) * tcols;
        uint8_t* dst = sStage + s * stage_bytes;
        mbar_wait_relaxed(&ms.b_empty[s], ((t / kStages) & 1) ^ 1);
        DYCON_TL(0, tl_on, 0, t, 0);
        mbar_expect_tx(&ms.b_full[s], stage_bytes);
        for (int c = 0; c < KC; ++c) {
          tma_load_2d(dst + c * kChunk64, &mapF, c * 64, row, &ms.b_full[s]);       // box: 64 rows, or 32 with a teacher
          if (teacher_on) tma_load_2d(dst + c * kChunk64 + kChunk32, &mapT, c * 64, row, &ms.b_full[s]);
        }
        DYCON_TL(0, tl_on, 0, t, 1);
        if (t == 0) {
          for (int h = 0; h < kRT; ++h)
            for (int c = (h == 0 ? 1 : 0); c < KC; ++c) load_a(h, c);
        }
      }
    }
  } else if (warp < 4) {
    // ================================ MMA issuers =================================
    // Three warps issue alternate sub-tiles (a remnant of the lane == 0 version, whose issue loop was three times
    // slower than the tensor pipe; harmless now).
    // (stored pairs: ONE issuer.  With three issuers and a 2- or 4-stage ring an issuer can reach use u + 1 of a
    //  stage before use u has even been loaded; its wait for phase parity p then passes at once, because the barrier
    //  has not completed phase p ^ 1 yet.  In the 3-stage ring every stage has its own issuer.)
    constexpr int kIssuers = kSt ? 1 : 3;
    if (nt > 0 && warp - 1 < kIssuers && elect_one()) {
      const uint32_t idesc = umma_idesc_16(128, 64, false, false, kBf16);
      const uint64_t a_desc0 = umma_desc_kmajor(smem_u32(sA));
      for (int t = warp - 1; t < nt; t += kIssuers) {
        const bool first = t == warp - 1;         // this issuer's first sub-tile: the A chunks may still be in flight
        const int s = t % kStages, a = t & (kSwSlots - 1);
        const uint64_t b_desc0 = umma_desc_kmajor(smem_u32(sStage + s * stage_bytes));
        mbar_wait_relaxed(&ms.b_full[s], (t / kStages) & 1);
        DYCON_TL(0, tl_on, 1, t, 0);
        mbar_wait_relaxed(&ms.acc_empty[a], ((t / kSwSlots) & 1) ^ 1);
        DYCON_TL(0, tl_on, 1, t, 1);
        tcgen05_after_sync();
#pragma unroll
        for (int h = 0; h < kRT; ++h) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < KC) {
              if (first) {
                mbar_wait_relaxed(&ms.a_full[h * 4 + c], 0);
                tcgen05_after_sync();
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem + a * kSlotCols + h * 64, desc_advance(a_desc0, h * a_tile + c * kChunk128 + k * 32),
                          desc_advance(b_desc0, c * kChunk64 + k * 32), idesc, (c | k) != 0);
            }
          }
        }
        umma_commit(&ms.b_empty[s]);
        umma_commit(&ms.acc_full[a]);
        DYCON_TL(0, tl_on, 1, t, 2);
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue teams ==============================
    const int team = (warp - 4) >> 3;             // 0: even sub-tiles, 1: odd sub-tiles
    const int tt = threadIdx.x - 128 - team * kSwTeamThreads;   // 0..255 inside the team
    const bool tl_e = tl_on && (warp == 4 || warp == 12);
    (void)tl_e;
    const int quarter = warp & 3, chalf = ((warp - 4) >> 2) & 1;
    // kRT = 1: the two warp groups of a team split the columns of a sub-tile; kRT = 2: they split the two
    // 128-row tiles and every thread walks all 64 columns in two chunks
    const int rh = kRT == 2 ? chalf : 0;
    const int r = rh * kTM + quarter * 32 + lane, i = i0 + r;
    const bool row_ok = i >= p.row_lo && i < p.row_hi;
    const size_t off = (size_t)b * p.N;
    const size_t g = off + (row_ok ? i : 0);
    const float* yb = p.labels + off;
    const float* mb = p.stat_m + off;
    const float qnan = __int_as_float(0x7fc00000);
    const float yi = row_ok ? __ldg(yb + i) : qnan;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    if (kMode != 0 && kMode != 3 && p.pdl) pdl_wait();          // the statistics of the previous sweep are complete from here on
    if (tt == 0 && p.pdl) pdl_trigger();          // (after the wait) the next kernel in the stream may set up
    // this thread's columns of a sub-tile: 32 of 64 (F only), or 16 of 32 of both S and CS (teacher)
    const int cbase = kRT == 2 ? 0 : teacher_on ? chalf * 16 : chalf * 32;

    // column statistics of sub-tile T (tcols columns): threads 0..63 fetch the label, 64..127 the scaled max
    auto fetch = [&](int T) -> float {
      if (tt >= 128) return 0.f;
      const int c = tt & 63, j = T * tcols + c;
      const bool ok = c < tcols && j < p.N;
      if (tt < 64) return ok ? __ldg(yb + j) : qnan;
      if (kMode == 0 || kMode == 3) return 0.f;
      return ok ? __ldg(mb + j) * kLog2e : INFINITY;
    };
    auto publish = [&](int slot, float val) {
      if (tt < 128) ms.col[team][slot][tt >> 6][tt & 63] = val;
    };

    float n_row = 0.f, kappa = 0.f;
    if (kMode == 2) {
      // n_i = the P1 launch's per-split partial sums, added in split order (deterministic for any split count);
      // the total is also what the backward reads
      {
        float part[kMaxSplits];
#pragma unroll
        for (int q = 0; q < kMaxSplits; ++q)        // all loads in flight at once
          part[q] = q < p.splits1 ? __ldg(p.npart + (size_t)q * gridDim.z * p.N + g) : 0.f;
#pragma unroll
        for (int q = 0; q < kMaxSplits; ++q) n_row += part[q];
      }
      const bool row_writer = split == 0 && team == 0 && (kRT == 2 || chalf == 0) && row_ok;   // one thread per row
      if (row_writer) p.stat_n[g] = n_row;
      // kappa_i = r_i c_i / (B N),  c_i = 1/(P_i - 1 + 1e-18)   (dycon_losses.py:192).  P_i = the size of the row's
      // label class (sorted rows), else the count of the P1 launch
      const float rw = p.row_weight ? __ldg(p.row_weight + g) : 1.f;
      const float P_i = p.cls_lo ? (float)(__ldg(p.cls_hi + g) - __ldg(p.cls_lo + g)) : __ldg(p.stat_p + g);
      kappa = rw / ((P_i - 1.f) + kTiny) * p.inv_rows;
      if (row_writer) p.stat_kappa[g] = kappa;
    }

    // stored pairs: positive pairs carry kappa_i h_mul px (the same scaling as the H tiles of the recomputing
    // backward), negative pairs 256 e_ij; rows outside [row_lo, row_hi) and the padding store zeros (yi = NaN makes every
    // one of their pairs a "negative")
    const float kx = (kStore && row_ok) ? kappa * p.sc.inv_tau * p.hscale : 0.f;
    const float xs = (kStore && row_ok) ? 256.f : 0.f;
    // The pair terms leave through shared memory: a thread owns a ROW, so direct stores would touch 32 lines per
    // instruction.  Each team has two staging buffers (X | Gc -- without a teacher both carry X) of two regions of
    // [128 rows][16 cols]; region = the 128-row tile (kRT = 2) or the column half of the team's sub-tile (kRT = 1).
    // After a team barrier one thread issues the TMA stores into the row-major matrices [B Npad][Npad].
    uint8_t* const stg = sStg + team * 16384;
    const int x_region = kRT == 2 ? rh : chalf;
    uint8_t* const stg_row = stg + x_region * 4096 + (quarter * 32 + lane) * 32;
    const int x_row0 = b * p.Npad + i0;
    const bool x_two = kRT == 1 || i0 + kTM < p.Npad;      // kRT = 2: the second row tile may lie beyond the padded sample
    // (col0 = first column of this chunk)
    auto stage_pairs = [&](const uint32_t (&a)[8], const uint32_t (&g)[8], int col0, bool with_g) {
      if (tt == 0) bulk_wait_read0();        // the previous chunk's stores have read the staging tiles
      sw_team_barrier(team);
      uint4* dx = reinterpret_cast<uint4*>(stg_row);
      dx[0] = make_uint4(a[0], a[1], a[2], a[3]);
      dx[1] = make_uint4(a[4], a[5], a[6], a[7]);
      uint4* dg = reinterpret_cast<uint4*>(stg_row + 8192);
      dg[0] = make_uint4(g[0], g[1], g[2], g[3]);
      dg[1] = make_uint4(g[4], g[5], g[6], g[7]);
      fence_async_smem();
      sw_team_barrier(team);
      if (tt == 0) {
        // teacher: buffer 1 is Gc (same coordinates, its own matrix); else buffer 1 = the next 16 columns of X
        const CUtensorMap* m1 = with_g ? &mapGs : &mapXs;
        const int c1 = with_g ? col0 : col0 + 16;
        if (kRT == 2) {
          tma_store_2d(&mapXs, col0, x_row0, stg);
          tma_store_2d(m1, c1, x_row0, stg + 8192);
          if (x_two) {
            tma_store_2d(&mapXs, col0, x_row0 + kTM, stg + 4096);
            tma_store_2d(m1, c1, x_row0 + kTM, stg + 8192 + 4096);
          }
        } else {                             // regions = the two column halves of the sub-tile
          const int step = with_g ? 16 : 32;
          tma_store_2d(&mapXs, col0, x_row0, stg);
          tma_store_2d(m1, c1, x_row0, stg + 8192);
          tma_store_2d(&mapXs, col0 + step, x_row0, stg + 4096);
          tma_store_2d(m1, c1 + step, x_row0, stg + 8192 + 4096);
        }
        bulk_commit();
      }
    };
    // The similarity sweep (kMode 3) stages per WARP: its 32 rows x 16 columns of each buffer (1 KB + 1 KB of the 32 KB)
    // leave through TMA stores that lane 0 issues -- no team barrier, and a warp only ever waits for its own previous
    // stores (the team-wide version above spent a third of the sweep in its two barriers per chunk, every thread waiting
    // for the slowest warp and for the TMA unit to have read the tile).
    uint8_t* const wbuf = sStg + (warp - 4) * 2048;
    const int w_row0 = x_row0 + rh * kTM + quarter * 32;        // row (of the whole matrix) of this warp's lane 0
    const bool w_rows_ok = kRT == 1 || rh == 0 || x_two;
    auto stage_warp = [&](const uint32_t (&a)[8], const uint32_t (&g)[8], int col_a, const CUtensorMap* mg, int col_g) {
      if (lane == 0) bulk_wait_read0();      // the previous chunk's stores have read this warp's tiles
      __syncwarp();
      uint4* dx = reinterpret_cast<uint4*>(wbuf + lane * 32);
      dx[0] = make_uint4(a[0], a[1], a[2], a[3]);
      dx[1] = make_uint4(a[4], a[5], a[6], a[7]);
      uint4* dg = reinterpret_cast<uint4*>(wbuf + 1024 + lane * 32);
      dg[0] = make_uint4(g[0], g[1], g[2], g[3]);
      dg[1] = make_uint4(g[4], g[5], g[6], g[7]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0 && w_rows_ok) {
        tma_store_2d(&mapXs, col_a, w_row0, wbuf);
        tma_store_2d(mg, col_g, w_row0, wbuf + 1024);
        bulk_commit();
      }
    };
    if (team < nt) publish(0, fetch(tile_of(tmap, u0 + team)));
    sw_team_barrier(team);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int it = 0;                                   // team-local iteration: sub-tile t = team + 2 * it
    for (int t = team; t < nt; t += 2, ++it) {
      const int T = tile_of(tmap, u0 + t);        // sub-tile index inside the sample
      const int slot = it & 1, a = t & (kSwSlots - 1), j0 = T * tcols;
      const int cls = (kMode == 0 || kSt) ? kClsMixed : tile_class(rc, T, tcols, p.N, ta, tb);      // uniform over the team
      DYCON_TL(0, tl_e, 2 + team, t, 0);
      const float nxt = fetch(tile_of(tmap, u0 + (t + 2 < nt ? t + 2 : t)));
      mbar_wait(&ms.acc_full[a], (t / kSwSlots) & 1);
      DYCON_TL(0, tl_e, 2 + team, t, 1);
      tcgen05_after_sync();
      if (!teacher_on) {
#pragma unroll 1
        for (int ch = 0; ch < kRT; ++ch) {        // kRT = 2: the two 32-column chunks of this thread's row
          const int cb = cbase + ch * 32;
          const float* cy = &ms.col[team][slot][0][cb];
          const float* cm = &ms.col[team][slot][1][cb];
          const int rdiag = i - j0 - cb;          // chunk-local column of the diagonal pair, if in range
          float v[32];
          tmem_ld32(tmem + lane_base + a * kSlotCols + rh * 64 + cb, v);
          tmem_ld_wait();
          if (ch == kRT - 1) {                    // last read of this accumulator: hand it back to the MMA issuers
            tcgen05_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ms.acc_empty[a]);
          }
          // does any row of this warp meet its diagonal inside this chunk?  (warp-uniform)
          const int w0 = i0 + rh * kTM + quarter * 32 - j0 - cb;     // rdiag of lane 0
          const bool diag_here = w0 + 31 >= 0 && w0 < 32;
          if (kMode == 0) {          // acc0 = max_j l_ij, the zeroed diagonal takes part (>= 0)   (dycon_losses.py:176-181)
            if (diag_here) {
#pragma unroll
              for (int c = 0; c < 32; ++c) acc0 = fmaxf(acc0, c == rdiag ? 0.f : v[c]);
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) acc0 = fmaxf(acc0, v[c]);     // padded columns hold S = 0 <= acc0
            }
          } else if (kMode == 1) {   // acc0 = n_i partial, acc1 = positive count                  (dycon_losses.py:183-184,192)
            if (cls == kClsDiff) {   // every pair is a negative (padded columns: m2 = +inf -> e = 0)
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 mm = *reinterpret_cast<const float4*>(cm + q * 4);
                acc0 += ex2_approx(fmaf(v[q * 4], p.c1, -mm.x)) + ex2_approx(fmaf(v[q * 4 + 1], p.c1, -mm.y));
                acc0 += ex2_approx(fmaf(v[q * 4 + 2], p.c1, -mm.z)) + ex2_approx(fmaf(v[q * 4 + 3], p.c1, -mm.w));
              }
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
                const float4 mm = *reinterpret_cast<const float4*>(cm + q * 4);
                const float ys[4] = {yy.x, yy.y, yy.z, yy.w}, m2[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float e = ex2_approx(fmaf(v[q * 4 + k], p.c1, -m2[k]));   // padded: m2 = +inf -> e = 0
                  const bool same = ys[k] == yi;
                  acc0 += same ? 0.f : e;
                  acc1 += same ? 1.f : 0.f;
                }
              }
            }
          } else if (kMode == 3) {   // acc0 = row max (as mode 0), and the similarities themselves go to pair_x (16-bit)
            if (diag_here) {
#pragma unroll
              for (int c = 0; c < 32; ++c) acc0 = fmaxf(acc0, c == rdiag ? 0.f : v[c]);
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) acc0 = fmaxf(acc0, v[c]);
            }
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              lo[q] = Cvt<kBf16>::two(v[2 * q], v[2 * q + 1]);
              hi[q] = Cvt<kBf16>::two(v[16 + 2 * q], v[16 + 2 * q + 1]);
            }
            stage_warp(lo, hi, j0 + cb, &mapXs, j0 + cb + 16);
          } else if (kStore) {       // loss / A partials, and the pair terms of the backward go to pair_x
            uint32_t xp[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
              const float4 mm = *reinterpret_cast<const float4*>(cm + q * 4);
              const float ys[4] = {yy.x, yy.y, yy.z, yy.w}, m2[4] = {mm.x, mm.y, mm.z, mm.w};
              float xv[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = q * 4 + k;
                const float tl = fmaf(v[c], p.c1, -m2[k]);
                const float e = ex2_approx(tl);                   // padded column: m2 = +inf -> e = 0
                float phi2, at, px, rom;
                pos_fwd_x<kFocal>(tl, e, n_row, p.sc.gamma, 1.f, phi2, at, px, rom);
                const bool same = ys[k] == yi;
                const bool pos = diag_here ? same && (c != rdiag) : same;
                acc0 += pos ? phi2 : 0.f;
                acc1 += pos ? at : 0.f;
                xv[k] = pos ? px * kx : same ? 0.f : e * xs;       // (same && !pos: the diagonal)
              }
              xp[q * 2] = Cvt<kBf16>::two(xv[0], xv[1]);
              xp[q * 2 + 1] = Cvt<kBf16>::two(xv[2], xv[3]);
            }
            {
              const uint32_t lo[8] = {xp[0], xp[1], xp[2], xp[3], xp[4], xp[5], xp[6], xp[7]};
              const uint32_t hi[8] = {xp[8], xp[9], xp[10], xp[11], xp[12], xp[13], xp[14], xp[15]};
              stage_pairs(lo, hi, j0 + (kRT == 2 ? cb : 0), false);
            }
          } else {                   // acc0 / acc1 = loss / A partials                              (dycon_losses.py:186-206)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
              const float4 mm = *reinterpret_cast<const float4*>(cm + q * 4);
              const float ys[4] = {yy.x, yy.y, yy.z, yy.w}, m2[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = q * 4 + k;
                const float tl = fmaf(v[c], p.c1, -m2[k]);
                float phi2, at;
                pos_fwd<kFocal>(tl, ex2_approx(tl), n_row, p.sc.gamma, phi2, at);
                const bool same = cls == kClsSame || ys[k] == yi;                 // (cls is uniform: no divergence)
                const bool pos = diag_here ? same && (c != rdiag) : same;         // select, never multiply
                acc0 += pos ? phi2 : 0.f;
                acc1 += pos ? at : 0.f;
              }
            }
          }
        }
      } else {
        // teacher mode (kMode == 2): chunks of 16 columns of S and the same 16 columns of CS (kRT = 1: one chunk,
        // this warp group's half of the 32-column sub-tile; kRT = 2: both halves, one after the other)
#pragma unroll 1
        for (int ch = 0; ch < kRT; ++ch) {
          const int cb = kRT == 2 ? ch * 16 : cbase;
          const float* cy = &ms.col[team][slot][0][cb];
          const float* cm = &ms.col[team][slot][1][cb];
          const int rdiag = i - j0 - cb;          // chunk-local column of the diagonal pair, if in range
          float v[16], w[16];
          if (cls != kClsDiff) tmem_ld16(tmem + lane_base + a * kSlotCols + rh * 64 + cb, v);        // S: positives
          if (cls != kClsSame) tmem_ld16(tmem + lane_base + a * kSlotCols + rh * 64 + 32 + cb, w);   // CS: negatives
          tmem_ld_wait();
          if (ch == kRT - 1) {
            tcgen05_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ms.acc_empty[a]);
          }
          const int w0 = i0 + rh * kTM + quarter * 32 - j0 - cb;
          const bool diag_here = w0 + 31 >= 0 && w0 < 16;
          if (kMode == 3) {
            // similarity sweep: row max of S, S itself to pair_x, and the WHOLE cross term -- it depends on no row
            // statistic: -log(1 - cs + 1e-18) over the hard negatives (dycon_losses.py:217-229) and pair_gc =
            // 1 / (64 (1 - cs)) for the backward
            if (diag_here) {
#pragma unroll
              for (int c = 0; c < 16; ++c) acc0 = fmaxf(acc0, c == rdiag ? 0.f : v[c]);
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c) acc0 = fmaxf(acc0, v[c]);
            }
            uint32_t xp[8], gp[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              xp[q] = Cvt<kBf16>::two(v[2 * q], v[2 * q + 1]);
              gp[q] = 0u;
            }
            // label masks (1.0f where the labels differ; the pair arithmetic below runs on pairs of columns in packed
            // fp32).  A chunk in which NO row of the warp meets a column of another label -- most chunks of a sample
            // whose foreground is a compact region -- carries no cross term at all: warp-uniform skip.
            float2 df[8], nd2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
              df[2 * q] = make_float2(mask_ne(yy.x, yi), mask_ne(yy.y, yi));
              df[2 * q + 1] = make_float2(mask_ne(yy.z, yi), mask_ne(yy.w, yi));
              nd2 = __fadd2_rn(nd2, __fadd2_rn(df[2 * q], df[2 * q + 1]));
            }
            if (__any_sync(0xffffffffu, row_ok && nd2.x + nd2.y > 0.f)) {
              float cmin = 1.f;
              float2 cprod2 = make_float2(1.f, 1.f), nh2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float2 cs2 = make_float2(w[2 * q], w[2 * q + 1]);
                const float2 hard = __fmul2_rn(df[q], make_float2(mask_gt(cs2.x, p.sc.cross_thresh), mask_gt(cs2.y, p.sc.cross_thresh)));
                // fac = 64 (1 - cs) on a hard negative, else 1
                const float2 fac = __ffma2_rn(hard, __ffma2_rn(cs2, make_float2(-64.f, -64.f), make_float2(63.f, 63.f)), make_float2(1.f, 1.f));
                cmin = fminf(cmin, fminf(fac.x, fac.y));
                cprod2 = __fmul2_rn(cprod2, fac);
                nh2 = __fadd2_rn(nh2, hard);
                const float2 gv = __fmul2_rn(hard, make_float2(rcp_approx(fac.x), rcp_approx(fac.y)));
                gp[q] = Cvt<kBf16>::two(gv.x, gv.y);
              }
              const float nh = nh2.x + nh2.y, cprod = cprod2.x * cprod2.y;
              const bool rowhard = row_ok && nh > 0.f;
              if (!rowhard) {
#pragma unroll
                for (int u = 0; u < 8; ++u) gp[u] = 0u;
              }
              if (__any_sync(0xffffffffu, rowhard) && lane == 0)
                p.gc_flag[((size_t)b * (p.Npad >> 7) + (i >> 7)) * (p.Npad >> 6) + ((j0 + cb) >> 6)] = 1u;
              acc3 += nh;
              if (cmin > 0.f && cprod >= 1e-30f) {
                acc2 += fmaf(-6.f, nh, lg2_approx(cprod));
              } else {
                // rare: a pair with cs >= 1 (NaN / log(1e-18), as in the reference) or collapsed pairs whose product
                // underflows -- one log per pair
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
                  const float ys[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float cs = w[q * 4 + k];
                    if (!(ys[k] == yi) && cs > p.sc.cross_thresh) acc2 += lg2_approx((1.f - cs) + kTiny);
                  }
                }
              }
            }
            stage_warp(xp, gp, j0 + cb, &mapGs, j0 + cb);
          } else if (kStore) {
            // general body + the pair terms of the backward: pair_x as above, pair_gc = 1 / (64 (1 - cs)) on hard
            // negatives (the 64 of cross_pair; the backward folds it into its scalar factor)
            float cprod = 1.f, cmin = 1.f, nh = 0.f;
            uint32_t xp[8], gp[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
              const float4 mm = *reinterpret_cast<const float4*>(cm + q * 4);
              const float ys[4] = {yy.x, yy.y, yy.z, yy.w}, m2[4] = {mm.x, mm.y, mm.z, mm.w};
              float xv[4], gv[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = q * 4 + k;
                const float tl = fmaf(v[c], p.c1, -m2[k]);
                const float e = ex2_approx(tl);
                const bool same = ys[k] == yi;
                const bool hard = !same && w[c] > p.sc.cross_thresh;
                const float fac = hard ? fmaf(-64.f, w[c], 64.f) : 1.f;
                cmin = fminf(cmin, fac);
                cprod *= fac;
                nh += hard ? 1.f : 0.f;
                float phi2, at, px, rom;
                pos_fwd_x<kFocal>(tl, e, n_row, p.sc.gamma, fac, phi2, at, px, rom);
                const bool pos = diag_here ? same && (c != rdiag) : same;
                acc0 += pos ? phi2 : 0.f;
                acc1 += pos ? at : 0.f;
                xv[k] = pos ? px * kx : same ? 0.f : e * xs;
                gv[k] = hard ? rom : 0.f;
              }
              xp[q * 2] = Cvt<kBf16>::two(xv[0], xv[1]);
              xp[q * 2 + 1] = Cvt<kBf16>::two(xv[2], xv[3]);
              gp[q * 2] = Cvt<kBf16>::two(gv[0], gv[1]);
              gp[q * 2 + 1] = Cvt<kBf16>::two(gv[2], gv[3]);
            }
            {
              const bool rowhard = row_ok && nh > 0.f;
              if (!rowhard) {
#pragma unroll
                for (int u = 0; u < 8; ++u) gp[u] = 0u;
              }
              if (__any_sync(0xffffffffu, rowhard) && lane == 0)
                p.gc_flag[((size_t)b * (p.Npad >> 7) + (i >> 7)) * (p.Npad >> 6) + ((j0 + cb) >> 6)] = 1u;
              stage_pairs(xp, gp, j0 + (kRT == 2 ? cb : 0), true);
            }
            acc3 += nh;
            if (cmin > 0.f && cprod >= 1e-30f) {
              acc2 += fmaf(-6.f, nh, lg2_approx(cprod));
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
                const float ys[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float cs = w[q * 4 + k];
                  if (!(ys[k] == yi) && cs > p.sc.cross_thresh) acc2 += lg2_approx((1.f - cs) + kTiny);
                }
              }
            }
          } else if (cls == kClsSame) {
            // all-positive sub-tile: student terms only, no label test, no hard negatives
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 mm = *reinterpret_cast<const float4*>(cm + q * 4);
              const float m2[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = q * 4 + k;
                const float tl = fmaf(v[c], p.c1, -m2[k]);
                float phi2, at;
                pos_fwd<kFocal>(tl, ex2_approx(tl), n_row, p.sc.gamma, phi2, at);
                const bool pos = !diag_here || c != rdiag;
                acc0 += pos ? phi2 : 0.f;
                acc1 += pos ? at : 0.f;
              }
            }
          } else {
            // cross term: -log(1 - cs + 1e-18) over labels differ && cs > thresh (dycon_losses.py:217-229).
            // Padded columns (y = NaN, so never "same") have cs == 0 exactly (zero teacher rows) and the host
            // guarantees thresh >= 0, so they are never hard negatives; padded rows are dropped at the end.
            float cprod = 1.f, cmin = 1.f, nh = 0.f;
            if (cls == kClsDiff) {
              // all-negative sub-tile: no student term at all (the loss sums over positives only)
#pragma unroll
              for (int c = 0; c < 16; ++c) cross_pair(w[c], w[c] > p.sc.cross_thresh, cprod, cmin, nh);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 yy = *reinterpret_cast<const float4*>(cy + q * 4);
                const float4 mm = *reinterpret_cast<const float4*>(cm + q * 4);
                const float ys[4] = {yy.x, yy.y, yy.z, yy.w}, m2[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int c = q * 4 + k;
                  const float tl = fmaf(v[c], p.c1, -m2[k]);
                  float phi2, at;
                  pos_fwd<kFocal>(tl, ex2_approx(tl), n_row, p.sc.gamma, phi2, at);
                  const bool same = ys[k] == yi;
                  const bool pos = diag_here ? same && (c != rdiag) : same;
                  acc0 += pos ? phi2 : 0.f;
                  acc1 += pos ? at : 0.f;
                  cross_pair(w[c], !same && w[c] > p.sc.cross_thresh, cprod, cmin, nh);
                }
              }
            }
            acc3 += nh;
            if (cmin > 0.f && cprod >= 1e-30f) {
              acc2 += fmaf(-6.f, nh, lg2_approx(cprod));      // sum of log2(1 - cs): undo the 64 per hard negative
            } else {
              // rare: a pair with cs >= 1 (NaN / log(1e-18), as in the reference) or sixteen collapsed pairs whose
              // product underflows -- one log per pair
              const float4* cy4 = reinterpret_cast<const float4*>(cy);
#pragma unroll
              for (int q = 0; q < 4; ++q) {      // (unrolled: w[] must stay in registers)
                const float4 yy = cy4[q];
                const float ys[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float cs = w[q * 4 + k];
                  const bool hard = (cls == kClsDiff || !(ys[k] == yi)) && cs > p.sc.cross_thresh;
                  if (hard) acc2 += lg2_approx((1.f - cs) + kTiny);
                }
              }
            }
          }
        }
      }
      DYCON_TL(0, tl_e, 2 + team, t, 2);
      publish(slot ^ 1, nxt);
      sw_team_barrier(team);
      DYCON_TL(0, tl_e, 2 + team, t, 3);
    }
    DYCON_TL(0, tl_e, 2 + team, 62, 0);
    if (kMode == 3 ? lane == 0 : (kStore && tt == 0)) bulk_wait0();      // this warp's / team's TMA stores are complete

    // ---- combine the threads that share a row (kRT = 1: 2 teams x 2 column halves; kRT = 2: the 2 teams),
    //      then the column splits ----
    sw_epi_barrier();                    // every sub-tile is consumed: all MMAs are done, the ring is free
    float* xch = reinterpret_cast<float*>(sStage);       // [writers - 1][rows of the CTA][4]
    constexpr int kRows = kTM * kRT, kWriters = kRT == 2 ? 2 : 4;
    const int wr = kRT == 2 ? team : team * 2 + chalf;
    if (wr != 0) *reinterpret_cast<float4*>(xch + ((wr - 1) * kRows + r) * 4) = make_float4(acc0, acc1, acc2, acc3);
    sw_epi_barrier();
    if (wr == 0) {
#pragma unroll
      for (int u = 0; u < kWriters - 1; ++u) {
        const float4 o = *reinterpret_cast<const float4*>(xch + (u * kRows + r) * 4);
        if (kMode == 0 || kMode == 3) acc0 = fmaxf(acc0, o.x); else acc0 += o.x;
        acc1 += o.y; acc2 += o.z; acc3 += o.w;
      }
      // (a CTA without sub-tiles -- more splits than sub-tiles left after skipping -- still writes its zero
      //  partials: the consumers add up every split's plane)
      if (row_ok) {
        if (kMode == 0 || kMode == 3) {
          const float m = acc0 * p.sc.inv_tau;                           // >= 0, so the int ordering is the float ordering
          atomicMax(reinterpret_cast<int*>(p.stat_m + g), __float_as_int(m));
          if (kMode == 3) {
            red[1] = (double)(-kLn2 * acc2);
            red[2] = (double)acc3;
          }
        } else if (kMode == 1) {
          p.npart[(size_t)split * gridDim.z * p.N + g] = acc0;           // summed in split order by P2 (deterministic)
          if (!p.cls_lo) atomicAdd(p.stat_p + g, acc1);                  // integer-valued: exact in any order
        } else {
          p.apart[(size_t)split * gridDim.z * p.N + g] = acc1;           // summed in split order by the backward
          red[0] = (double)(kappa * (-kLn2 * acc0));                      // kappa_i = r_i c_i inv_rows
          red[1] = (double)(-kLn2 * acc2);
          red[2] = (double)acc3;
        }
      }
    }
  }

  // ---- block / grid reduction of {student, cross_sum, cross_cnt} (mode 2) ----
  if (kMode == 2) {
    double* scratch = reinterpret_cast<double*>(sStage + 8192);    // the ring is idle by now (xch uses its first 6 KB)
    double total_[3];
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (grid_sum_last_block<3>(red, total_, p.ticket, p.partials, nblocks, bid, scratch, reinterpret_cast<int*>(sStage + 12288)) &&
        warp == 0) {
      // totals are in thread 0; with a sharded batch warp 0 runs the exchange (rank-ordered sums of all ranks)
      double t0 = __shfl_sync(0xffffffffu, total_[0], 0), t1 = __shfl_sync(0xffffffffu, total_[1], 0),
             t2 = __shfl_sync(0xffffffffu, total_[2], 0);
      if (p.x.world > 1) {
        const double tot = exchange_warp(p.x, lane == 0 ? t0 : lane == 1 ? t1 : t2, 3);
        t0 = __shfl_sync(0xffffffffu, tot, 0);
        t1 = __shfl_sync(0xffffffffu, tot, 1);
        t2 = __shfl_sync(0xffffffffu, tot, 2);
      }
      if (lane == 0) {
        const double student = t0 / p.inv_rows_d;
        p.sums_out[0] = student;
        p.sums_out[1] = t1;
        p.sums_out[2] = t2;
        if (p.loss_out) {
          const double cross = p.has_teacher ? t1 / (t2 + 1e-18) : 0.0;
          *p.loss_out = (float)(student * p.inv_rows_d + (double)p.sc.lambda_cross * cross);
        }
      }
    }
  }
  // similarity sweep: the LOCAL cross sums wait in sums_out[1..2] for the row kernel, whose last block adds the
  // student sum, runs the exchange of a sharded batch and writes the loss
  if (kMode == 3) {
    double* scratch = reinterpret_cast<double*>(sStage + 8192);
    double total_[3];
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (grid_sum_last_block<3>(red, total_, p.ticket, p.partials, nblocks, bid, scratch, reinterpret_cast<int*>(sStage + 12288)) &&
        threadIdx.x == 0) {
      p.sums_out[1] = total_[1];
      p.sums_out[2] = total_[2];
    }
  }
  DYCON_TL(0, tl_on && warp == 4, 2, 63, 0);
  DYCON_SPAN(1, kMode == 3 && threadIdx.x == 128, 2);           // (warp 4: its epilogue loop and the grid sum are done)
  DYCON_SPAN_V(1, kMode == 3 && threadIdx.x == 128, 5, nt);
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kSwSlots * kSlotCols);
  DYCON_SPAN(1, kMode == 3 && threadIdx.x == 0, 3);
}

// =================================================================================================
//  Row kernel: everything of the student term that needs the row statistics, from the stored similarities
// =================================================================================================
// The similarity sweep (kMode 3) has left S = F F^T in pair_x as 16-bit numbers (rounding S to fp16 costs less than
// the rounding of the operands did: |dS| <= 2.4e-4 against a logit scale of 1/tau) together with every row maximum
// m.  What remains of the forward -- n_i, kappa_i, the row losses, A_i -- and the per-pair gradient terms X need
// complete ROWS, not tiles: a warp owns a row, keeps its e_ij = 2^(S_ij c1 - m2_j) in registers and walks it three
// times,
//   pass 1   n_i = sum of the negatives' e_ij, P_i = number of same-label columns        (dycon_losses.py:183-184,192)
//   pass 2   positives: phi(d_ij) -> row loss, A_i, X_ij = kappa_i h phi'(d) d (1 - d)     (:186-206)
//   pass 3   negatives: X_ij = -kappa_i A_i h e_ij                                        (A_i is complete now)
// and writes X over S in place.  No grid-wide dependency between the passes (a row is warp-local), no tensor core,
// no TMEM round trip: a plain SIMT kernel at full occupancy, where MUFU and issue slots overlap.  The backward
// then is a pure streaming GEMM over X (no fix-up).  Columns are handled in groups of eight (one 16-byte load per
// lane and 256-column chunk); kChunks = ceil(Npad / 256) <= 8 keeps a row in registers (N <= 2048; longer rows use
// the three-sweep forward).
// Threads per CTA x resident CTAs per SM.  Measured on the B200 at the BraTS19 shape (forward, us): 256 x 1 -> 55.6,
// 256 x 2 -> 57.7, 128 x 3 -> 57.7, 128 x 4 -> 61.2: eight warps per SM with the whole register file (no spills, the pair
// loops interleaved as deep as ptxas likes) beat sixteen warps at 128 registers.
#ifndef DYCON_ROW_THREADS
#define DYCON_ROW_THREADS 256
#define DYCON_ROW_MINB 1
#endif
constexpr int kRowThreads = DYCON_ROW_THREADS;
constexpr int kRowMinBlocks = DYCON_ROW_MINB;
constexpr int kRowMaxChunks = 8;

struct RowParams {
  int N, Npad, ncol;         // ncol = the columns the similarity sweep wrote (N rounded up to 64)
  int rows_per_cta;
  int pdl;
  float c1, gamma, kh, inv_rows;      // kh = inv_tau * hscale
  double inv_rows_d;
  float lambda_cross;
  int has_teacher;
  const float* labels;
  const float* row_weight;
  const float* stat_m;
  void* pair_x;              // in: S (16-bit), out: X
  float* stat_n;
  float* stat_kappa;
  float* stat_a;             // plane 0 of the partial-A planes (hdr[1] = 1)
  unsigned int* ticket;
  double* partials;
  double* sums_out;
  float* loss_out;
  ExchangeCtx x;
};

// The student terms of one positive pair from e = e_ij and n = n_i (pos_fwd_x without the cross factor):
// phi2 (phi = -ln2 phi2), the A-summand phi'(d) d / T and px = phi'(d) d (1 - d).
template <int kFocal>
__device__ __forceinline__ void pos_terms(float e, float n, float gamma, float& phi2, float& a_term, float& px) {
  const float T = e + n;
  const float rT = rcp_approx(T);
  const float d = e * rT;
  const float L = lg2_approx(fmaxf(d, 1e-30f));      // log2 d (finite even where e underflows)
  if (kFocal == kNoFocal) {
    phi2 = L;
    a_term = -rT;
    px = -(n * rT);                       // -(1 - d)
  } else {
    float omd = 1.f - d;
    const float w1 = kFocal == kFocalG2 ? omd : pow_gm1(omd, gamma);
    const float w = w1 * omd;
    const float y = fmaf(gamma * kLn2 * d, L, -omd);
    phi2 = L * w;
    a_term = w1 * y * rT;
    px = w * y;
  }
}

// The same on a pair of columns in packed fp32: from d = e / T, L = log2 d and rT = 1 / T.
template <int kFocal>
__device__ __forceinline__ void pos_tail2(float2 d, float2 L, float2 rT, float2 n, float gamma, float2& phi2, float2& a_term,
                                          float2& px) {
  if (kFocal == kNoFocal) {
    phi2 = L;
    a_term = make_float2(-rT.x, -rT.y);
    const float2 nr = __fmul2_rn(n, rT);
    px = make_float2(-nr.x, -nr.y);
  } else {
    float2 omd = __fadd2_rn(make_float2(1.f, 1.f), make_float2(-d.x, -d.y));
    float2 w1 = omd;
    if (kFocal != kFocalG2) {
      w1.x = pow_gm1(omd.x, gamma);      // (clamps omd at 0)
      w1.y = pow_gm1(omd.y, gamma);
    }
    const float gl = gamma * kLn2;
    const float2 w = __fmul2_rn(w1, omd);
    const float2 y = __ffma2_rn(__fmul2_rn(make_float2(gl, gl), d), L, make_float2(-omd.x, -omd.y));
    phi2 = __fmul2_rn(L, w);
    a_term = __fmul2_rn(__fmul2_rn(w1, y), rT);
    px = __fmul2_rn(w, y);
  }
}
// d, L, rT through the special-function unit: one MUFU.RCP and one MUFU.LG2 per pair (any n).
template <int kFocal>
__device__ __forceinline__ void pos_terms2(float2 e, float2 n, float gamma, float2& phi2, float2& a_term, float2& px) {
  const float2 T = __fadd2_rn(e, n);
  const float2 rT = make_float2(rcp_approx(T.x), rcp_approx(T.y));
  const float2 d = __fmul2_rn(e, rT);
  const float2 L = make_float2(lg2_approx(fmaxf(d.x, 1e-30f)), lg2_approx(fmaxf(d.y, 1e-30f)));
  pos_tail2<kFocal>(d, L, rT, n, gamma, phi2, a_term, px);
}
// c ? a : b as ONE selp: left to itself nvcc turns the nested selects of the pair loops into a branch region per pair
// (BSSY / BRA / BSYNC), which serialises the MUFU latencies of the 56 pairs a lane walks
__device__ __forceinline__ float fsel(bool c, float a, float b) {
  float r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f32 %0, %1, %2, p;\n\t}" : "=f"(r) : "f"(a), "f"(b), "r"((int)c));
  return r;
}

__device__ __forceinline__ void rp_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}


template <int kFocal, int kChunks>
__global__ void __launch_bounds__(kRowThreads, kRowMinBlocks)
fecl_row_pairs_kernel(const RowParams p) {
  // column statistics of the sample, permuted so that the two 16-byte reads of a lane's group of eight columns are
  // conflict-free: column 256 k + 8 l + q lives at [k][q >> 2][l][q & 3]
  extern __shared__ __align__(16) float rp_smem[];
  float* const ys = rp_smem;                       // label (padding: NaN -- never "same")
  float* const nm2 = rp_smem + kChunks * 256;      // -m_j log2(e) (padding: -inf -> e = 0)
  // per warp: the NEXT row of S, fetched with cp.async while the current row is being worked on (a warp walks its rows
  // one after the other, four warps per scheduler: without the prefetch a third of the kernel is load latency)
  uint8_t* const wb = reinterpret_cast<uint8_t*>(rp_smem + 2 * kChunks * 256) + (threadIdx.x >> 5) * (kChunks * 512);
  __shared__ double rp_scratch[32];
  __shared__ int rp_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, N = p.N, Npad = p.Npad;
  const size_t off = (size_t)b * N;
  const float qnan = __int_as_float(0x7fc00000);
  const int r0 = blockIdx.x * p.rows_per_cta;
  DYCON_SPAN(2, tid == 0, 0);
  DYCON_SPAN_V(2, tid == 0, 4, g_smid());
  const int r1 = min(r0 + p.rows_per_cta, p.ncol);     // rows up to ncol are read (as X_JI tiles) by the backward
  double red[1] = {0.0};
  auto prefetch = [&](int r) {
    if (r < r1 && r < N) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.pair_x) + ((size_t)b * Npad + r) * Npad * 2;
#pragma unroll
      for (int k = 0; k < kChunks; ++k) {
        const int c0 = k * 256 + lane * 8;
        if (c0 < p.ncol) rp_cp_async16(wb + c0 * 2, src + c0 * 2);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // the labels are inputs of the step: fetched while the sweep still runs (programmatic dependent launch)
  // (all loads of a thread first, then the stores: one memory round trip per array, not one per element)
  constexpr int kFill = kChunks * 256 / kRowThreads;
  {
    float v[kFill];
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int j = tid + u * kRowThreads;
      v[u] = j < N ? __ldg(p.labels + off + j) : qnan;
    }
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int j = tid + u * kRowThreads, k = j >> 8, l = (j >> 3) & 31, q = j & 7;
      ys[k * 256 + (q >> 2) * 128 + l * 4 + (q & 3)] = v[u];
    }
  }
  if (p.pdl) pdl_wait();                           // row maxima and similarities of the sweep are complete
  DYCON_SPAN(2, tid == 0, 1);
  if (tid == 0 && p.pdl) pdl_trigger();
  prefetch(r0 + warp);                             // the first row travels while the column statistics are set up
  {
    float v[kFill];
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int j = tid + u * kRowThreads;
      v[u] = j < N ? __ldcg(p.stat_m + off + j) : INFINITY;
    }
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int j = tid + u * kRowThreads, k = j >> 8, l = (j >> 3) & 31, q = j & 7;
      nm2[k * 256 + (q >> 2) * 128 + l * 4 + (q & 3)] = -(v[u] * kLog2e);
    }
  }
  __syncthreads();
  for (int i = r0 + warp; i < r1; i += kRowThreads / 32) {
    uint4* row = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.pair_x) + ((size_t)b * Npad + i) * Npad);
    if (i >= N) {                                   // padded rows: X = 0
#pragma unroll
      for (int k = 0; k < kChunks; ++k) {
        const int c0 = k * 256 + lane * 8;
        if (c0 < p.ncol) row[c0 >> 3] = make_uint4(0u, 0u, 0u, 0u);
      }
      continue;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();                                   // (the diagonal below was fetched by another lane)
    uint4 sv[kChunks];
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const int c0 = k * 256 + lane * 8;
      sv[k] = c0 < p.ncol ? *reinterpret_cast<const uint4*>(wb + c0 * 2) : make_uint4(0u, 0u, 0u, 0u);
    }
    const __half s_diag = *reinterpret_cast<const __half*>(wb + i * 2);
    __syncwarp();                                   // every lane has read: the buffer may take the next row
    prefetch(i + kRowThreads / 32);
    const float yi = __ldg(p.labels + off + i);
    const float rw = p.row_weight ? __ldg(p.row_weight + off + i) : 1.f;
    // the diagonal pair (i, i) belongs to lane (i / 8) % 32: it counts as a same-label column (P_i), carries neither
    // loss nor gradient (l_ii is multiplied by 0, dycon_losses.py:176-178).  The pair loops treat it like any positive
    // pair; its terms are taken out of the row sums afterwards and its X is cleared after the row's stores.
    const bool diag_lane = ((i >> 3) & 31) == lane;

    // The pair arithmetic runs on PAIRS of columns in packed fp32 (FFMA2 / FMUL2 / FADD2: two lanes' worth of fp32 per
    // issue slot); MUFU, the label masks and the 16-bit conversions stay scalar.
    // ---- pass 1: e_ij, n_i, P_i ----
    float2 e[kChunks * 4];
    float2 n2 = make_float2(0.f, 0.f), cnt2 = make_float2(0.f, 0.f);   // cnt: columns with a DIFFERENT label (padding included)
    const float2 c1c1 = make_float2(p.c1, p.c1);
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const uint32_t w4[4] = {sv[k].x, sv[k].y, sv[k].z, sv[k].w};
      const float4 ya = *reinterpret_cast<const float4*>(ys + k * 256 + lane * 4);
      const float4 yb = *reinterpret_cast<const float4*>(ys + k * 256 + 128 + lane * 4);
      const float4 ma = *reinterpret_cast<const float4*>(nm2 + k * 256 + lane * 4);
      const float4 mb = *reinterpret_cast<const float4*>(nm2 + k * 256 + 128 + lane * 4);
      const float2 yy[4] = {make_float2(ya.x, ya.y), make_float2(ya.z, ya.w), make_float2(yb.x, yb.y), make_float2(yb.z, yb.w)};
      const float2 mm[4] = {make_float2(ma.x, ma.y), make_float2(ma.z, ma.w), make_float2(mb.x, mb.y), make_float2(mb.z, mb.w)};
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        const float2 tl = __ffma2_rn(__half22float2(*reinterpret_cast<const __half2*>(&w4[q2])), c1c1, mm[q2]);
        const float2 ev = make_float2(ex2_approx(tl.x), ex2_approx(tl.y));            // padded columns: e = 0
        e[k * 4 + q2] = ev;
        const float2 df = make_float2(mask_ne(yy[q2].x, yi), mask_ne(yy[q2].y, yi));
        n2 = __ffma2_rn(df, ev, n2);
        cnt2 = __fadd2_rn(cnt2, df);
      }
    }
    const float n_i = warp_sum(n2.x + n2.y);
    const float cnt = (float)(kChunks * 256) - warp_sum(cnt2.x + cnt2.y);       // P_i: same-label columns (exact: small integers)
    // kappa_i = r_i c_i / (B N),  c_i = 1/(P_i - 1 + 1e-18)   (dycon_losses.py:192)
    const float kappa = rw / ((cnt - 1.f) + kTiny) * p.inv_rows;
    const float kx = kappa * p.kh;

    // ---- pass 2: the positives ----
    // (the label tests are repeated in every pass from the labels in shared memory; an opaque copy of y_i per pass keeps
    //  the compiler from carrying 56 masks from pass to pass in registers)
    // T = e + n is taken with n + 1e-30: the terms of EVERY pair stay finite (a padded column of a row without negatives
    // has e = n = 0), so that the masks may multiply instead of select -- the reference adds 1e-18 there (:186-187)
    const float n_t = n_i + 1e-30f;
    const float2 nt2 = make_float2(n_t, n_t), kx2 = make_float2(kx, kx);
    float yi2 = yi, yi3 = yi;
    asm volatile("" : "+f"(yi2));
    asm volatile("" : "+f"(yi3));
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const float4 ya = *reinterpret_cast<const float4*>(ys + k * 256 + lane * 4);
      const float4 yb = *reinterpret_cast<const float4*>(ys + k * 256 + 128 + lane * 4);
      const float2 yy[4] = {make_float2(ya.x, ya.y), make_float2(ya.z, ya.w), make_float2(yb.x, yb.y), make_float2(yb.z, yb.w)};
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        float2 phi2, at, px;
        pos_terms2<kFocal>(e[k * 4 + q2], nt2, p.gamma, phi2, at, px);
        const float2 same = make_float2(mask_eq(yy[q2].x, yi2), mask_eq(yy[q2].y, yi2));
        a0 = __ffma2_rn(same, phi2, a0);
        a1 = __ffma2_rn(same, at, a1);
        const float2 ev = e[k * 4 + q2];
        e[k * 4 + q2] = __ffma2_rn(same, __ffma2_rn(px, kx2, make_float2(-ev.x, -ev.y)), ev);    // positives: kappa h px; negatives keep e
      }
    }
    float acc0 = a0.x + a0.y, acc1 = a1.x + a1.y;
    {                                                   // the diagonal's terms: e_ii = 2^(S_ii c1 - m2_i)
      float phi2, at, px;
      pos_terms<kFocal>(ex2_approx(fmaf(__half2float(s_diag), p.c1, -(__ldcg(p.stat_m + off + i) * kLog2e))), n_t, p.gamma, phi2, at, px);
      acc0 -= fsel(diag_lane, phi2, 0.f);
      acc1 -= fsel(diag_lane, at, 0.f);
    }
    acc0 = warp_sum(acc0);
    acc1 = warp_sum(acc1);
    if (cnt <= 1.f) acc0 = acc1 = 0.f;                  // a row alone in its class has no positive pair: exactly 0 (kappa is ~1e18 there)
    const float fneg1 = -(kx * acc1) - 1.f;            // negatives: -kappa_i A_i h e_ij (as a factor 1 + mask (f - 1))
    const float2 fn2 = make_float2(fneg1, fneg1), one2 = make_float2(1.f, 1.f);

    // ---- pass 3: the negatives, and X over S ----
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const float4 ya = *reinterpret_cast<const float4*>(ys + k * 256 + lane * 4);
      const float4 yb = *reinterpret_cast<const float4*>(ys + k * 256 + 128 + lane * 4);
      const float2 yy[4] = {make_float2(ya.x, ya.y), make_float2(ya.z, ya.w), make_float2(yb.x, yb.y), make_float2(yb.z, yb.w)};
      uint32_t o4[4];
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        const float2 df = make_float2(mask_ne(yy[q2].x, yi3), mask_ne(yy[q2].y, yi3));
        const float2 xv = __fmul2_rn(e[k * 4 + q2], __ffma2_rn(df, fn2, one2));
        o4[q2] = Cvt<false>::two(xv.x, xv.y);
      }
      const int c0 = k * 256 + lane * 8;
      if (c0 < p.ncol) row[c0 >> 3] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    }
    if (diag_lane) reinterpret_cast<uint16_t*>(row)[i] = 0;        // same thread, after its vector store: program order
    if (lane == 0) {
      p.stat_n[off + i] = n_i;
      p.stat_kappa[off + i] = kappa;
      p.stat_a[off + i] = acc1;
      red[0] += (double)(kappa * (-kLn2 * acc0));
    }
  }

  DYCON_SPAN(2, tid == 0, 2);
  // ---- grid sum of the student term; the last block adds the cross sums of the sweep, runs the exchange of a
  //      sharded batch and writes the loss (same protocol as the tail of the loss sweep) ----
  double total_[1];
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (grid_sum_last_block<1>(red, total_, p.ticket, p.partials, nblocks, bid, rp_scratch, &rp_last) && warp == 0) {
    double t0 = __shfl_sync(0xffffffffu, total_[0], 0);
    double t1 = __ldcg(p.sums_out + 1), t2 = __ldcg(p.sums_out + 2);
    if (p.x.world > 1) {
      const double tot = exchange_warp(p.x, lane == 0 ? t0 : lane == 1 ? t1 : t2, 3);
      t0 = __shfl_sync(0xffffffffu, tot, 0);
      t1 = __shfl_sync(0xffffffffu, tot, 1);
      t2 = __shfl_sync(0xffffffffu, tot, 2);
    }
    if (lane == 0) {
      const double student = t0 / p.inv_rows_d;
      p.sums_out[0] = student;
      p.sums_out[1] = t1;
      p.sums_out[2] = t2;
      if (p.loss_out) {
        const double cross = p.has_teacher ? t1 / (t2 + 1e-18) : 0.0;
        *p.loss_out = (float)(student * p.inv_rows_d + (double)p.lambda_cross * cross);
      }
    }
  }
  DYCON_SPAN(2, tid == 0, 3);
}

// =================================================================================================
//  Backward
// =================================================================================================
// A CTA owns 128 rows of one sample and walks 32-column sub-tiles J of its column range:
//   MMA1: S = F_I F_J^T, CS = F_I T_J^T (128x32, K = D)   -> TMEM slot J & 1
//   epilogue: H = G + G^T, Gc (16-bit) -> column half J & 1 of the K-major smem tiles sH / sG
//   MMA2: dF_I += H F_J + Gc T_J  (B operands read MN-major from the very tiles that produced S / CS)
// TMA is a high-latency path (~1.5-3 us per tile when few are in flight), so the operand ring is four
// 32 KB stages deep, and the work is decoupled over independent agents that only meet on mbarriers:
//   warp 0 TMA producer | warps 1, 3 MMA1 issuers (even / odd sub-tiles) | warp 2 MMA2 issuer |
//   warps 4..11 epilogue team 0 (even sub-tiles) | warps 12..19 epilogue team 1 (odd sub-tiles)
// An epilogue thread owns one row and 16 of the 32 columns of its team's sub-tile (TMEM lane quadrant =
// warp % 4).  The two teams run out of phase: while one is in its MUFU-heavy arithmetic the other waits for
// TMEM or stores H.
constexpr int kBwdThreads = 640;
constexpr int kBwdTeamThreads = 256;
constexpr int kBwdStages = 4;
__device__ __forceinline__ void bwd_team_barrier(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kBwdTeamThreads) : "memory");
}

struct BwdParams {
  int N, Npad, KC, D, has_teacher;
  int splits;          // column splits; with > 1 the partial dF are summed with red.global.add onto zeros
  int pdl;             // the zero fill of grad_feat may still be running: griddepcontrol.wait before the first RED
  FeclScalars sc;
  float c1;
  const float* labels;
  const float* stat_m;
  const float* stat_n;
  const float* apart;  // P2's per-split partial A_i (hdr[1] planes of B*N floats)
  const float* stat_kappa;
  const float* hdr;
  const double* cross_cnt;
  const float* grad_out;
  float* grad_feat;
  int64_t g_sb, g_sn, g_sd;
  int rb_lo, row_lo, row_hi;   // row blocks / rows of this launch (see SweepParams)
  int grad_rows;               // rows per sample of grad_feat in global-negatives mode (0: row i of sample b)
  const int* cls_lo;           // rows sorted by label (nullptr: the caller's order): class bounds by row, and the
  const int* cls_hi;           // original row of every sorted position -- the gradient is scattered back through it
  const int* perm;
  int use_classes;             // specialise the pair arithmetic by the label class of a sub-tile
};

struct BwdMisc {
  uint64_t a_full[4];      // one per 16 KB K chunk of the A tile
  uint64_t b_full[kBwdStages], b_empty[kBwdStages];
  uint64_t sc_full[2], sc_empty[2];      // per team: S / CS accumulators of its sub-tile
  uint64_t h_full[2], h_free[2];         // per team: its column half of sH / sG
  uint64_t df_full;
  uint32_t tmem_slot;
  uint32_t pad_;
  alignas(16) float col[2][2][5][32];    // [team][slot][y, m2, n 2^-m2, -kappa A h 2^m2, kappa h][column]
};

// One pair (i, j) of a sub-tile, both directions at once (H = G + G^T is built from S_ij = S_ji):
//   e_ij = 2^(x c1 - m2_j) costs the ONE ex2 of the pair; the other direction, e_ji = 2^(x c1 - m2_i) =
//   e_ij 2^m2_j 2^-m2_i, only ever appears as d_ji = e_ji / (e_ji + n_j) and as P_j e_ji, so it is carried
//   divided by 2^m2_j:  ejs = e_ij ui (ui = 2^-m2_i, per row) against the column statistics njs = n_j 2^-m2_j and
//   Pjs = P_j 2^m2_j, which the publishing threads scale once per column (m2 clamped at 120 in these factors, far
//   above anything 16-bit operands can resolve).
// kCls: the label class of the sub-tile (uniform over the team) -- all pairs positive (SAME: the focal terms of
// both directions, no teacher term), all negative (DIFF: two FMAs, plus the teacher's 1/(1 - cs)), or MIXED
// (both, then a select).
template <int kFocal, bool kTeacher, int kCls>
__device__ __forceinline__ void bwd_pair(float x, float cs, float yi, float m2i, float ui, float ni, float Pi, float ki,
                                         float yj, float m2j, float njs, float Pjs, float kj, float c1, float gamma,
                                         float gl, float thresh, float gcs, float& h, float& g) {
  const float tij = fmaf(x, c1, -m2j);
  const float eij = ex2_approx(tij);                             // padded column: m2j = +inf -> eij = 0
  const float ejs = eij * ui;
  if (kCls == kClsDiff) {
    h = fmaf(Pi, eij, Pjs * ejs);
    g = 0.f;
    if (kTeacher) {
      // padded columns have cs == 0 exactly (zero teacher rows) and thresh >= 0 (host check): never hard
      const float om = (1.f - cs) + kTiny;                       // dycon_losses.py:228
      g = cs > thresh ? gcs * rcp_approx(om) : 0.f;
    }
    return;
  }
  const float Tij = eij + ni, Tjs = ejs + njs;
  const float p2 = Tij * Tjs;
  float r2, rom = 0.f;
  if (kTeacher && kCls == kClsMixed) {
    const float om = (1.f - cs) + kTiny;
    const float r = rcp_approx(p2 * om);                         // 1/T_ij, 1/T_ji, 1/(1-cs) from ONE reciprocal
    rom = r * p2;
    r2 = r * om;
  } else {
    r2 = rcp_approx(p2);
  }
  const float rij = r2 * Tjs, rjs = r2 * Tij;
  float pij, pji;                                                // phi'(d) d (1 - d)
  if (kFocal == kNoFocal) {
    pij = fmaf(eij, rij, -1.f);
    pji = fmaf(ejs, rjs, -1.f);
  } else {
    const float dij = eij * rij, dji = ejs * rjs;
    float oij = fmaf(-eij, rij, 1.f), oji = fmaf(-ejs, rjs, 1.f);
    const float Lij = tij - lg2_approx(Tij), Lji = (tij - m2i) - lg2_approx(Tjs);   // log2 d
    if (kFocal == kFocalG2) {
      const float aij = fmaf(gl, dij * Lij, -oij), aji = fmaf(gl, dji * Lji, -oji);
      pij = oij * oij * aij;
      pji = oji * oji * aji;
    } else {
      const float wij = pow_gm1(oij, gamma), wji = pow_gm1(oji, gamma);     // clamps oij / oji to >= 0
      const float aij = fmaf(gl, dij * Lij, -oij), aji = fmaf(gl, dji * Lji, -oji);
      pij = wij * oij * aij;
      pji = wji * oji * aji;
    }
  }
  const float gpos = fmaf(ki, pij, kj * pji);
  if (kCls == kClsSame) {
    h = gpos;
    g = 0.f;
    return;
  }
  const float gneg = fmaf(Pi, eij, Pjs * ejs);
  const bool same = yj == yi;                                    // NaN labels (padding) compare unequal
  h = same ? gpos : gneg;                                        // select: the unused branch may be NaN
  g = (kTeacher && !same && cs > thresh) ? gcs * rom : 0.f;
}

// Zero fill of the gradient for the split backward.  It releases its dependent at once: the backward is
// launched with programmatic stream serialization, runs its whole main loop next to this grid and only waits
// for it (griddepcontrol.wait) before its first red.global.add.
__global__ void __launch_bounds__(256) zero_fill_kernel(float* dst, size_t n, int pdl) {
  if (pdl) pdl_trigger();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (size_t q = tid; q < n / 4; q += nth) d4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t q = (n & ~(size_t)3) + tid; q < n; q += nth) dst[q] = 0.f;
  } else {
    for (size_t q = tid; q < n; q += nth) dst[q] = 0.f;
  }
}

template <bool kBf16, int kFocal>
__global__ void __launch_bounds__(kBwdThreads, 1)
fecl_tc_bwd_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapF,
                   const __grid_constant__ CUtensorMap mapT, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int KC = p.KC, Dpad = KC * 64;
  const uint32_t a_bytes = (uint32_t)KC * kChunk128, j_bytes = (uint32_t)KC * kChunk32;
  const uint32_t stage_bytes = j_bytes * 2;       // F_J sub-tile + T_J sub-tile, interleaved per K chunk
  uint8_t* const sA = smem;
  uint8_t* const sStage = smem + a_bytes;         // kBwdStages x stage_bytes
  uint8_t* const sH = sStage + kBwdStages * stage_bytes;     // [128 i][64 j] 16-bit, K-major SW128
  uint8_t* const sG = sH + kChunk128;
  BwdMisc& ms = *reinterpret_cast<BwdMisc*>(sG + kChunk128);
  const int b = blockIdx.z, i0 = (blockIdx.x + p.rb_lo) * kTM, split = blockIdx.y;
  const int nt_all = (p.N + 31) / 32;           // sub-tiles that hold at least one real column
  const int t0 = (int)((long long)split * nt_all / p.splits), t1 = (int)((long long)(split + 1) * nt_all / p.splits);
  const int nt = t1 - t0;                       // this CTA's 32-column sub-tiles: t0 .. t1-1
  const bool teacher = p.has_teacher != 0;
  // label class of every sub-tile (rows sorted by label): all-positive sub-tiles need no teacher term at all --
  // no Gc tile, no Gc T_J MMAs -- and every class has its own pair arithmetic (bwd_pair)
  RowClass rc{0, 0x7fffffff, 0};
  if (p.use_classes) rc = load_row_class(p.cls_lo, p.cls_hi, (size_t)b * p.N, i0, kTM, p.N);
  int ta, tb;
  same_range(rc, 32, ta, tb);
  const bool tl_on = blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 1 && lane == 0;     // (timeline build only)
  (void)tl_on;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int q = 0; q < 4; ++q) mbar_init(&ms.a_full[q], 1);
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&ms.b_full[s], 1);
      mbar_init(&ms.b_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&ms.sc_full[g], 1);
      mbar_init(&ms.sc_empty[g], kBwdTeamThreads / 32);
      mbar_init(&ms.h_full[g], kBwdTeamThreads / 32);
      mbar_init(&ms.h_free[g], 1);
    }
    mbar_init(&ms.df_full, 1);
    fence_mbar_init();
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapF);
    if (teacher) prefetch_tmap(&mapT);
  }
  if (warp == 1) tmem_alloc(&ms.tmem_slot, 512);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ms.tmem_slot, 0);
  // TMEM columns: team g's S | CS at g*64 (32 + 32 columns), dF at 256 (Dpad columns)
  const uint32_t tm_sc = tmem, tm_df = tmem + 256;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (nt > 0 && elect_one()) {
      auto load_a = [&](int c) {
        mbar_expect_tx(&ms.a_full[c], kChunk128);
        tma_load_2d(sA + c * kChunk128, &mapA, c * 64, b * p.Npad + i0, &ms.a_full[c]);
      };
      load_a(0);
      for (int t = 0; t < nt; ++t) {
        const int s = t % kBwdStages, row = b * p.Npad + (t0 + t) * 32;
        uint8_t* dst = sStage + s * stage_bytes;
        mbar_wait_relaxed(&ms.b_empty[s], ((t / kBwdStages) & 1) ^ 1);
        DYCON_TL(1, tl_on, 0, t, 0);
        mbar_expect_tx(&ms.b_full[s], teacher ? stage_bytes : j_bytes);
        for (int c = 0; c < KC; ++c) {     // chunk c = [32 F rows | 32 T rows] x 64 K: one N = 64 B operand for MMA1
          tma_load_2d(dst + c * kChunk64, &mapF, c * 64, row, &ms.b_full[s]);
          if (teacher) tma_load_2d(dst + c * kChunk64 + kChunk32, &mapT, c * 64, row, &ms.b_full[s]);
        }
        if (t == 0) {
          for (int c = 1; c < KC; ++c) load_a(c);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ================================ MMA1 issuers: S, CS (warp 1: team 0's sub-tiles, warp 3: team 1's) ====
    if (nt > 0 && elect_one()) {
      // one MMA per K step computes S | CS side by side (N = 64: the F and T rows of a chunk are contiguous)
      const uint32_t idesc_s = umma_idesc_16(128, teacher ? 64 : 32, false, false, kBf16);
      const uint64_t a_desc0 = umma_desc_kmajor(smem_u32(sA));
      for (int t = warp >> 1; t < nt; t += 2) {
        const bool first = t == (warp >> 1);      // this issuer's first sub-tile: the A chunks may still be in flight
        const int s = t % kBwdStages, g = t & 1;
        const uint64_t b_desc0 = umma_desc_kmajor(smem_u32(sStage + s * stage_bytes));
        mbar_wait_relaxed(&ms.b_full[s], (t / kBwdStages) & 1);
        DYCON_TL(1, tl_on, 1, t, 0);
        mbar_wait_relaxed(&ms.sc_empty[g], ((t >> 1) & 1) ^ 1);
        DYCON_TL(1, tl_on, 1, t, 1);
        tcgen05_after_sync();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < KC) {
            if (first) {
              mbar_wait_relaxed(&ms.a_full[c], 0);
              tcgen05_after_sync();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tm_sc + g * 64, desc_advance(a_desc0, c * kChunk128 + k * 32),
                        desc_advance(b_desc0, c * kChunk64 + k * 32), idesc_s, (c | k) != 0);
          }
        }
        umma_commit(&ms.sc_full[g]);
        DYCON_TL(1, tl_on, 1, t, 2);
      }
    }
  } else if (warp == 2) {
    // ================================ MMA2 issuer: dF += H F_J + Gc T_J ===========
    if (nt > 0 && elect_one()) {
      const uint32_t idesc_d = umma_idesc_16(128, Dpad, false, true, kBf16);   // B = F_J / T_J read MN-major
      const uint64_t h_desc0 = umma_desc_kmajor(smem_u32(sH)), g_desc0 = umma_desc_kmajor(smem_u32(sG));
      for (int t = 0; t < nt; ++t) {
        const int s = t % kBwdStages, g = t & 1;
        // MN-major B: 64-element MN blocks are the chunks (LBO = 8 KB), 8 K rows per 1 KB atom (SBO)
        const uint64_t f_desc0 = umma_desc(smem_u32(sStage + s * stage_bytes), kChunk64, 1024);
        mbar_wait_relaxed(&ms.h_full[g], (t >> 1) & 1);
        DYCON_TL(1, tl_on, 1, t, 3);
        tcgen05_after_sync();
        // K = the 32 columns of the sub-tile = 2 steps of 16: columns g*32 + k*16 of sH, rows k*16 (2048 B) of F_J
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16(tm_df, desc_advance(h_desc0, (g * 2 + k) * 32), desc_advance(f_desc0, k * 2048), idesc_d, (t | k) != 0);
        if (teacher && tile_class(rc, t0 + t, 32, p.N, ta, tb) != kClsSame) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(tm_df, desc_advance(g_desc0, (g * 2 + k) * 32), desc_advance(f_desc0, kChunk32 + k * 2048), idesc_d, true);
        }
        umma_commit(&ms.b_empty[s]);
        umma_commit(&ms.h_free[g]);
        DYCON_TL(1, tl_on, 1, t, 4);
      }
      umma_commit(&ms.df_full);
    }
  } else if (warp >= 4 && nt > 0) {
    // ================================ epilogue teams ==============================
    const int team = (warp - 4) >> 3;             // 0: even sub-tiles, 1: odd sub-tiles
    const int tt = threadIdx.x - 128 - team * kBwdTeamThreads;   // 0..255 inside the team
    const int quarter = warp & 3, chalf = ((warp - 4) >> 2) & 1;
    const int r = quarter * 32 + lane, i = i0 + r;
    const bool row_ok = i >= p.row_lo && i < p.row_hi;
    const size_t off = (size_t)b * p.N;
    const int ic = row_ok ? i : p.N - 1;
    const float qnan = __int_as_float(0x7fc00000);
    // H and Gc are scaled by a power of two (~ B N tau / 8, so |H| stays O(1)) before the 16-bit
    // conversion: exact, and it keeps fp16 out of its subnormal range.  Undone when dF is read out.
    const float hscale = __ldg(p.hdr);
    const float h_mul = p.sc.inv_tau * hscale;
    const float yi = row_ok ? __ldg(p.labels + off + ic) : qnan;
    const float m2i = __ldg(p.stat_m + off + ic) * kLog2e;
    const float ui = ex2_approx(-fminf(m2i, 120.f));          // e_ji / 2^m2_j = e_ij ui, see bwd_pair
    const float ni = __ldg(p.stat_n + off + ic);
    const float ki = row_ok ? __ldg(p.stat_kappa + off + ic) * h_mul : 0.f;
    const size_t bn = (size_t)gridDim.z * p.N;      // plane stride of the statistics
    const int splits2 = (int)__ldg(p.hdr + 1);
    auto a_of = [&](size_t g) {                     // A = P2's per-split partials, added in split order
      float part[kMaxSplits];
#pragma unroll
      for (int q = 0; q < kMaxSplits; ++q)          // all loads in flight at once (one L2 round trip, not eight)
        part[q] = q < splits2 ? __ldg(p.apart + (size_t)q * bn + g) : 0.f;
      float acc = part[0];
#pragma unroll
      for (int q = 1; q < kMaxSplits; ++q) acc += part[q];
      return acc;
    };
    const float Pi = -ki * a_of(off + ic);
    const float gcs = (teacher && row_ok) ? hscale * p.sc.lambda_cross / ((float)(*p.cross_cnt) + kTiny) : 0.f;
    const float gl = p.sc.gamma * kLn2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int cbase = chalf * 16;                 // this thread's 16 columns of the 32-column sub-tile
    const int hcol = team * 32 + cbase;           // ... and of the 64-column sH / sG tiles
    const bool tl_e = tl_on && (warp == 4 || warp == 12);
    (void)tl_e;

    // column statistics of sub-tile t: 5 planes x 32 columns, fetched by threads 0..159 of the team and
    // published through shared memory.  Padded columns: y = NaN, m2 = +inf, rest 0 (-> e_ij = 0, h = g = 0).
    auto fetch_one = [&](int t) -> float {
      if (tt >= 160) return 0.f;
      const int st = tt >> 5, j = t * 32 + (tt & 31);
      if (j >= p.N) return st == 0 ? qnan : st == 1 ? INFINITY : 0.f;
      const size_t g = off + j;
      if (st == 0) return __ldg(p.labels + g);
      const float m2 = __ldg(p.stat_m + g) * kLog2e;
      if (st == 1) return m2;
      if (st == 2) return __ldg(p.stat_n + g) * ex2_approx(-fminf(m2, 120.f));          // n_j 2^-m2_j
      const float kh = __ldg(p.stat_kappa + g) * h_mul;
      return st == 3 ? -kh * a_of(g) * ex2_approx(fminf(m2, 120.f)) : kh;              // P_j 2^m2_j | kappa_j h
    };
    auto publish = [&](int slot, float v) {
      if (tt < 160) ms.col[team][slot][tt >> 5][tt & 31] = v;
    };
    if (team < nt) publish(0, fetch_one(t0 + team));
    bwd_team_barrier(team);

    int it = 0;                                   // team-local iteration: sub-tile t = team + 2 * it
    for (int t = team; t < nt; t += 2, ++it) {
      const int slot = it & 1, j0 = (t0 + t) * 32;
      const float nx = fetch_one(t0 + (t + 2 < nt ? t + 2 : t));
      const int cls = tile_class(rc, t0 + t, 32, p.N, ta, tb);       // uniform over the team: no divergence
      const bool with_g = teacher && cls != kClsSame;                 // all-positive sub-tiles have no hard negatives
      DYCON_TL(1, tl_e, 2 + team, t, 0);
      mbar_wait(&ms.sc_full[team], it & 1);
      DYCON_TL(1, tl_e, 2 + team, t, 1);
      tcgen05_after_sync();
      float sv[16], cv[16];
      tmem_ld16(tm_sc + lane_base + team * 64 + cbase, sv);
      if (with_g) tmem_ld16(tm_sc + lane_base + team * 64 + 32 + cbase, cv);
      tmem_ld_wait();
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms.sc_empty[team]);

      const float* cy = &ms.col[team][slot][0][cbase];
      const float* cm = &ms.col[team][slot][1][cbase];
      const float* cn = &ms.col[team][slot][2][cbase];
      const float* cp = &ms.col[team][slot][3][cbase];
      const float* ck = &ms.col[team][slot][4][cbase];
      uint32_t hp[8], gp[8];
      auto pairs = [&](auto cls_tag) {
        constexpr int kCls = decltype(cls_tag)::value;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 m4 = *reinterpret_cast<const float4*>(cm + q * 4), p4 = *reinterpret_cast<const float4*>(cp + q * 4);
          const float m2[4] = {m4.x, m4.y, m4.z, m4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
          float ys[4] = {0.f, 0.f, 0.f, 0.f}, ns[4] = {0.f, 0.f, 0.f, 0.f}, ks[4] = {0.f, 0.f, 0.f, 0.f};
          if (kCls == kClsMixed) {
            const float4 y4 = *reinterpret_cast<const float4*>(cy + q * 4);
            ys[0] = y4.x; ys[1] = y4.y; ys[2] = y4.z; ys[3] = y4.w;
          }
          if (kCls != kClsDiff) {
            const float4 n4 = *reinterpret_cast<const float4*>(cn + q * 4), k4 = *reinterpret_cast<const float4*>(ck + q * 4);
            ns[0] = n4.x; ns[1] = n4.y; ns[2] = n4.z; ns[3] = n4.w;
            ks[0] = k4.x; ks[1] = k4.y; ks[2] = k4.z; ks[3] = k4.w;
          }
          float hv[4], gv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = q * 4 + k;
            if (teacher)
              bwd_pair<kFocal, true, kCls>(sv[c], cv[c], yi, m2i, ui, ni, Pi, ki, ys[k], m2[k], ns[k], ps[k], ks[k],
                                           p.c1, p.sc.gamma, gl, p.sc.cross_thresh, gcs, hv[k], gv[k]);
            else
              bwd_pair<kFocal, false, kCls>(sv[c], 0.f, yi, m2i, ui, ni, Pi, ki, ys[k], m2[k], ns[k], ps[k], ks[k],
                                            p.c1, p.sc.gamma, gl, p.sc.cross_thresh, gcs, hv[k], gv[k]);
          }
          hp[q * 2] = Cvt<kBf16>::two(hv[0], hv[1]);
          hp[q * 2 + 1] = Cvt<kBf16>::two(hv[2], hv[3]);
          gp[q * 2] = Cvt<kBf16>::two(gv[0], gv[1]);
          gp[q * 2 + 1] = Cvt<kBf16>::two(gv[2], gv[3]);
        }
      };
      if (cls == kClsSame) pairs(std::integral_constant<int, kClsSame>{});
      else if (cls == kClsDiff) pairs(std::integral_constant<int, kClsDiff>{});
      else pairs(std::integral_constant<int, kClsMixed>{});
      DYCON_TL(1, tl_e, 2 + team, t, 2);
      // the team's previous MMA2 must have finished reading this column half of sH / sG
      if (it > 0) mbar_wait(&ms.h_free[team], (it - 1) & 1);
      DYCON_TL(1, tl_e, 2 + team, t, 3);
#pragma unroll
      for (int u = 0; u < 2; ++u) {     // two 16-byte units = 16 16-bit columns
        const uint32_t o = sw128_offset(r, hcol + u * 8);
        *reinterpret_cast<uint4*>(sH + o) = make_uint4(hp[u * 4], hp[u * 4 + 1], hp[u * 4 + 2], hp[u * 4 + 3]);
        if (with_g) *reinterpret_cast<uint4*>(sG + o) = make_uint4(gp[u * 4], gp[u * 4 + 1], gp[u * 4 + 2], gp[u * 4 + 3]);
      }
      // the diagonal pair carries no gradient (l_ii is multiplied by 0, dycon_losses.py:176-178): clear it in
      // place instead of testing every pair.  (Its Gc is already 0: same label, never a hard negative.)
      const int rdiag = i - j0 - cbase;
      if (rdiag >= 0 && rdiag < 16) *reinterpret_cast<uint16_t*>(sH + sw128_offset(r, hcol + rdiag)) = 0;
      fence_async_smem();
      publish(slot ^ 1, nx);
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms.h_full[team]);
      bwd_team_barrier(team);
      DYCON_TL(1, tl_e, 2 + team, t, 4);
      if (tl_e) g_tl_cls(t, cls);
    }
    DYCON_TL(1, tl_e, 2 + team, 62, 0);

    // ---- dF (TMEM) * go -> grad_feat: the 16 epilogue warps split the Dpad columns four ways ----
    mbar_wait(&ms.df_full, 0);
    DYCON_TL(1, tl_e, 2 + team, 62, 1);
    tcgen05_after_sync();
    if (p.pdl) pdl_wait();                          // the zero fill of grad_feat is complete and visible
    const float go = __ldg(p.grad_out) / hscale;
    const int cgrp = (warp - 4) >> 2;               // 0..3
    const int dq = Dpad / 4;                         // columns per column group (multiple of 16)
    for (int c0 = cgrp * dq; c0 < (cgrp + 1) * dq; c0 += 16) {
      float v[16];
      tmem_ld16(tm_df + lane_base + c0, v);
      tmem_ld_wait();
      if (row_ok) {
        // global-negatives mode: row i of the merged batch is row (i - row_lo) % grad_rows of local sample
        // (i - row_lo) / grad_rows of this rank's gradient tensor
        const int il = i - p.row_lo;
        // sorted rows: position i of sample b is the caller's row perm[i]
        const int gb = p.grad_rows ? il / p.grad_rows : b;
        const int gn = p.grad_rows ? il - gb * p.grad_rows : p.perm ? __ldg(p.perm + off + i) : i;
        float* dst = p.grad_feat + (int64_t)gb * p.g_sb + (int64_t)gn * p.g_sn + (int64_t)c0 * p.g_sd;
        if (p.g_sd == 1 && ((p.g_sn | p.g_sb) & 3) == 0 &&
            (reinterpret_cast<uintptr_t>(p.grad_feat) & 15) == 0) {  // rows contiguous and 16-byte aligned: vector stores
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float o0 = go * v[q * 4], o1 = go * v[q * 4 + 1], o2 = go * v[q * 4 + 2], o3 = go * v[q * 4 + 3];
            if (c0 + q * 4 + 3 < p.D) {
              if (p.splits == 1) {
                *reinterpret_cast<float4*>(dst + q * 4) = make_float4(o0, o1, o2, o3);
              } else {   // two partial sums onto a zero-filled buffer: a + b == b + a, still bit-reproducible
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q * 4), "f"(o0), "f"(o1),
                             "f"(o2), "f"(o3) : "memory");
              }
            } else {
              const float o[4] = {o0, o1, o2, o3};
              for (int k = 0; k < 4; ++k) {
                if (c0 + q * 4 + k < p.D) {
                  if (p.splits == 1) dst[q * 4 + k] = o[k]; else atomicAdd(dst + q * 4 + k, o[k]);
                }
              }
            }
          }
        } else {   // columns contiguous (the caller's (D*N, 1, N) layout): a warp stores 32 consecutive rows
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c0 + c < p.D) {
              float* q = dst + (int64_t)c * p.g_sd;
              if (p.splits == 1) *q = go * v[c];
              else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(go * v[c]) : "memory");
            }
          }
        }
      }
    }
  }
  DYCON_TL(1, tl_on && warp == 4, 2, 63, 0);
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// =================================================================================================
//  Backward from stored pairs: a streaming GEMM with an elementwise fix-up
// =================================================================================================
// The loss sweep (kStore) has left X and Gc in the state: everything MUFU-heavy the recomputing backward evaluates a
// fourth time.  What it could not finish is the negative pairs' gradient -kappa_i A_i e_ij: A_i is complete only
// when the sweep is.  So a CTA (128 rows I, a column split) streams, per 64-column tile J,
//   X_IJ  [128 i][64 j]  K-major A           X_JI  [64 j][128 i]  the transposed direction, read as an MN-major A
//   Gc_IJ [128 i][64 j]  K-major A           F_J, T_J [64 j][Dpad]  MN-major B                       (all by TMA)
// the fix-up warps scale the negative pairs of both X tiles IN PLACE by their row's -kappa A h_mul / 256 (a pair is
// negative where the labels differ -- the test the sweep used), and one thread issues
//   dF1 += X_IJ F_J + X_JI^T F_J            dF2 += Gc_IJ T_J     (two 128 x Dpad fp32 accumulators = 512 TMEM columns)
// grad = go / hscale (dF1 + 64 lambda hscale / cnt dF2).  No similarity tile, no TMEM round trip per pair, no MUFU.
// Gc tiles without a hard negative (flagged by the sweep) are neither loaded nor multiplied.
constexpr int kGbThreads = 640;
constexpr int kGbFixThreads = 512;
constexpr int kGbStages = 2;

struct GbParams {
  int N, Npad, KC, D, has_teacher;
  int splits, pdl;
  int fixup;           // 0: X is final (written by the row kernel) -- a pure streaming GEMM
  float inv_tau, lambda_cross;
  const float* labels;
  const float* apart;
  const float* stat_kappa;
  const float* hdr;
  const double* cross_cnt;
  const unsigned int* gc_flag;
  const float* grad_out;
  float* grad_feat;
  int64_t g_sb, g_sn, g_sd;
};

struct GbMisc {
  uint64_t full[kGbStages], fixed[kGbStages], empty[kGbStages];
  uint64_t df_full;
  uint32_t tmem_slot;
  uint32_t pad_;
  alignas(16) float y_own[128];
  alignas(16) float s_own[128];
  alignas(16) float y_col[kGbStages][64];
  alignas(16) float s_col[kGbStages][64];
};

template <bool kBf16>
__device__ __forceinline__ void gb_unpack(uint32_t v, float& a, float& b) {
  if (kBf16) {
    a = __uint_as_float(v << 16);
    b = __uint_as_float(v & 0xffff0000u);
  } else {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
    a = f.x;
    b = f.y;
  }
}

// One 16-byte unit (8 pairs of one tile row): pairs whose labels differ are scaled by the row's factor.
template <bool kBf16>
__device__ __forceinline__ void gb_fix_unit(uint8_t* unit, float y_row, float s_row, const float* y8) {
  uint4 v = *reinterpret_cast<uint4*>(unit);
  const float4 ya = *reinterpret_cast<const float4*>(y8), yb = *reinterpret_cast<const float4*>(y8 + 4);
  const float ys[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float a, b;
    gb_unpack<kBf16>(w[q], a, b);
    a = ys[2 * q] == y_row ? a : a * s_row;
    b = ys[2 * q + 1] == y_row ? b : b * s_row;
    w[q] = Cvt<kBf16>::two(a, b);
  }
  *reinterpret_cast<uint4*>(unit) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool kBf16>
__global__ void __launch_bounds__(kGbThreads, 1)
fecl_tc_bwd_gemm_kernel(const __grid_constant__ CUtensorMap mapF, const __grid_constant__ CUtensorMap mapT,
                        const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapG,
                        const GbParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int KC = p.KC, Dpad = KC * 64;
  const uint32_t ft_bytes = (uint32_t)KC * kChunk64;                 // F_J (or T_J): [KC chunks][64 j][64 d]
  const uint32_t stage_bytes = 2 * ft_bytes + 3 * kChunk128;         // F_J | T_J | X_IJ | X_JI | Gc_IJ
  GbMisc& ms = *reinterpret_cast<GbMisc*>(smem + kGbStages * stage_bytes);
  const int b = blockIdx.z, i0 = blockIdx.x * kTM, split = blockIdx.y;
  const int nt_all = (p.N + 63) / 64;
  const int t0 = (int)((long long)split * nt_all / p.splits), t1 = (int)((long long)(split + 1) * nt_all / p.splits);
  const int nt = t1 - t0;
  const bool teacher = p.has_teacher != 0;
  const unsigned int* flags = p.gc_flag + ((size_t)b * (p.Npad >> 7) + blockIdx.x) * (p.Npad >> 6) + t0;
  const bool gtl_on = blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 1 && lane == 0;     // (timeline build only)
  (void)gtl_on;
  DYCON_SPAN(3, threadIdx.x == 0, 0);
  DYCON_SPAN_V(3, threadIdx.x == 0, 4, g_smid());

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < kGbStages; ++s) {
      mbar_init(&ms.full[s], 1);
      mbar_init(&ms.fixed[s], kGbFixThreads / 32);
      mbar_init(&ms.empty[s], 1);
    }
    mbar_init(&ms.df_full, 1);
    fence_mbar_init();
    prefetch_tmap(&mapF);
    prefetch_tmap(&mapX);
    if (teacher) prefetch_tmap(&mapT), prefetch_tmap(&mapG);
  }
  if (warp == 1) tmem_alloc(&ms.tmem_slot, 512);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, ms.tmem_slot, 0);
  const uint32_t tm_d1 = tmem, tm_d2 = tmem + 256;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (nt > 0 && elect_one()) {
      for (int t = 0; t < nt; ++t) {
        const int s = t & 1, j0 = (t0 + t) * 64, rowj = b * p.Npad + j0, rowi = b * p.Npad + i0;
        uint8_t* st = smem + s * stage_bytes;
        const bool gc_on = teacher && __ldg(flags + t) != 0u;
        mbar_wait_relaxed(&ms.empty[s], ((t >> 1) & 1) ^ 1);
        DYCON_TL(1, gtl_on, 0, t, 0);
        mbar_expect_tx(&ms.full[s], ft_bytes + 2 * kChunk128 + (gc_on ? ft_bytes + kChunk128 : 0u));
        uint8_t* sx = st + 2 * ft_bytes;
        tma_load_2d(sx, &mapX, j0, rowi, &ms.full[s]);                           // X_IJ rows i0 .. +63
        tma_load_2d(sx + kChunk64, &mapX, j0, rowi + 64, &ms.full[s]);           //      rows i0 + 64 .. +127
        tma_load_2d(sx + kChunk128, &mapX, i0, rowj, &ms.full[s]);               // X_JI columns i0 .. +63
        tma_load_2d(sx + kChunk128 + kChunk64, &mapX, i0 + 64, rowj, &ms.full[s]);
        for (int c = 0; c < KC; ++c) tma_load_2d(st + c * kChunk64, &mapF, c * 64, rowj, &ms.full[s]);
        if (gc_on) {
          tma_load_2d(sx + 2 * kChunk128, &mapG, j0, rowi, &ms.full[s]);
          tma_load_2d(sx + 2 * kChunk128 + kChunk64, &mapG, j0, rowi + 64, &ms.full[s]);
          for (int c = 0; c < KC; ++c) tma_load_2d(st + ft_bytes + c * kChunk64, &mapT, c * 64, rowj, &ms.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (nt > 0 && elect_one()) {
      const uint32_t idesc_k = umma_idesc_16(128, Dpad, false, true, kBf16);   // A K-major, B = F_J / T_J MN-major
      const uint32_t idesc_m = umma_idesc_16(128, Dpad, true, true, kBf16);    // A = X_JI MN-major (the transposed tile)
      bool d2_on = false;
      for (int t = 0; t < nt; ++t) {
        const int s = t & 1;
        const uint32_t st = smem_u32(smem + s * stage_bytes), sx = st + 2 * ft_bytes;
        const bool gc_on = teacher && __ldg(flags + t) != 0u;
        // MN-major: 64-element MN blocks are the chunks / boxes (LBO = 8 KB), 8 K rows per 1 KB atom (SBO)
        const uint64_t f_desc = umma_desc(st, kChunk64, 1024), t_desc = umma_desc(st + ft_bytes, kChunk64, 1024);
        const uint64_t xij_desc = umma_desc_kmajor(sx), xji_desc = umma_desc(sx + kChunk128, kChunk64, 1024),
                       gc_desc = umma_desc_kmajor(sx + 2 * kChunk128);
        mbar_wait_relaxed(&ms.full[s], (t >> 1) & 1);
        DYCON_TL(1, gtl_on, 1, t, 0);
        tcgen05_after_sync();
        if (gc_on) {             // needs no fix-up: runs while the fix-up warps work on the X tiles
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tm_d2, desc_advance(gc_desc, k * 32), desc_advance(t_desc, k * 2048), idesc_k, d2_on || k != 0);
          d2_on = true;
        }
        if (p.fixup) {
          mbar_wait_relaxed(&ms.fixed[s], (t >> 1) & 1);
          tcgen05_after_sync();
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm_d1, desc_advance(xij_desc, k * 32), desc_advance(f_desc, k * 2048), idesc_k, (t | k) != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm_d1, desc_advance(xji_desc, k * 2048), desc_advance(f_desc, k * 2048), idesc_m, true);
        umma_commit(&ms.empty[s]);
        DYCON_TL(1, gtl_on, 1, t, 2);
      }
      umma_commit(&ms.df_full);
    }
  } else if (warp >= 4 && nt > 0) {
    // ================================ fix-up warps, then the read-out =============
    const int tt = threadIdx.x - 128;             // 0..511
    const size_t off = (size_t)b * p.N;
    const float qnan = __int_as_float(0x7fc00000);
    const float hscale = __ldg(p.hdr);
    const float h_mul = p.inv_tau * hscale;
    const int splits2 = (int)__ldg(p.hdr + 1);
    const size_t bn = (size_t)gridDim.z * p.N;    // plane stride of the statistics
    // row j of the sample: its factor for the negative pairs, -kappa_j A_j h_mul / 256 (A = the sweep's per-split
    // partials, added in split order)
    auto row_factor = [&](int j) -> float {
      if (j >= p.N) return 0.f;
      float part[kMaxSplits];
#pragma unroll
      for (int q = 0; q < kMaxSplits; ++q) part[q] = q < splits2 ? __ldg(p.apart + (size_t)q * bn + off + j) : 0.f;
      float acc = part[0];
#pragma unroll
      for (int q = 1; q < kMaxSplits; ++q) acc += part[q];
      return -(__ldg(p.stat_kappa + off + j) * h_mul) * acc * (1.f / 256.f);
    };
    auto row_label = [&](int j) -> float { return j < p.N ? __ldg(p.labels + off + j) : qnan; };
    // column statistics of tile t: threads 128..191 fetch the labels, 192..255 the factors
    auto fetch = [&](int t) -> float {
      if (tt < 128 || tt >= 256) return 0.f;
      const int j = (t0 + t) * 64 + (tt & 63);
      return tt < 192 ? row_label(j) : row_factor(j);
    };
    auto publish = [&](int slot, float v) {
      if (tt >= 128 && tt < 192) ms.y_col[slot][tt & 63] = v;
      else if (tt >= 192 && tt < 256) ms.s_col[slot][tt & 63] = v;
    };
    if (p.fixup) {
    if (tt < 128) {
      ms.y_own[tt] = row_label(i0 + tt);
      ms.s_own[tt] = row_factor(i0 + tt);
    }
    publish(0, fetch(0));
    asm volatile("bar.sync 1, 512;" ::: "memory");
    }

    for (int t = 0; p.fixup && t < nt; ++t) {
      const int s = t & 1;
      const float nx = fetch(t + 1 < nt ? t + 1 : t);
      uint8_t* sx = smem + s * stage_bytes + 2 * ft_bytes;
      mbar_wait(&ms.full[s], (t >> 1) & 1);
      const int row = tt >> 3, uir = tt & 7;
      const uint32_t uo = (uint32_t)((uir ^ (row & 7)) << 4);       // (row + 64) & 7 == row & 7
      // X_IJ: tile row = row i of the CTA, columns = the tile's j
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = row + h * 64;
        gb_fix_unit<kBf16>(sx + h * kChunk64 + row * 128 + uo, ms.y_own[r], ms.s_own[r], &ms.y_col[s][uir * 8]);
      }
      // X_JI: tile row = column j of the tile, columns = the CTA's rows i (two boxes of 64)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        gb_fix_unit<kBf16>(sx + kChunk128 + h * kChunk64 + row * 128 + uo, ms.y_col[s][row], ms.s_col[s][row],
                           &ms.y_own[h * 64 + uir * 8]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms.fixed[s]);
      publish(s ^ 1, nx);
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }

    // ---- (dF1 + gcs dF2) * go -> grad_feat: the 16 warps split the Dpad columns four ways ----
    bool any_gc = false;
    if (teacher) {
      for (int t = lane; t < nt; t += 32) any_gc |= __ldg(flags + t) != 0u;
      any_gc = __any_sync(0xffffffffu, any_gc);
    }
    const int quarter = warp & 3, r = quarter * 32 + lane, i = i0 + r;
    const bool row_ok = i < p.N;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    mbar_wait(&ms.df_full, 0);
    DYCON_TL(1, gtl_on && warp == 4, 2, 62, 1);
    DYCON_SPAN(3, threadIdx.x == 128, 1);                        // accumulators complete
    tcgen05_after_sync();
    if (p.pdl) pdl_wait();                          // the zero fill of grad_feat is complete and visible
    DYCON_SPAN(3, threadIdx.x == 128, 2);
    const float go = __ldg(p.grad_out) / hscale;
    const float gcs = any_gc ? 64.f * hscale * p.lambda_cross / ((float)(*p.cross_cnt) + kTiny) : 0.f;
    const int cgrp = (warp - 4) >> 2;               // 0..3
    const int dq = Dpad / 4;                         // columns per column group (multiple of 16)
    for (int c0 = cgrp * dq; c0 < (cgrp + 1) * dq; c0 += 16) {
      float v[16];
      tmem_ld16(tm_d1 + lane_base + c0, v);
      if (any_gc) {
        float v2[16];
        tmem_ld16(tm_d2 + lane_base + c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = fmaf(gcs, v2[c], v[c]);
      } else {
        tmem_ld_wait();
      }
      if (row_ok) {
        float* dst = p.grad_feat + (int64_t)b * p.g_sb + (int64_t)i * p.g_sn + (int64_t)c0 * p.g_sd;
        if (p.g_sd == 1 && ((p.g_sn | p.g_sb) & 3) == 0 &&
            (reinterpret_cast<uintptr_t>(p.grad_feat) & 15) == 0) {  // rows contiguous and 16-byte aligned: vector stores
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float o0 = go * v[q * 4], o1 = go * v[q * 4 + 1], o2 = go * v[q * 4 + 2], o3 = go * v[q * 4 + 3];
            if (c0 + q * 4 + 3 < p.D) {
              if (p.splits == 1) {
                *reinterpret_cast<float4*>(dst + q * 4) = make_float4(o0, o1, o2, o3);
              } else {   // two partial sums onto a zero-filled buffer: a + b == b + a, still bit-reproducible
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q * 4), "f"(o0), "f"(o1),
                             "f"(o2), "f"(o3) : "memory");
              }
            } else {
              const float o[4] = {o0, o1, o2, o3};
              for (int k = 0; k < 4; ++k) {
                if (c0 + q * 4 + k < p.D) {
                  if (p.splits == 1) dst[q * 4 + k] = o[k]; else atomicAdd(dst + q * 4 + k, o[k]);
                }
              }
            }
          }
        } else {   // columns contiguous (the caller's (D*N, 1, N) layout): a warp stores 32 consecutive rows
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c0 + c < p.D) {
              float* q = dst + (int64_t)c * p.g_sd;
              if (p.splits == 1) *q = go * v[c];
              else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(go * v[c]) : "memory");
            }
          }
        }
      }
    }
  }
  DYCON_TL(1, gtl_on && warp == 4, 2, 63, 0);
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  DYCON_SPAN(3, threadIdx.x == 0, 3);
#ifdef DYCON_TIMELINE
  if (threadIdx.x == 0 && teacher) {         // work of this CTA: tiles, and how many of them carry a Gc tile
    int ng = 0;
    for (int t = 0; t < nt; ++t) ng += __ldg(flags + t) != 0u;
    DYCON_SPAN_V(3, true, 5, ng * 1000 + nt);
  }
#endif
}

// ---- host side -----------------------------------------------------------------------------------
struct TcState {
  float* hdr;
  void* F;
  void* T;
  float* stats;
  // rows sorted by label (fecl_rank_kernel): planes of B*N entries behind the statistics
  int* rank;
  int* perm;
  int* cls_lo;
  int* cls_hi;
  float* ys;
  float* rws;
  // stored pairs (see the stored-pairs backward): X, Gc and the tile flags behind everything else
  void* pair_x;
  void* pair_gc;
  unsigned int* gc_flag;
};
constexpr size_t kHdrBytes = 1024;   // keeps the operand arrays 1024-B aligned relative to the state base
inline int npad_of(int N) { return (N + 127) / 128 * 128; }
inline int dpad_of(int D) { return (D + 63) / 64 * 64; }
size_t operand_bytes(int B, int N, int D) { return align_up((size_t)B * npad_of(N) * dpad_of(D) * 2, 1024); }
// kNumStats planes of row statistics + kMaxSplits planes each of P1's partial n_i and P2's partial A_i + the six
// planes of the label sort
constexpr int kSortPlanes = 6;
size_t stats_bytes(int B, int N) {
  return align_up((size_t)(kNumStats + 2 * kMaxSplits + kSortPlanes) * B * N * sizeof(float), 128);
}
// Stored pairs: B * Npad^2 16-bit entries per matrix (X, and Gc with a teacher) + the tile flags.  On by default
// for the per-sample loss (never for a merged batch, whose transposed tiles live on other ranks, nor for rows sorted
// by label); DYCON_FECL_BWD=recompute selects the recomputing backward, DYCON_FECL_PAIR_GB caps the matrices (GiB).
bool getenv_sort_on() {
  static const bool on = [] {
    const char* e = getenv("DYCON_FECL_SORT");
    return e && e[0] == '1';
  }();
  return on;
}
size_t pair_matrix_bytes(int B, int N) { return align_up((size_t)B * npad_of(N) * npad_of(N) * 2, 1024); }
size_t pair_flag_bytes(int B, int N) { return align_up((size_t)B * (npad_of(N) / 128) * (npad_of(N) / 64) * 4, 1024); }
bool stored_pairs(int B, int N, int has_teacher) {
  static const bool off = [] {
    const char* e = getenv("DYCON_FECL_BWD");
    return e && strcmp(e, "recompute") == 0;
  }();
  static const double cap_gb = [] {
    const char* e = getenv("DYCON_FECL_PAIR_GB");
    return e ? atof(e) : 16.0;
  }();
  if (off || getenv_sort_on()) return false;
  return (double)pair_matrix_bytes(B, N) * (has_teacher ? 2 : 1) <= cap_gb * (double)(1ull << 30);
}
// Similarity sweep + row kernel + pure-GEMM backward (the default where pairs are stored and a row fits the row
// kernel's registers); DYCON_FECL_FWD=sweeps keeps the three-sweep forward with the fix-up backward.
bool fused_rows(int B, int N, int has_teacher) {
  static const bool off = [] {
    const char* e = getenv("DYCON_FECL_FWD");
    return e && strcmp(e, "sweeps") == 0;
  }();
  return !off && stored_pairs(B, N, has_teacher) && npad_of(N) <= 256 * kRowMaxChunks;
}
size_t pair_bytes(int B, int N, int has_teacher) {
  return stored_pairs(B, N, has_teacher) ? pair_matrix_bytes(B, N) * (has_teacher ? 2 : 1) + pair_flag_bytes(B, N) : 0;
}

TcState carve(void* state, int B, int N, int D, int has_teacher) {
  char* p = reinterpret_cast<char*>(state);
  TcState s;
  s.hdr = reinterpret_cast<float*>(p);
  p += kHdrBytes;
  s.F = p;
  p += operand_bytes(B, N, D);
  s.T = has_teacher ? p : nullptr;
  if (has_teacher) p += operand_bytes(B, N, D);
  s.stats = reinterpret_cast<float*>(p);
  const size_t plane = (size_t)B * N;
  int* sp = reinterpret_cast<int*>(s.stats + (size_t)(kNumStats + 2 * kMaxSplits) * plane);
  s.rank = sp;
  s.perm = sp + plane;
  s.cls_lo = sp + 2 * plane;
  s.cls_hi = sp + 3 * plane;
  s.ys = reinterpret_cast<float*>(sp + 4 * plane);
  s.rws = reinterpret_cast<float*>(sp + 5 * plane);
  char* q = reinterpret_cast<char*>(state) + kHdrBytes + operand_bytes(B, N, D) * (has_teacher ? 2 : 1) +
            align_up(stats_bytes(B, N), 1024);
  s.pair_x = q;
  s.pair_gc = has_teacher ? q + pair_matrix_bytes(B, N) : nullptr;
  s.gc_flag = reinterpret_cast<unsigned int*>(q + pair_matrix_bytes(B, N) * (has_teacher ? 2 : 1));
  return s;
}

// Rows packed sorted by label (DYCON_FECL_SORT=1; never for a merged batch -- global negatives exchange the row
// statistics between ranks by merged row index -- nor beyond the rank kernel's shared memory).  OFF by default: on
// the B200 the class-specialised bodies cut the executed instructions of the loss sweep by 28 % and of the backward
// by 20 % and shorten the arithmetic phase of a sub-tile by a third, but both kernels are bound by their MUFU + issue
// floor per SM and by the life time of an operand stage, not by those instructions, so the step only pays for the
// extra launch (138 vs 150 us; profiles/r2_fecl_timeline.md).  DYCON_FECL_CLASSES=0 sorts but runs the general
// pair arithmetic on every sub-tile.
constexpr int kMaxSortRows = 49152;
bool sort_rows(int N, bool merged) { return getenv_sort_on() && !merged && N <= kMaxSortRows; }
bool use_classes() {
  static const bool off = [] {
    const char* e = getenv("DYCON_FECL_CLASSES");
    return e && e[0] == '0';
  }();
  return !off;
}

// The kernels treat zero-padded columns as "cs == 0 <= thresh": a negative threshold would turn them into hard
// negatives.  The reference's schedule is sigmoid_rampup(.., 0.3, 0.5) >= 0.3 (dycon_losses.py:222).
int check_tc_thresh(const FeclProblem& p) {
  DYCON_REQUIRE(!p.has_teacher || p.sc.cross_thresh >= 0.f, DYCON_ERR_UNSUPPORTED,
                "FeCL tensor-core path: cross_thresh=%g < 0 (use precision fp32)", (double)p.sc.cross_thresh);
  return DYCON_OK;
}

int check_tc_shape(int B, int N, int D) {
  DYCON_REQUIRE(D % 4 == 0 && D <= 256, DYCON_ERR_UNSUPPORTED,
                "FeCL bf16: D=%d must be a multiple of 4 and <= 256 (the reference projection head has D=256)", D);
  DYCON_REQUIRE((long long)(npad_of(N) / 128) * B * 2 <= kMaxPartials && B <= 65535, DYCON_ERR_UNSUPPORTED,
                "FeCL tensor-core path: %d row blocks x B=%d exceeds %d CTAs", npad_of(N) / 128, B, kMaxPartials / 2);
  return DYCON_OK;
}

// Opt the kernel into the full 227 KB of shared memory per CTA (minus its static part).
template <typename K>
int set_smem(K kernel) {
  cudaFuncAttributes attr;
  DYCON_CUDA(cudaFuncGetAttributes(&attr, kernel));
  DYCON_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024 - (int)attr.sharedSizeBytes));
  return DYCON_OK;
}

// cudaFuncSetAttribute applies to the CURRENT device: run `f` once per device (a process may drive several GPUs).
template <typename F>
int once_per_device(F f) {
  static std::mutex mu;
  static uint64_t done = 0;      // bit d: device d is set up (devices >= 64 are set up on every call)
  int dev = 0;
  DYCON_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 64 && ((done >> dev) & 1)) return DYCON_OK;
  if (int rc = f()) return rc;
  if (dev < 64) done |= uint64_t(1) << dev;
  return DYCON_OK;
}

}  // namespace

size_t fecl_tc_state_bytes(int B, int N, int D, int has_teacher, bool pairs) {
  if (D % 4 != 0 || D > 256) return 0;
  return kHdrBytes + operand_bytes(B, N, D) * (has_teacher ? 2 : 1) + align_up(stats_bytes(B, N), 1024) +
         (pairs ? pair_bytes(B, N, has_teacher) : 0);
}
size_t fecl_tc_workspace_bytes(int, int, int) { return 16 + sizeof(double) * 3 * kMaxPartials; }
void fecl_tc_layout(int B, int N, int D, int has_teacher, size_t out[6]) {
  const size_t stats = kHdrBytes + operand_bytes(B, N, D) * (has_teacher ? 2 : 1), plane = (size_t)B * N * sizeof(float);
  out[0] = 0;
  out[1] = stats + kStatM * plane;
  out[2] = stats + kStatN * plane;
  out[3] = stats + kStatKappa * plane;
  out[4] = stats + (size_t)(kNumStats + kMaxSplits) * plane;
  out[5] = plane;
}

namespace {

// Power of two ~ B_global N tau / 8: |H| <= 2 (1 + gamma/e) r_max / (B N tau), so scaled entries are O(1).
// Column splits per row block: 2 when that still fits one wave of SMs (small batches), else 1.  Never more
// than 2, so every split reduction has at most two float contributors and stays order-independent.
int pick_splits(int row_blocks, int B) {
  static const int forced = [] {
    const char* e = getenv("DYCON_FECL_SPLITS");
    return e ? atoi(e) : 0;
  }();
  if (forced == 1 || forced == 2) return forced;
  return (long long)row_blocks * B * 2 <= sm_count() ? 2 : 1;
}

int focal_kind(const FeclScalars& sc) { return !sc.focal ? kNoFocal : sc.gamma == 2.f ? kFocalG2 : kFocalAny; }

float pick_hscale(double inv_rows, float inv_tau) {
  const double x = 1.0 / (inv_rows * (double)inv_tau * 8.0);
  int e = 0;
  if (x > 1.0) {
    double y = x;
    while (y >= 2.0 && e < 40) { y *= 0.5; ++e; }
  }
  return (float)(1ull << e);
}

template <bool kBf16>
int tc_fwd_impl(const FeclProblem& p, const FeclFwdArgs& a, cudaStream_t st) {
  using T16 = typename Cvt<kBf16>::type;
  const int B = p.B, N = p.N, D = p.D;
  const int Npad = npad_of(N), Dpad = dpad_of(D), KC = Dpad / 64;
  TcState s = carve(a.state, B, N, D, p.has_teacher);
  const size_t plane = (size_t)B * N;
  PackParams pk;
  pk.src[0] = a.feat; pk.sb[0] = a.f_sb; pk.sn[0] = a.f_sn; pk.sd[0] = a.f_sd; pk.dst[0] = s.F;
  pk.src[1] = a.teacher; pk.sb[1] = a.t_sb; pk.sn[1] = a.t_sn; pk.sd[1] = a.t_sd; pk.dst[1] = s.T;
  pk.stats = s.stats; pk.B = B; pk.N = N; pk.D = D; pk.Npad = Npad; pk.Dpad = Dpad;
  pk.merge = 0;
  pk.scale[0] = a.feat_scale; pk.scale[1] = a.teacher_scale;
  if (a.merge_B > 0) {       // global negatives: merge_B samples of N / merge_B rows -> one sample of N rows
    pk.B = a.merge_B; pk.N = N / a.merge_B; pk.Npad = npad_of(pk.N); pk.merge = 1;
  }
  const int row_lo = a.row_lo, row_hi = a.row_hi < 0 ? N : a.row_hi;
  // DYCON_NO_PDL=1 launches the forward kernels fully serialised and without any griddepcontrol instruction
  // (ncu's kernel replay of a --set full capture does not get along with programmatic dependent launches)
  static const bool no_pdl = [] {
    const char* e = getenv("DYCON_NO_PDL");
    return e && e[0] == '1';
  }();
  pk.pdl = no_pdl ? 0 : 1;
  const bool sorted = sort_rows(N, a.merge_B > 0);
  pk.rank = sorted ? s.rank : nullptr;
  // stored pairs: the loss sweep also leaves the pair terms of the backward in the state (tc_bwd_impl decides the same way)
  const bool store = !kBf16 && a.merge_B == 0 && row_lo == 0 && row_hi == N && !sorted && stored_pairs(B, N, p.has_teacher);
  pk.gc_flag = (store && p.has_teacher) ? s.gc_flag : nullptr;
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  dim3 pgrid(pk.Npad / 64, Dpad / 64, pk.B * (p.has_teacher ? 2 : 1));
  if (a.phase_mask & 1) {
    if (sorted) {
      RankParams rp;
      rp.labels = a.labels; rp.row_weight = a.row_weight; rp.N = N;
      rp.rank = s.rank; rp.perm = s.perm; rp.cls_lo = s.cls_lo; rp.cls_hi = s.cls_hi; rp.ys = s.ys; rp.rws = s.rws;
      rp.pdl = pk.pdl;
      const size_t rsmem = (size_t)((N + 3) & ~3) * sizeof(uint32_t);
      if (int rc = once_per_device([] {
            DYCON_CUDA(cudaFuncSetAttribute(fecl_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kMaxSortRows * (int)sizeof(uint32_t)));
            return (int)DYCON_OK;
          }))
        return rc;
      fecl_rank_kernel<<<dim3((N + 31) / 32, B), 256, rsmem, st>>>(rp);
      DYCON_CUDA(cudaGetLastError());
      count_launches(1);
    }
    cudaLaunchConfig_t pcfg = {};
    pcfg.gridDim = pgrid;
    pcfg.blockDim = dim3(256);
    pcfg.stream = st;
    pcfg.attrs = pdl_attr;
    pcfg.numAttrs = (sorted && pk.pdl) ? 1 : 0;      // the pack kernel overlaps its loads with the rank kernel
    DYCON_CUDA(cudaLaunchKernelEx(&pcfg, pack16_kernel<kBf16>, pk));
  }
  // boxes: 128 rows (A tile), 64 rows (a sub-tile of F alone), 32 rows (F | T interleaved in teacher mode)
  CUtensorMap mapA, mapF64, mapF32, mapT32, mapXs, mapGs;
  if (int rc = make_tmap_16_2d(&mapA, s.F, (uint64_t)B * Npad, Dpad, 128, kBf16)) return rc;
  if (int rc = make_tmap_16_2d(&mapF64, s.F, (uint64_t)B * Npad, Dpad, 64, kBf16)) return rc;
  if (int rc = make_tmap_16_2d(&mapF32, s.F, (uint64_t)B * Npad, Dpad, 32, kBf16)) return rc;
  if (int rc = make_tmap_16_2d(&mapT32, p.has_teacher ? s.T : s.F, (uint64_t)B * Npad, Dpad, 32, kBf16)) return rc;
  // store side of the pair matrices (dummies when nothing is stored: the kernels then never touch them)
  // (the similarity sweep stages per warp: 32-row boxes; the stored-pairs loss sweep per team: 128-row boxes)
  const uint32_t sbox = (!kBf16 && store && fused_rows(B, N, p.has_teacher)) ? 32u : 128u;
  if (int rc = store ? make_tmap_16_store(&mapXs, s.pair_x, (uint64_t)B * Npad, Npad, kBf16, sbox) : (mapXs = mapA, (int)DYCON_OK)) return rc;
  if (int rc = (store && p.has_teacher) ? make_tmap_16_store(&mapGs, s.pair_gc, (uint64_t)B * Npad, Npad, kBf16, sbox)
                                        : (mapGs = mapXs, (int)DYCON_OK))
    return rc;
  ReduceWorkspace ws = carve_reduce_workspace(a.workspace);
  SweepParams sp;
  sp.N = N; sp.Npad = Npad; sp.KC = KC; sp.has_teacher = p.has_teacher;
  // all three sweeps: 256-row tiles when the sample has at least two 128-row blocks, and as many column splits
  // (<= 8, at least one 64-column sub-tile each) as fit one wave of SMs
  static const int rt_env = [] {                      // DYCON_FECL_RT=1 | 2 forces the row tiles per CTA (experiments)
    const char* e = getenv("DYCON_FECL_RT");
    return e ? atoi(e) : 0;
  }();
  const int rt01 = rt_env == 1 ? 1 : Npad / 128 >= 2 ? 2 : 1;
  const int rb_lo = row_lo / (128 * rt01);                                   // row blocks that hold rows of this launch
  const int rb01 = (row_hi + 128 * rt01 - 1) / (128 * rt01) - rb_lo;
  int splits01 = sm_count() / (rb01 * B);
  if (splits01 > kMaxSplits) splits01 = kMaxSplits;
  if (splits01 > (N + 63) / 64) splits01 = (N + 63) / 64;
  if (splits01 < 1) splits01 = 1;
  sp.splits = splits01;
  sp.splits1 = splits01;
  sp.rb_lo = rb_lo; sp.row_lo = row_lo; sp.row_hi = row_hi;
  sp.npart = s.stats + (size_t)kNumStats * plane;
  sp.apart = s.stats + (size_t)(kNumStats + kMaxSplits) * plane;
  sp.sc = p.sc;
  sp.c1 = p.sc.inv_tau * kLog2e;
  sp.inv_rows = (float)p.inv_rows;
  sp.inv_rows_d = p.inv_rows;
  sp.hscale = pick_hscale(p.inv_rows, p.sc.inv_tau);
  sp.hdr = s.hdr;
  sp.labels = sorted ? s.ys : a.labels;
  sp.row_weight = a.row_weight ? (sorted ? s.rws : a.row_weight) : nullptr;
  sp.cls_lo = sorted ? s.cls_lo : nullptr;
  sp.cls_hi = sorted ? s.cls_hi : nullptr;
  sp.use_classes = sorted && use_classes();
  sp.stat_m = s.stats + kStatM * plane; sp.stat_n = s.stats + kStatN * plane;
  sp.stat_kappa = s.stats + kStatKappa * plane;
  sp.stat_p = s.stats + kStatP * plane;
  sp.ticket = ws.ticket; sp.partials = ws.partials; sp.sums_out = a.sums_out; sp.loss_out = a.loss_out;
  if (a.xc) sp.x = *a.xc; else make_exchange_ctx(&sp.x, nullptr, 0, 1, nullptr, DYCON_CHANNEL_FECL, 0.0);
  sp.pair_x = s.pair_x; sp.pair_gc = s.pair_gc; sp.gc_flag = s.gc_flag;
  // >= 120 KB of dynamic smem also pins one CTA per SM
  // (the stored-pairs sweep trades one ring stage for 32 KB of staging tiles: never more than this)
  size_t smem = rt01 == 2 ? (size_t)2 * KC * kChunk128 + (size_t)3 * KC * kChunk64 + sizeof(SweepMisc)
                          : (size_t)KC * kChunk128 + (size_t)kSwStages * KC * kChunk64 + sizeof(SweepMisc);
  if (store && KC < 4) smem += 32768;
  if (smem < 120 * 1024) smem = 120 * 1024;
  if (int rc = once_per_device([] {
        return set_smem(fecl_tc_sweep_kernel<0, kBf16, kNoFocal, 1>) | set_smem(fecl_tc_sweep_kernel<1, kBf16, kNoFocal, 1>) |
               set_smem(fecl_tc_sweep_kernel<2, kBf16, kNoFocal, 1>) | set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalG2, 1>) |
               set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalAny, 1>) | set_smem(fecl_tc_sweep_kernel<0, kBf16, kNoFocal, 2>) |
               set_smem(fecl_tc_sweep_kernel<1, kBf16, kNoFocal, 2>) | set_smem(fecl_tc_sweep_kernel<2, kBf16, kNoFocal, 2>) |
               set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalG2, 2>) | set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalAny, 2>) |
               set_smem(fecl_tc_sweep_kernel<2, kBf16, kNoFocal, 1, true>) | set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalG2, 1, true>) |
               set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalAny, 1, true>) | set_smem(fecl_tc_sweep_kernel<2, kBf16, kNoFocal, 2, true>) |
               set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalG2, 2, true>) | set_smem(fecl_tc_sweep_kernel<2, kBf16, kFocalAny, 2, true>) |
               set_smem(fecl_tc_sweep_kernel<3, kBf16, kNoFocal, 1>) | set_smem(fecl_tc_sweep_kernel<3, kBf16, kNoFocal, 2>);
      }))
    return rc;
  DYCON_REQUIRE(smem <= 227 * 1024, DYCON_ERR_UNSUPPORTED, "FeCL tensor-core fwd: %zu bytes of shared memory needed", smem);
  dim3 grid(rb01, splits01, B);
  const CUtensorMap& mapF2 = p.has_teacher ? mapF32 : mapF64;      // mode 2 walks 32-column sub-tiles with a teacher
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kSwThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = pdl_attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  sp.pdl = no_pdl ? 0 : 1;
  const int fk = focal_kind(p.sc);
  if (!kBf16 && store && fused_rows(B, N, p.has_teacher)) {
    // ---- similarity sweep (row max, S -> pair_x, the whole cross term) + row kernel (n, kappa, loss, A, X over S) ----
    if (rt01 == 2)
      DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_sweep_kernel<3, kBf16, kNoFocal, 2>, mapA, mapF2, mapT32, mapXs, mapGs, sp));
    else
      DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_sweep_kernel<3, kBf16, kNoFocal, 1>, mapA, mapF2, mapT32, mapXs, mapGs, sp));
    RowParams rp;
    rp.N = N; rp.Npad = Npad; rp.ncol = (N + 63) / 64 * 64;
    const int chunks = (Npad + 255) / 256;
    constexpr int kRowWarps = kRowThreads / 32;
    int want = kRowMinBlocks * sm_count() / B;           // CTAs per sample: the resident CTAs of every SM over the batch
    if (want < 1) want = 1;
    rp.rows_per_cta = ((rp.ncol + want - 1) / want + kRowWarps - 1) / kRowWarps * kRowWarps;
    rp.pdl = no_pdl ? 0 : 1;
    rp.c1 = sp.c1; rp.gamma = p.sc.gamma; rp.kh = p.sc.inv_tau * sp.hscale; rp.inv_rows = (float)p.inv_rows;
    rp.inv_rows_d = p.inv_rows; rp.lambda_cross = p.sc.lambda_cross; rp.has_teacher = p.has_teacher;
    rp.labels = a.labels; rp.row_weight = a.row_weight; rp.stat_m = sp.stat_m; rp.pair_x = s.pair_x;
    rp.stat_n = sp.stat_n; rp.stat_kappa = sp.stat_kappa; rp.stat_a = sp.apart;
    rp.ticket = ws.ticket; rp.partials = ws.partials; rp.sums_out = a.sums_out; rp.loss_out = a.loss_out;
    rp.x = sp.x;
    cudaLaunchConfig_t rcfg = {};
    rcfg.gridDim = dim3((rp.ncol + rp.rows_per_cta - 1) / rp.rows_per_cta, B);
    rcfg.blockDim = dim3(kRowThreads);
    rcfg.dynamicSmemBytes = (size_t)chunks * 256 * 2 * sizeof(float) + (size_t)(kRowThreads / 32) * chunks * 512;
    rcfg.stream = st;
    rcfg.attrs = pdl_attr;
    rcfg.numAttrs = no_pdl ? 0 : 1;
    DYCON_REQUIRE((long long)rcfg.gridDim.x * B <= kMaxPartials, DYCON_ERR_UNSUPPORTED, "FeCL row kernel: %u x %d CTAs", rcfg.gridDim.x, B);
#define DYCON_ROWK(FK, CH)                                                                                                 \
  do {                                                                                                                     \
    if (int rc = once_per_device([] {                                                                                      \
          DYCON_CUDA(cudaFuncSetAttribute(fecl_row_pairs_kernel<FK, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                          CH * 2048 + (kRowThreads / 32) * CH * 512));                                     \
          return (int)DYCON_OK;                                                                                            \
        }))                                                                                                                \
      return rc;                                                                                                           \
    DYCON_CUDA(cudaLaunchKernelEx(&rcfg, fecl_row_pairs_kernel<FK, CH>, rp));                                              \
  } while (0)
#define DYCON_ROWS(CH)                                                          \
  do {                                                                          \
    if (fk == kNoFocal) DYCON_ROWK(kNoFocal, CH);                               \
    else if (fk == kFocalG2) DYCON_ROWK(kFocalG2, CH);                          \
    else DYCON_ROWK(kFocalAny, CH);                                             \
  } while (0)
    switch (chunks) {
      case 1: DYCON_ROWS(1); break;
      case 2: DYCON_ROWS(2); break;
      case 3: DYCON_ROWS(3); break;
      case 4: DYCON_ROWS(4); break;
      case 5: DYCON_ROWS(5); break;
      case 6: DYCON_ROWS(6); break;
      case 7: DYCON_ROWS(7); break;
      default: DYCON_ROWS(8); break;
    }
#undef DYCON_ROWS
#undef DYCON_ROWK
    DYCON_CUDA(cudaGetLastError());
    count_launches(3);
    return DYCON_OK;
  }
#define DYCON_SWEEPS(RT, ST)                                                                                             \
  do {                                                                                                                   \
    if (a.phase_mask & 2)                                                                                                \
      DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_sweep_kernel<0, kBf16, kNoFocal, RT>, mapA, mapF64, mapT32, mapXs, mapGs, sp));       \
    if (a.phase_mask & 4)                                                                                                \
      DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_sweep_kernel<1, kBf16, kNoFocal, RT>, mapA, mapF64, mapT32, mapXs, mapGs, sp));       \
    if (!(a.phase_mask & 8)) break;                                                                                      \
    if (fk == kNoFocal)                                                                                                  \
      DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_sweep_kernel<2, kBf16, kNoFocal, RT, ST>, mapA, mapF2, mapT32, mapXs, mapGs, sp));    \
    else if (fk == kFocalG2)                                                                                             \
      DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_sweep_kernel<2, kBf16, kFocalG2, RT, ST>, mapA, mapF2, mapT32, mapXs, mapGs, sp));    \
    else                                                                                                                 \
      DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_sweep_kernel<2, kBf16, kFocalAny, RT, ST>, mapA, mapF2, mapT32, mapXs, mapGs, sp));   \
  } while (0)
  if (rt01 == 2) {
    if (store) DYCON_SWEEPS(2, true); else DYCON_SWEEPS(2, false);
  } else {
    if (store) DYCON_SWEEPS(1, true); else DYCON_SWEEPS(1, false);
  }
#undef DYCON_SWEEPS
  DYCON_CUDA(cudaGetLastError());
  count_launches((a.phase_mask & 1) + ((a.phase_mask >> 1) & 1) + ((a.phase_mask >> 2) & 1) + ((a.phase_mask >> 3) & 1));
  return DYCON_OK;
}

template <bool kBf16>
int tc_bwd_impl(const FeclProblem& p, const FeclBwdArgs& a, cudaStream_t st) {
  const int B = p.B, N = p.N, D = p.D;
  const int Npad = npad_of(N), Dpad = dpad_of(D), KC = Dpad / 64;
  TcState s = carve(const_cast<void*>(a.state), B, N, D, p.has_teacher);
  const size_t plane = (size_t)B * N;
  CUtensorMap mapA, mapF, mapT;
  if (int rc = make_tmap_16_2d(&mapA, s.F, (uint64_t)B * Npad, Dpad, 128, kBf16)) return rc;
  if (int rc = make_tmap_16_2d(&mapF, s.F, (uint64_t)B * Npad, Dpad, 32, kBf16)) return rc;
  if (int rc = make_tmap_16_2d(&mapT, p.has_teacher ? s.T : s.F, (uint64_t)B * Npad, Dpad, 32, kBf16)) return rc;
  BwdParams bp;
  bp.N = N; bp.Npad = Npad; bp.KC = KC; bp.D = D; bp.has_teacher = p.has_teacher;
  // column splits accumulate into a zero-filled gradient: only for a dense layout (memset of B*N*D floats)
  const int row_lo = a.row_lo, row_hi = a.row_hi < 0 ? N : a.row_hi;
  const int gn = a.grad_rows > 0 ? a.grad_rows : N;          // rows per sample of grad_feat
  const int rb_lo = row_lo / 128, rbs = (row_hi + 127) / 128 - rb_lo;
  const bool dense = a.g_sb == (int64_t)gn * D && ((a.g_sn == D && a.g_sd == 1) || (a.g_sn == 1 && a.g_sd == gn));
  bp.splits = dense ? pick_splits(rbs, B) : 1;
  static const bool no_pdl = [] {
    const char* e = getenv("DYCON_NO_PDL");
    return e && e[0] == '1';
  }();
  auto zero_fill = [&](int pdl) -> int {
    const size_t n = (size_t)B * (row_hi - row_lo) * D;
    const unsigned zgrid = (unsigned)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)sm_count() * 8);
    zero_fill_kernel<<<zgrid, 256, 0, st>>>(a.grad_feat, n, pdl);
    DYCON_CUDA(cudaGetLastError());
    count_launches(1);
    return DYCON_OK;
  };
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (!kBf16 && a.grad_rows == 0 && row_lo == 0 && row_hi == N && !sort_rows(N, false) && stored_pairs(B, N, p.has_teacher)) {
    // ---- stored pairs: streaming GEMM over what the loss sweep left in the state ----
    CUtensorMap mapF, mapT, mapX, mapG;
    if (int rc = make_tmap_16_2d(&mapF, s.F, (uint64_t)B * Npad, Dpad, 64, kBf16)) return rc;
    if (int rc = make_tmap_16_2d(&mapT, p.has_teacher ? s.T : s.F, (uint64_t)B * Npad, Dpad, 64, kBf16)) return rc;
    if (int rc = make_tmap_16_2d(&mapX, s.pair_x, (uint64_t)B * Npad, Npad, 64, kBf16)) return rc;
    if (int rc = make_tmap_16_2d(&mapG, p.has_teacher ? s.pair_gc : s.pair_x, (uint64_t)B * Npad, Npad, 64, kBf16)) return rc;
    GbParams gp;
    gp.N = N; gp.Npad = Npad; gp.KC = KC; gp.D = D; gp.has_teacher = p.has_teacher;
    gp.splits = bp.splits;
    gp.fixup = fused_rows(B, N, p.has_teacher) ? 0 : 1;
    gp.pdl = (gp.splits > 1 && !no_pdl) ? 1 : 0;
    gp.inv_tau = p.sc.inv_tau; gp.lambda_cross = p.sc.lambda_cross;
    gp.labels = a.labels;
    gp.apart = s.stats + (size_t)(kNumStats + kMaxSplits) * plane;
    gp.stat_kappa = s.stats + kStatKappa * plane;
    gp.hdr = s.hdr;
    gp.cross_cnt = a.cross_cnt; gp.gc_flag = s.gc_flag; gp.grad_out = a.grad_out; gp.grad_feat = a.grad_feat;
    gp.g_sb = a.g_sb; gp.g_sn = a.g_sn; gp.g_sd = a.g_sd;
    const size_t gsmem = (size_t)kGbStages * (2 * (size_t)KC * kChunk64 + 3 * kChunk128) + sizeof(GbMisc);
    if (int rc = once_per_device([] { return set_smem(fecl_tc_bwd_gemm_kernel<kBf16>); })) return rc;
    DYCON_REQUIRE(gsmem <= 227 * 1024, DYCON_ERR_UNSUPPORTED, "FeCL stored-pairs bwd: %zu bytes of shared memory needed", gsmem);
    if (gp.splits > 1) {
      if (int rc = zero_fill(gp.pdl)) return rc;
    }
    cudaLaunchConfig_t gcfg = {};
    gcfg.gridDim = dim3(Npad / 128, gp.splits, B);
    gcfg.blockDim = dim3(kGbThreads);
    gcfg.dynamicSmemBytes = gsmem < 120 * 1024 ? 120 * 1024 : gsmem;
    gcfg.stream = st;
    gcfg.attrs = pdl_attr;
    gcfg.numAttrs = gp.pdl ? 1 : 0;
    DYCON_CUDA(cudaLaunchKernelEx(&gcfg, fecl_tc_bwd_gemm_kernel<kBf16>, mapF, mapT, mapX, mapG, gp));
    DYCON_CUDA(cudaGetLastError());
    count_launches(1);
    return DYCON_OK;
  }
  bp.rb_lo = rb_lo; bp.row_lo = row_lo; bp.row_hi = row_hi; bp.grad_rows = a.grad_rows;
  bp.sc = p.sc;
  bp.c1 = p.sc.inv_tau * kLog2e;
  const bool sorted = sort_rows(N, a.grad_rows > 0);
  bp.labels = sorted ? s.ys : a.labels;
  bp.cls_lo = sorted ? s.cls_lo : nullptr;
  bp.cls_hi = sorted ? s.cls_hi : nullptr;
  bp.perm = sorted ? s.perm : nullptr;
  bp.use_classes = sorted && use_classes();
  bp.stat_m = s.stats + kStatM * plane; bp.stat_n = s.stats + kStatN * plane;
  bp.apart = s.stats + (size_t)(kNumStats + kMaxSplits) * plane;
  bp.stat_kappa = s.stats + kStatKappa * plane;
  bp.hdr = s.hdr;
  bp.cross_cnt = a.cross_cnt; bp.grad_out = a.grad_out; bp.grad_feat = a.grad_feat;
  bp.g_sb = a.g_sb; bp.g_sn = a.g_sn; bp.g_sd = a.g_sd;
  size_t smem = (size_t)KC * kChunk128 + (size_t)kBwdStages * 2 * KC * kChunk32 + 2 * kChunk128 + sizeof(BwdMisc);
  if (smem < 120 * 1024) smem = 120 * 1024;
  if (int rc = once_per_device([] {
        return set_smem(fecl_tc_bwd_kernel<kBf16, kNoFocal>) | set_smem(fecl_tc_bwd_kernel<kBf16, kFocalG2>) |
               set_smem(fecl_tc_bwd_kernel<kBf16, kFocalAny>);
      }))
    return rc;
  DYCON_REQUIRE(smem <= 227 * 1024, DYCON_ERR_UNSUPPORTED, "FeCL tensor-core bwd: %zu bytes of shared memory needed", smem);
  bp.pdl = 0;
  if (bp.splits > 1) {
    bp.pdl = no_pdl ? 0 : 1;
    if (int rc = zero_fill(bp.pdl)) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(rbs, bp.splits, B);
  cfg.blockDim = dim3(kBwdThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = pdl_attr;
  cfg.numAttrs = bp.pdl ? 1 : 0;
  switch (focal_kind(p.sc)) {
    case kNoFocal: DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_bwd_kernel<kBf16, kNoFocal>, mapA, mapF, mapT, bp)); break;
    case kFocalG2: DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_bwd_kernel<kBf16, kFocalG2>, mapA, mapF, mapT, bp)); break;
    default: DYCON_CUDA(cudaLaunchKernelEx(&cfg, fecl_tc_bwd_kernel<kBf16, kFocalAny>, mapA, mapF, mapT, bp)); break;
  }
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // namespace

// copies the timeline of the last launches to the host (timeline build only; returns 0 bytes otherwise)
size_t fecl_tc_debug_timeline(void* host_out, size_t bytes) {
#ifdef DYCON_TIMELINE
  const size_t n = sizeof(unsigned long long) * 2 * 4 * 64 * 8, n2 = sizeof(unsigned long long) * 4 * 512 * 6;
  if (bytes < n || cudaMemcpyFromSymbol(host_out, g_timeline, n) != cudaSuccess) return 0;
  if (bytes >= n + n2) {         // the per-CTA spans follow the timeline when the caller made room for them
    if (cudaMemcpyFromSymbol(static_cast<char*>(host_out) + n, g_span, n2) != cudaSuccess) return 0;
    return n + n2;
  }
  return n;
#else
  (void)host_out; (void)bytes;
  return 0;
#endif
}

int fecl_tc_fwd(const FeclProblem& p, const FeclFwdArgs& a, cudaStream_t st) {
  if (int rc = check_tc_shape(p.B, p.N, p.D)) return rc;
  if (int rc = check_tc_thresh(p)) return rc;
  return p.precision == DYCON_FECL_BF16 ? tc_fwd_impl<true>(p, a, st) : tc_fwd_impl<false>(p, a, st);
}

int fecl_tc_bwd(const FeclProblem& p, const FeclBwdArgs& a, cudaStream_t st) {
  if (int rc = check_tc_shape(p.B, p.N, p.D)) return rc;
  if (int rc = check_tc_thresh(p)) return rc;
  return p.precision == DYCON_FECL_BF16 ? tc_bwd_impl<true>(p, a, st) : tc_bwd_impl<false>(p, a, st);
}

}  // namespace dycon
