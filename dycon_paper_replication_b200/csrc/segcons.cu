// K4 -- the voxel-wise losses of the DyCON step loop in ONE pass over the logits.
//
// The reference step (code/train_DyCON_BraTS19.py:308-314,351-352) reads the same student / teacher logits four
// times through separate op chains:
//   u_loss        = UnCLoss(stud_logits, ema_logits, beta)                                dycon_losses.py:94-118
//   loss_seg      = F.cross_entropy(stud_logits[:lb], label[:lb])                         train_DyCON_BraTS19.py:313
//   loss_seg_dice = dice_loss(softmax(stud_logits)[:lb, 1], label[:lb] == 1)              utils/losses.py:8-16
//   consistency   = softmax_mse_loss(stud_probs[lb:], ema_probs[lb:]).mean()              utils/losses.py:65-82, :352
// (the last one is handed PROBABILITIES and takes the softmax again: a softmax of a softmax -- kept as is).
// Here one kernel computes the six sums all four need from a single read of s, t (and the labels of the labelled
// half), and one backward kernel re-reads them and writes d(sum_k go_k loss_k)/ds: 20-28 B/voxel forward,
// 28-36 B/voxel backward instead of the ~60 tensor passes of the op chains.  Two classes only (the scripts
// hard-set num_classes = 2, train_DyCON_BraTS19.py:146); the host wrapper composes the unfused kernels otherwise.
// Sums are two-stage and fixed-order (bit-reproducible, no float atomics).
#include "common.cuh"

namespace dycon {
namespace {

constexpr float kEps = 1e-6f;      // dycon_losses.py:95
constexpr float kSmooth = 1e-5f;   // utils/losses.py:10
constexpr int kThreads = 256;
enum { kSumUncl = 0, kSumCe = 1, kSumI = 2, kSumZ = 3, kSumY = 4, kSumCons = 5, kNumSums = 6 };

// Everything the four losses need from one voxel (two classes).  p1 = softmax(s)[1].
struct Voxel {
  float l_uncl;     // q W + beta (Hs + Ht)
  float du;         // d l_uncl / d s1   (d/ds0 = -du: softmax Jacobian with two classes)
  float p1, pp;     // student probability of class 1, p1 (1 - p1)
  float nlog0, nlog1;   // -log p0, -log p1 (cross entropy)
  float delta, qq;  // consistency: q_s1 - q_t1 with q = softmax(p)[1] = sigmoid(2 p1 - 1);  q_s1 (1 - q_s1)
};

struct Two { float p_hi, p_lo, l_hi, l_lo, H, a, l1p; bool hi1; };
__device__ __forceinline__ Two two_class(float x0, float x1) {
  Two o;
  const float d = x1 - x0;
  o.a = fabsf(d);
  o.hi1 = d >= 0.f;
  const float e = __expf(-o.a) + (o.a - o.a);          // +-inf logits -> NaN like torch.softmax
  o.l1p = log1pf(e);
  o.p_hi = 1.f / (1.f + e);
  o.p_lo = e * o.p_hi;
  o.l_hi = __logf(o.p_hi + kEps);
  o.l_lo = __logf(o.p_lo + kEps);
  o.H = -(o.p_hi * o.l_hi + o.p_lo * o.l_lo);
  return o;
}

__device__ __forceinline__ Voxel voxel(float s0, float s1, float t0, float t1, float beta) {
  const Two S = two_class(s0, s1), T = two_class(t0, t1);
  Voxel v;
  const float ps1 = S.hi1 ? S.p_hi : S.p_lo, pt1 = T.hi1 ? T.p_hi : T.p_lo;
  const float es = __expf(beta * S.H), et = __expf(beta * T.H);
  const float w = 1.f / (es + et);
  const float d = ps1 - pt1;
  const float q = 2.f * d * d;
  v.l_uncl = fmaf(q, w, beta * (S.H + T.H));
  // dL/dHs = beta (1 - q es W^2);  g_c = 2 (ps_c - pt_c) W - dL/dHs (log(ps_c + eps) + ps_c / (ps_c + eps))
  const float dl_dh = beta - q * beta * es * w * w;
  const float t_hi = S.l_hi + S.p_hi / (S.p_hi + kEps), t_lo = S.l_lo + S.p_lo / (S.p_lo + kEps);
  const float t1m0 = S.hi1 ? t_hi - t_lo : t_lo - t_hi;           // term of class 1 minus term of class 0
  v.pp = S.p_hi * S.p_lo;
  v.du = v.pp * fmaf(-dl_dh, t1m0, 4.f * d * w);
  v.p1 = ps1;
  // -log p_hi = log1p(e), -log p_lo = a + log1p(e)
  v.nlog1 = S.hi1 ? S.l1p : S.a + S.l1p;
  v.nlog0 = S.hi1 ? S.a + S.l1p : S.l1p;
  // consistency on the probabilities: softmax((p0, p1))[1] = sigmoid(p1 - p0) = sigmoid(2 p1 - 1)
  const float qs = 1.f / (1.f + __expf(1.f - 2.f * ps1)), qt = 1.f / (1.f + __expf(1.f - 2.f * pt1));
  v.delta = qs - qt;
  v.qq = qs * (1.f - qs);
  return v;
}

struct SegParams {
  const float* s;
  const float* t;
  const long long* label;      // (labeled_bs, V) int64 class indices, or nullptr
  int B, labeled_bs;
  long long V;
  float beta;
};

__global__ void __launch_bounds__(kThreads)
segcons_fwd_kernel(const SegParams p, unsigned int* ticket, double* partials, double* __restrict__ sums_out) {
  __shared__ double scratch[kNumSums * 32];
  float acc[kNumSums] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * kThreads;
  for (int b = blockIdx.y; b < p.B; b += gridDim.y) {
    const float* s0 = p.s + (2ll * b) * p.V;
    const float* t0 = p.t + (2ll * b) * p.V;
    const bool labelled = b < p.labeled_bs && p.label != nullptr;
    const long long* lab = p.label + (long long)b * p.V;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < p.V; i += stride) {
      const Voxel v = voxel(__ldcs(s0 + i), __ldcs(s0 + p.V + i), __ldcs(t0 + i), __ldcs(t0 + p.V + i), p.beta);
      acc[kSumUncl] += v.l_uncl;
      if (labelled) {
        const bool y = __ldcs(lab + i) == 1;
        acc[kSumCe] += y ? v.nlog1 : v.nlog0;
        acc[kSumI] += y ? v.p1 : 0.f;
        acc[kSumZ] += v.p1 * v.p1;
        acc[kSumY] += y ? 1.f : 0.f;
      } else if (b >= p.labeled_bs) {
        acc[kSumCons] += v.delta * v.delta;
      }
    }
  }
  double red[kNumSums], total[kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) red[k] = (double)acc[k];
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (grid_sum_last_block<kNumSums>(red, total, ticket, partials, nblocks, bid, scratch) && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) sums_out[k] = total[k];
  }
}

// losses_out[4] = {uncl, ce, dice, consistency} from the six sums (a second, one-thread kernel would cost a launch:
// the winner block of the forward calls this instead)
__device__ __forceinline__ void seg_losses(const double* sums, int B, int labeled_bs, long long V, float* out) {
  const double nl = (double)labeled_bs * (double)V, nu = (double)(B - labeled_bs) * (double)V;
  out[0] = (float)(sums[kSumUncl] / ((double)B * (double)V));
  out[1] = (float)(labeled_bs > 0 ? sums[kSumCe] / nl : 0.0);
  out[2] = (float)(1.0 - (2.0 * sums[kSumI] + (double)kSmooth) / (sums[kSumZ] + sums[kSumY] + (double)kSmooth));
  out[3] = (float)(B > labeled_bs ? sums[kSumCons] / nu : 0.0);      // mean over (B_u, 2, V) of the squared differences
}

__global__ void seg_losses_kernel(const double* sums, int B, int labeled_bs, long long V, float* out) {
  seg_losses(sums, B, labeled_bs, V, out);
}

// grad_s = go[0] dUnCL + go[1] dCE + go[2] dDice + go[3] dConsistency, recomputed per voxel.
__global__ void __launch_bounds__(kThreads)
segcons_bwd_kernel(const SegParams p, const double* __restrict__ sums, const float* __restrict__ go,
                   float* __restrict__ grad_s) {
  const float g_u = __ldg(go) / ((float)p.B * (float)p.V);
  const float nl = (float)p.labeled_bs * (float)p.V, nu = (float)(p.B - p.labeled_bs) * (float)p.V;
  const float g_ce = p.labeled_bs > 0 ? __ldg(go + 1) / nl : 0.f;
  // dice = 1 - Nn / Dn, Nn = 2 I + eps, Dn = Z + Y + eps:  d dice / d p_v = -(2 y_v Dn - 2 Nn p_v) / Dn^2
  const double Nn = 2.0 * sums[kSumI] + (double)kSmooth, Dn = sums[kSumZ] + sums[kSumY] + (double)kSmooth;
  const float g_dy = (float)(-2.0 * (double)__ldg(go + 2) / Dn), g_dp = (float)(2.0 * (double)__ldg(go + 2) * Nn / (Dn * Dn));
  const float g_c = p.B > p.labeled_bs ? 4.f * __ldg(go + 3) / nu : 0.f;      // d(delta^2)/dp1 = 2 delta * 2 q (1 - q)
  const long long stride = (long long)gridDim.x * kThreads;
  for (int b = blockIdx.y; b < p.B; b += gridDim.y) {
    const float* s0 = p.s + (2ll * b) * p.V;
    const float* t0 = p.t + (2ll * b) * p.V;
    float* g0 = grad_s + (2ll * b) * p.V;
    const bool labelled = b < p.labeled_bs && p.label != nullptr;
    const long long* lab = p.label + (long long)b * p.V;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < p.V; i += stride) {
      const Voxel v = voxel(__ldcs(s0 + i), __ldcs(s0 + p.V + i), __ldcs(t0 + i), __ldcs(t0 + p.V + i), p.beta);
      float g1 = g_u * v.du;
      if (labelled) {
        const float y = __ldcs(lab + i) == 1 ? 1.f : 0.f;
        g1 = fmaf(g_ce, v.p1 - y, g1);                                   // d(-log p_label)/ds1 = p1 - y
        g1 = fmaf(fmaf(g_dy, y, g_dp * v.p1), v.pp, g1);                 // dice through dp1/ds1 = p1 (1 - p1)
      } else if (b >= p.labeled_bs) {
        g1 = fmaf(g_c * v.delta * v.qq, v.pp, g1);
      }
      __stcs(g0 + i, -g1);
      __stcs(g0 + p.V + i, g1);
    }
  }
}

dim3 seg_grid(int B, long long V) {
  long long gy = B < 64 ? B : 64;
  long long gx = (V + kThreads - 1) / kThreads;
  long long cap = (long long)sm_count() * 8 / gy;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  while (gx * gy * kNumSums > (long long)kMaxPartials * kNumSums) --gx;       // partial slots: kNumSums per block
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

int seg_check(const float* s, const float* t, const long long* label, int B, int labeled_bs, int C, long long V) {
  DYCON_REQUIRE(s && t, DYCON_ERR_ARG, "segcons: NULL logits");
  DYCON_REQUIRE(B > 0 && V > 0 && labeled_bs >= 0 && labeled_bs <= B, DYCON_ERR_ARG,
                "segcons: B=%d labeled_bs=%d V=%lld", B, labeled_bs, (long long)V);
  DYCON_REQUIRE(C == 2, DYCON_ERR_UNSUPPORTED, "segcons: the fused step losses are built for 2 classes (got C=%d)", C);
  DYCON_REQUIRE(labeled_bs == 0 || label, DYCON_ERR_ARG, "segcons: labeled_bs=%d needs labels", labeled_bs);
  DYCON_REQUIRE(aligned(s, 4) && aligned(t, 4) && aligned(label, 8), DYCON_ERR_ARG, "segcons: misaligned pointer");
  return DYCON_OK;
}

}  // namespace
}  // namespace dycon

using namespace dycon;

extern "C" {

size_t dycon_segcons_workspace_bytes(void) { return 16 + sizeof(double) * kNumSums * kMaxPartials; }

int dycon_segcons_fwd(const float* s, const float* t, const long long* label, int B, int labeled_bs, int C, int64_t V,
                      float beta, double* sums_out, float* losses_out, void* workspace, size_t workspace_bytes,
                      dycon_stream_t stream) {
  if (int rc = seg_check(s, t, label, B, labeled_bs, C, V)) return rc;
  DYCON_REQUIRE(sums_out && losses_out && workspace && aligned(sums_out, 8) && aligned(workspace, 16), DYCON_ERR_ARG,
                "segcons fwd: NULL / misaligned sums_out / losses_out / workspace");
  DYCON_REQUIRE(workspace_bytes >= dycon_segcons_workspace_bytes(), DYCON_ERR_WORKSPACE, "segcons fwd: workspace %zu < %zu bytes",
                workspace_bytes, dycon_segcons_workspace_bytes());
  ReduceWorkspace ws = carve_reduce_workspace(workspace);
  SegParams p{s, t, label, B, labeled_bs, (long long)V, beta};
  cudaStream_t st = as_stream(stream);
  segcons_fwd_kernel<<<seg_grid(B, V), kThreads, 0, st>>>(p, ws.ticket, ws.partials, sums_out);
  seg_losses_kernel<<<1, 1, 0, st>>>(sums_out, B, labeled_bs, (long long)V, losses_out);
  DYCON_CUDA(cudaGetLastError());
  count_launches(2);
  return DYCON_OK;
}

int dycon_segcons_bwd(const float* s, const float* t, const long long* label, int B, int labeled_bs, int C, int64_t V,
                      float beta, const double* sums, const float* grad_out4, float* grad_s, dycon_stream_t stream) {
  if (int rc = seg_check(s, t, label, B, labeled_bs, C, V)) return rc;
  DYCON_REQUIRE(sums && grad_out4 && grad_s && aligned(grad_s, 4), DYCON_ERR_ARG, "segcons bwd: NULL sums / grad_out / grad_s");
  SegParams p{s, t, label, B, labeled_bs, (long long)V, beta};
  segcons_bwd_kernel<<<seg_grid(B, V), kThreads, 0, as_stream(stream)>>>(p, sums, grad_out4, grad_s);
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // extern "C"
