// K5 -- gradient clipping + SGD-momentum step + mean-teacher EMA as two multi-tensor launches, and the device-side
// finite check that lets the step loop drop its host synchronisations.
//
// The reference step touches every parameter tensor three times through separate per-tensor launches
// (code/train_DyCON_BraTS19.py:366-372):
//   torch.nn.utils.clip_grad_norm_(params, max_norm=1.0)     total L2 norm of all gradients, grads *= min(1, c/(n+1e-6))
//   optimizer.step()                                         SGD(lr, momentum=0.9, weight_decay=1e-4), :268
//   update_ema_variables(model, ema_model, 0.99, iter_num)   :155-164
// and decides on the host whether to run them at all (`if isnan(loss) or isinf(loss): continue`, :360-362: a
// device->host sync per step).  Here:
//   dycon_grad_norm     one launch over all gradient tensors: fixed-order two-stage sum of squares -> the norm and
//                       the clip coefficient, both left on the device
//   dycon_sgd_ema_step  one launch over all tensors: reads p, g, momentum buffer, teacher; writes p, buffer, teacher
//                       (28 B/param instead of ~60), with the rounding order of the PyTorch ops it replaces:
//                         g' = rn(g c);  g' = fma(wd, p, g');  buf = rn(rn(buf mu) + g')  (first step: buf = g');
//                         p = fma(-lr, buf, p);  ema = fma(1 - a, p, rn(ema a))
//                       `skip` (device int, may be NULL): non-zero -> the whole step is a no-op, which is what the
//                       reference's `continue` does for a non-finite loss -- without the host round trip.
//   dycon_finite_check  flag = any value non-finite (one thread; also counts the skipped steps)
#include "common.cuh"

namespace dycon {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 1024;        // elements per chunk: 256 threads x one float4
constexpr int kMaxTensors = 160;    // per launch: 160 * 44 B = 7 KB of kernel parameters

struct StepTable {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];       // nullptr: the tensor has no gradient (the teacher still follows it)
  float* buf[kMaxTensors];
  float* ema[kMaxTensors];           // nullptr: no teacher copy of this tensor
  long long numel[kMaxTensors];
  int chunk_start[kMaxTensors + 1];
  int n;
};

struct StepScalars {
  float lr, momentum, weight_decay, alpha, one_minus_alpha;
  int first;          // momentum buffers are uninitialised: buf = g'  (torch.optim.SGD's first step)
  int nesterov;
  int scale_grads;    // also write the clipped gradient back, as clip_grad_norm_ does
};

__device__ __forceinline__ int tensor_of_chunk(const StepTable& tab, int c) {
  int lo = 0, hi = tab.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.chunk_start[mid] <= c) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- total gradient norm (clip_grad_norm_: norm_type 2 over all tensors) ----------------------------------
__global__ void __launch_bounds__(kThreads)
grad_norm_kernel(const __grid_constant__ StepTable tab, int total_chunks, float max_norm, unsigned int* ticket,
                 double* partials, float* __restrict__ out /* {norm, clip coefficient} */) {
  __shared__ double scratch[32];
  const int per = total_chunks / (int)gridDim.x, rem = total_chunks % (int)gridDim.x;
  const int bx = (int)blockIdx.x;
  const int c_begin = bx * per + (bx < rem ? bx : rem), c_end = c_begin + per + (bx < rem ? 1 : 0);
  float acc = 0.f;
  if (c_begin < c_end) {
    int k = tensor_of_chunk(tab, c_begin);
    for (int c = c_begin; c < c_end; ++c) {
      while (tab.chunk_start[k + 1] <= c) ++k;
      const float* g = tab.g[k];
      if (g == nullptr) continue;
      const long long base = (long long)(c - tab.chunk_start[k]) * kChunk, left = tab.numel[k] - base;
      if (left >= kChunk && (reinterpret_cast<uintptr_t>(g + base) & 15) == 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + base) + threadIdx.x);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      } else {
        const long long n = left < kChunk ? left : kChunk;
        for (long long i = threadIdx.x; i < n; i += kThreads) acc += g[base + i] * g[base + i];
      }
    }
  }
  double v[1] = {(double)acc}, total[1];
  if (grid_sum_last_block<1>(v, total, ticket, partials, gridDim.x, blockIdx.x, scratch) && threadIdx.x == 0) {
    const float norm = (float)sqrt(total[0]);
    out[0] = norm;
    out[1] = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.f) : 1.f;      // clip_grad_norm_: clamp(c/(n+1e-6), max=1)
  }
}

__device__ __forceinline__ void step_one(float& p, float g, float& buf, float& ema, bool has_g, bool has_ema, float coef,
                                         const StepScalars& sc) {
  if (has_g) {
    float gg = __fmul_rn(g, coef);
    if (sc.weight_decay != 0.f) gg = fmaf(sc.weight_decay, p, gg);
    float d = gg;
    if (sc.momentum != 0.f) {
      buf = sc.first ? gg : __fadd_rn(__fmul_rn(buf, sc.momentum), gg);
      d = sc.nesterov ? fmaf(sc.momentum, buf, gg) : buf;
    }
    p = fmaf(-sc.lr, d, p);
  }
  if (has_ema) ema = fmaf(sc.one_minus_alpha, p, __fmul_rn(ema, sc.alpha));
}

__global__ void __launch_bounds__(kThreads)
sgd_ema_kernel(const __grid_constant__ StepTable tab, int total_chunks, const StepScalars sc,
               const float* __restrict__ clip /* {norm, coefficient} or nullptr */, const int* __restrict__ skip) {
  if (skip != nullptr && __ldg(skip) != 0) return;      // non-finite loss: the reference `continue`s past the whole step
  const float coef = clip ? __ldg(clip + 1) : 1.f;
  const int per = total_chunks / (int)gridDim.x, rem = total_chunks % (int)gridDim.x;
  const int bx = (int)blockIdx.x;
  const int c_begin = bx * per + (bx < rem ? bx : rem), c_end = c_begin + per + (bx < rem ? 1 : 0);
  if (c_begin >= c_end) return;
  int k = tensor_of_chunk(tab, c_begin);
  for (int c = c_begin; c < c_end; ++c) {
    while (tab.chunk_start[k + 1] <= c) ++k;
    const long long base = (long long)(c - tab.chunk_start[k]) * kChunk, left = tab.numel[k] - base;
    float* p = tab.p[k] + base;
    float* g = const_cast<float*>(tab.g[k]);
    float* buf = tab.buf[k];
    float* ema = tab.ema[k];
    const bool has_g = g != nullptr, has_ema = ema != nullptr, has_buf = has_g && buf != nullptr && sc.momentum != 0.f;
    if (has_g) g += base;
    if (has_buf) buf += base;
    if (has_ema) ema += base;
    uintptr_t bits = reinterpret_cast<uintptr_t>(p);
    if (has_g) bits |= reinterpret_cast<uintptr_t>(g);
    if (has_buf) bits |= reinterpret_cast<uintptr_t>(buf);
    if (has_ema) bits |= reinterpret_cast<uintptr_t>(ema);
    if (left >= kChunk && (bits & 15) == 0) {
      const int t = threadIdx.x;
      float4 pv = reinterpret_cast<float4*>(p)[t];
      float4 gv = has_g ? __ldcs(reinterpret_cast<const float4*>(g) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 bv = (has_buf && !sc.first) ? reinterpret_cast<float4*>(buf)[t] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 ev = has_ema ? reinterpret_cast<float4*>(ema)[t] : make_float4(0.f, 0.f, 0.f, 0.f);
      step_one(pv.x, gv.x, bv.x, ev.x, has_g, has_ema, coef, sc);
      step_one(pv.y, gv.y, bv.y, ev.y, has_g, has_ema, coef, sc);
      step_one(pv.z, gv.z, bv.z, ev.z, has_g, has_ema, coef, sc);
      step_one(pv.w, gv.w, bv.w, ev.w, has_g, has_ema, coef, sc);
      if (has_g) reinterpret_cast<float4*>(p)[t] = pv;
      if (has_buf) reinterpret_cast<float4*>(buf)[t] = bv;
      if (has_ema) reinterpret_cast<float4*>(ema)[t] = ev;
      if (has_g && sc.scale_grads)
        reinterpret_cast<float4*>(g)[t] = make_float4(__fmul_rn(gv.x, coef), __fmul_rn(gv.y, coef), __fmul_rn(gv.z, coef),
                                                      __fmul_rn(gv.w, coef));
    } else {
      const long long n = left < kChunk ? left : kChunk;
      for (long long i = threadIdx.x; i < n; i += kThreads) {
        float pv = p[i], gv = has_g ? g[i] : 0.f, bv = (has_buf && !sc.first) ? buf[i] : 0.f, ev = has_ema ? ema[i] : 0.f;
        step_one(pv, gv, bv, ev, has_g, has_ema, coef, sc);
        if (has_g) p[i] = pv;
        if (has_buf) buf[i] = bv;
        if (has_ema) ema[i] = ev;
        if (has_g && sc.scale_grads) g[i] = __fmul_rn(gv, coef);
      }
    }
  }
}

struct FiniteTable { const float* v[16]; };
__global__ void finite_check_table_kernel(const __grid_constant__ FiniteTable tab, int n, int* flag, unsigned long long* skipped) {
  int bad = 0;
  for (int k = 0; k < n; ++k) bad |= !(fabsf(*tab.v[k]) <= 3.402823466e38f);
  *flag = bad;
  if (bad && skipped) *skipped += 1ull;
}

// builds launch tables of at most kMaxTensors tensors and calls `launch(table, chunks)` for each
template <typename F>
int for_each_table(float* const* p, const float* const* g, float* const* buf, float* const* ema, const int64_t* numels,
                   int n_tensors, F launch) {
  int k = 0;
  while (k < n_tensors) {
    StepTable tab;
    tab.n = 0;
    long long chunks = 0;
    while (k < n_tensors && tab.n < kMaxTensors) {
      const long long nb = (numels[k] + kChunk - 1) / kChunk;
      if (nb == 0) { ++k; continue; }
      if (chunks + nb > 0x7fffffffLL / 2) break;
      tab.p[tab.n] = p ? p[k] : nullptr;
      tab.g[tab.n] = g ? g[k] : nullptr;
      tab.buf[tab.n] = buf ? buf[k] : nullptr;
      tab.ema[tab.n] = ema ? ema[k] : nullptr;
      tab.numel[tab.n] = numels[k];
      tab.chunk_start[tab.n] = (int)chunks;
      chunks += nb;
      ++tab.n;
      ++k;
    }
    if (tab.n == 0) {
      DYCON_REQUIRE(k >= n_tensors, DYCON_ERR_UNSUPPORTED, "sgd/ema step: tensor %d too large for one launch", k);
      break;
    }
    tab.chunk_start[tab.n] = (int)chunks;
    if (int rc = launch(tab, (int)chunks)) return rc;
  }
  return DYCON_OK;
}

long long resident_grid(long long chunks) {
  long long grid = (long long)sm_count() * 6;
  return grid > chunks ? chunks : grid;
}

}  // namespace
}  // namespace dycon

using namespace dycon;

extern "C" {

size_t dycon_grad_norm_workspace_bytes(void) { return 16 + sizeof(double) * kMaxPartials; }

int dycon_grad_norm(const float* const* grad_ptrs, const int64_t* numels, int n_tensors, float max_norm, float* out2,
                    void* workspace, size_t workspace_bytes, dycon_stream_t stream) {
  DYCON_REQUIRE(n_tensors >= 0 && out2 && workspace, DYCON_ERR_ARG, "grad norm: bad arguments");
  DYCON_REQUIRE(n_tensors <= kMaxTensors, DYCON_ERR_UNSUPPORTED, "grad norm: %d tensors (at most %d per call)", n_tensors,
                kMaxTensors);
  DYCON_REQUIRE(workspace_bytes >= dycon_grad_norm_workspace_bytes() && aligned(workspace, 16), DYCON_ERR_WORKSPACE,
                "grad norm: workspace too small / misaligned");
  DYCON_REQUIRE(n_tensors == 0 || (grad_ptrs && numels), DYCON_ERR_ARG, "grad norm: NULL table");
  for (int k = 0; k < n_tensors; ++k)
    DYCON_REQUIRE(numels[k] >= 0 && aligned(grad_ptrs[k], 4), DYCON_ERR_ARG, "grad norm: tensor %d", k);
  ReduceWorkspace ws = carve_reduce_workspace(workspace);
  cudaStream_t st = as_stream(stream);
  bool launched = false;
  int rc = for_each_table(nullptr, grad_ptrs, nullptr, nullptr, numels, n_tensors, [&](const StepTable& tab, int chunks) {
    long long grid = resident_grid(chunks);
    if (grid > kMaxPartials) grid = kMaxPartials;
    grad_norm_kernel<<<(unsigned)grid, kThreads, 0, st>>>(tab, chunks, max_norm, ws.ticket, ws.partials, out2);
    DYCON_CUDA(cudaGetLastError());
    count_launches(1);
    launched = true;
    return (int)DYCON_OK;
  });
  if (rc) return rc;
  if (!launched) {      // no gradient at all: norm 0, coefficient 1
    const float init[2] = {0.f, 1.f};
    DYCON_CUDA(cudaMemcpyAsync(out2, init, sizeof(init), cudaMemcpyHostToDevice, st));
  }
  return DYCON_OK;
}

int dycon_sgd_ema_step(float* const* param_ptrs, const float* const* grad_ptrs, float* const* buf_ptrs,
                       float* const* ema_ptrs, const int64_t* numels, int n_tensors, float lr, float momentum,
                       float weight_decay, int nesterov, int first_step, float alpha, float one_minus_alpha,
                       const float* clip2, const int* skip_flag, int scale_grads, dycon_stream_t stream) {
  DYCON_REQUIRE(n_tensors >= 0, DYCON_ERR_ARG, "sgd/ema step: n_tensors=%d", n_tensors);
  if (n_tensors == 0) return DYCON_OK;
  DYCON_REQUIRE(param_ptrs && numels, DYCON_ERR_ARG, "sgd/ema step: NULL table");
  for (int k = 0; k < n_tensors; ++k) {
    DYCON_REQUIRE(numels[k] >= 0 && (numels[k] == 0 || param_ptrs[k]), DYCON_ERR_ARG, "sgd/ema step: tensor %d", k);
    const bool has_g = grad_ptrs && grad_ptrs[k];
    DYCON_REQUIRE(!has_g || momentum == 0.f || (buf_ptrs && buf_ptrs[k]), DYCON_ERR_ARG,
                  "sgd/ema step: tensor %d has a gradient but no momentum buffer", k);
    DYCON_REQUIRE(aligned(param_ptrs[k], 4) && (!grad_ptrs || aligned(grad_ptrs[k], 4)) && (!buf_ptrs || aligned(buf_ptrs[k], 4)) &&
                      (!ema_ptrs || aligned(ema_ptrs[k], 4)), DYCON_ERR_ARG, "sgd/ema step: tensor %d misaligned", k);
  }
  StepScalars sc{lr, momentum, weight_decay, alpha, one_minus_alpha, first_step, nesterov, scale_grads};
  cudaStream_t st = as_stream(stream);
  return for_each_table(param_ptrs, grad_ptrs, buf_ptrs, ema_ptrs, numels, n_tensors, [&](const StepTable& tab, int chunks) {
    sgd_ema_kernel<<<(unsigned)resident_grid(chunks), kThreads, 0, st>>>(tab, chunks, sc, clip2, skip_flag);
    DYCON_CUDA(cudaGetLastError());
    count_launches(1);
    return (int)DYCON_OK;
  });
}

int dycon_finite_check(const float* const* value_ptrs, int n, int* flag_out, unsigned long long* skipped_count,
                       dycon_stream_t stream) {
  DYCON_REQUIRE(value_ptrs && flag_out && n >= 1 && n <= 16, DYCON_ERR_ARG, "finite check: 1..16 device scalars, flag_out");
  FiniteTable tab;
  for (int k = 0; k < 16; ++k) tab.v[k] = k < n ? value_ptrs[k] : nullptr;
  for (int k = 0; k < n; ++k) DYCON_REQUIRE(tab.v[k] && aligned(tab.v[k], 4), DYCON_ERR_ARG, "finite check: value %d", k);
  finite_check_table_kernel<<<1, 1, 0, as_stream(stream)>>>(tab, n, flag_out, skipped_count);
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // extern "C"
