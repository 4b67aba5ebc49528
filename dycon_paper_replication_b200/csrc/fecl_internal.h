// Internal (non-ABI) declarations shared by the FeCL translation units.
#pragma once

#include "common.cuh"
#include "fecl_math.cuh"

namespace dycon {

struct ExchangeCtx;   // exchange.cuh

struct FeclProblem {
  int B, N, D;
  int has_teacher;
  FeclScalars sc;
  double inv_rows;
  int precision;
};

struct FeclFwdArgs {
  const float* feat;
  int64_t f_sb, f_sn, f_sd;
  const float* teacher;
  int64_t t_sb, t_sn, t_sd;
  const float* labels;
  const float* row_weight;
  void* state;
  double* sums_out;
  float* loss_out;
  void* workspace;
  // global-negatives mode (tensor-core path only): feat / teacher / labels hold merge_B samples of N / merge_B
  // rows that are contrasted as ONE sample of N rows; this call runs the phases in phase_mask (1 pack, 2 P0,
  // 4 P1, 8 P2) for the rows [row_lo, row_hi) -- the caller all-gathers the row statistics between phases
  int merge_B = 0;
  int phase_mask = 15;
  int row_lo = 0, row_hi = -1;
  const ExchangeCtx* xc = nullptr;   // sharded batch: exchange the three sums in the tail of the loss sweep
  // F.normalize folded into the operand staging: per-row factors (B*N floats each, or nullptr) applied to feat / teacher
  const float* feat_scale = nullptr;
  const float* teacher_scale = nullptr;
};

struct FeclBwdArgs {
  const void* state;
  const float* labels;
  const double* cross_cnt;
  const float* grad_out;
  float* grad_feat;
  int64_t g_sb, g_sn, g_sd;   // element strides of grad_feat
  int row_lo = 0, row_hi = -1;   // global-negatives mode: the rows this rank owns ...
  int grad_rows = 0;             // ... and the rows per sample of its (local) grad_feat
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Per-row statistics kept between forward and backward: planes of B*N floats (P = positive count,
// only used by the tensor-core path).
enum { kStatM = 0, kStatN = 1, kStatA = 2, kStatKappa = 3, kStatP = 4, kNumStats = 5 };

// fp32 SIMT path (fecl_simt.cu)
size_t fecl_simt_state_bytes(int B, int N, int D, int has_teacher);
size_t fecl_simt_workspace_bytes(int B, int N, int D);
int fecl_simt_fwd(const FeclProblem& p, const FeclFwdArgs& a, cudaStream_t st);
int fecl_simt_bwd(const FeclProblem& p, const FeclBwdArgs& a, cudaStream_t st);

// bf16 tcgen05 path (fecl_tc.cu)
// pairs: include the stored-pairs matrices of the per-sample loss (fecl_tc.cu: stored_pairs); never for a merged batch
size_t fecl_tc_state_bytes(int B, int N, int D, int has_teacher, bool pairs);
size_t fecl_tc_workspace_bytes(int B, int N, int D);
int fecl_tc_fwd(const FeclProblem& p, const FeclFwdArgs& a, cudaStream_t st);
int fecl_tc_bwd(const FeclProblem& p, const FeclBwdArgs& a, cudaStream_t st);
// byte offsets inside the tensor-core state: {hdr, m, n, kappa, A partial plane 0, plane stride}
void fecl_tc_layout(int B, int N, int D, int has_teacher, size_t out[6]);
size_t fecl_tc_debug_timeline(void* host_out, size_t bytes);

}  // namespace dycon
