// Internal (non-ABI) declarations shared by the FeCL translation units.
#pragma once

#include "common.cuh"
#include "fecl_math.cuh"

namespace dycon {

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
};

struct FeclBwdArgs {
  const void* state;
  const float* labels;
  const double* cross_cnt;
  const float* grad_out;
  float* grad_feat;
  int64_t g_sb, g_sn, g_sd;   // element strides of grad_feat
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
size_t fecl_tc_state_bytes(int B, int N, int D, int has_teacher);
size_t fecl_tc_workspace_bytes(int B, int N, int D);
int fecl_tc_fwd(const FeclProblem& p, const FeclFwdArgs& a, cudaStream_t st);
int fecl_tc_bwd(const FeclProblem& p, const FeclBwdArgs& a, cudaStream_t st);

}  // namespace dycon
