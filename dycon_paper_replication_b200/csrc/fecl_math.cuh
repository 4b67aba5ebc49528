// Per-pair FeCL arithmetic shared by the fp32 (SIMT) and bf16 (tcgen05) kernels.
// Reference: code/utils/dycon_losses.py:172-231; closed forms in SURVEY.md section 0.2.
//
// Notation (one sample, row i, column j, i != j):
//   l_ij = (f_i . f_j)/tau          m_j = max(0, max_{i != j} l_ij)   (column max == row max: S is symmetric)
//   e_ij = exp(l_ij - m_j)          n_i = sum_k neg_ik e_ik
//   T_ij = e_ij + n_i               d_ij = e_ij / T_ij
//   phi(d) = -log(d) * (1-d)^gamma  (focal)   or   -log(d)
//   row loss_i = c_i sum_j pos_ij phi(d_ij),   c_i = 1/(P_i - 1 + 1e-18)
//   A_i = sum_j pos_ij phi'(d_ij) d_ij / T_ij                     (backward row scalar)
//   dL/dl_ij = kappa_i [ pos_ij phi'(d_ij) d_ij (1-d_ij) - neg_ij e_ij A_i ],  kappa_i = r_i c_i /(B N)
#pragma once

#include <cuda_runtime.h>

namespace dycon {

constexpr float kTiny = 1e-18f;  // the reference's epsilon (dycon_losses.py:187,189,192,229)

struct FeclScalars {
  float inv_tau;
  float gamma;
  float cross_thresh;
  float lambda_cross;
  int focal;  // use_focal && no row_weight  (dycon_losses.py:196,209-211)
};

template <bool kFast>
__device__ __forceinline__ float f_exp(float x) { return kFast ? __expf(x) : expf(x); }
template <bool kFast>
__device__ __forceinline__ float f_log(float x) { return kFast ? __logf(x) : logf(x); }
template <bool kFast>
__device__ __forceinline__ float f_div(float a, float b) { return kFast ? __fdividef(a, b) : a / b; }
template <bool kFast>
__device__ __forceinline__ float f_pow(float x, float g) {
  if (g == 2.f) return x * x;
  if (g == 1.f) return x;
  if (g == 0.f) return 1.f;      // torch.pow(0., 0.) = 1; __powf(0, 0) would be NaN
  if (g == 3.f) return x * x * x;
  return kFast ? __powf(x, g) : powf(x, g);
}

// Positive pair, forward: phi(d) and the A_i summand phi'(d) d / T.  One reciprocal, one log.
template <bool kFast>
__device__ __forceinline__ void fecl_pos_fwd(float e, float n, const FeclScalars& sc, float& phi, float& a_term) {
  const float rT = f_div<kFast>(1.f, e + n + kTiny);
  const float d = e * rT;
  const float logd = f_log<kFast>(d + kTiny);
  if (sc.focal) {
    const float omd = fmaxf(1.f - d, 0.f);               // a row without negatives has d = 1 (+ ulp): (1-1)^gamma = 0
    const float w1 = f_pow<kFast>(omd, sc.gamma - 1.f);  // (1-d)^(gamma-1)
    const float w = w1 * omd;
    phi = -logd * w;
    a_term = (sc.gamma * w1 * d * logd - w) * rT;        // phi'(d) = gamma w1 log d - w/d
  } else {
    phi = -logd;
    a_term = -rT;                                        // phi'(d) = -1/d
  }
}

// Positive pair, backward: phi'(d) d (1-d)  (multiply by kappa_i outside).
template <bool kFast>
__device__ __forceinline__ float fecl_pos_bwd(float e, float n, const FeclScalars& sc) {
  const float d = e * f_div<kFast>(1.f, e + n + kTiny);
  const float omd = fmaxf(1.f - d, 0.f);
  if (sc.focal) {
    const float logd = f_log<kFast>(d + kTiny);
    const float w1 = f_pow<kFast>(omd, sc.gamma - 1.f);
    // phi'(d) d (1-d) = gamma (1-d)^gamma d log d - (1-d)^(gamma+1)
    return w1 * omd * (sc.gamma * d * logd - omd);
  }
  return -omd;  // (-1/d) d (1-d)
}

// Cross (teacher) pair: hard negative iff labels differ and cs > thresh (dycon_losses.py:223).
template <bool kFast>
__device__ __forceinline__ float fecl_cross_term(float cs) { return -f_log<kFast>(1.f - cs + kTiny); }

}  // namespace dycon
