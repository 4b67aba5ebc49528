// C-ABI entry points of FeCL (include/dycon_b200.h): argument validation + dispatch on the
// similarity arithmetic (fp32 SIMT tiles / bf16 tcgen05 tiles).
#include "exchange.cuh"
#include "fecl_internal.h"

using namespace dycon;

namespace {

int check_shape(int B, int N, int D, int precision) {
  DYCON_REQUIRE(B > 0 && N > 0 && D > 0, DYCON_ERR_ARG, "FeCL: B=%d N=%d D=%d must be positive", B, N, D);
  DYCON_REQUIRE(precision == DYCON_FECL_FP32 || precision == DYCON_FECL_BF16 || precision == DYCON_FECL_FP16, DYCON_ERR_ARG,
                "FeCL: unknown precision %d", precision);
  DYCON_REQUIRE((long long)N * N < (1LL << 40), DYCON_ERR_UNSUPPORTED, "FeCL: N=%d too large", N);
  return DYCON_OK;
}


int fecl_fwd_checked(const float* feat, int64_t f_sb, int64_t f_sn, int64_t f_sd, const float* teacher, int64_t t_sb,
                   int64_t t_sn, int64_t t_sd, const float* labels, const float* row_weight, int B, int N, int D,
                   float inv_tau, float gamma, int use_focal, float cross_thresh, float lambda_cross, double inv_rows,
                   int precision, void* state, size_t state_bytes, double* sums_out, float* loss_out, void* workspace,
                   size_t workspace_bytes, const ExchangeCtx* xc, dycon_stream_t stream, const float* feat_scale = nullptr,
                   const float* teacher_scale = nullptr) {
  if (int rc = check_shape(B, N, D, precision)) return rc;
  DYCON_REQUIRE(feat && labels && state && sums_out && workspace, DYCON_ERR_ARG,
                "FeCL fwd: NULL feat/labels/state/sums_out/workspace");
  DYCON_REQUIRE(aligned(feat, 4) && aligned(teacher, 4) && aligned(labels, 4) && aligned(row_weight, 4) &&
                    aligned(state, 128) && aligned(sums_out, 8) && aligned(workspace, 16),
                DYCON_ERR_ARG, "FeCL fwd: misaligned pointer");
  DYCON_REQUIRE(f_sn > 0 && f_sd > 0 && (teacher == nullptr || (t_sn > 0 && t_sd > 0)), DYCON_ERR_ARG,
                "FeCL fwd: non-positive strides");
  const int has_teacher = teacher != nullptr;
  DYCON_REQUIRE(state_bytes >= dycon_fecl_state_bytes(B, N, D, has_teacher, precision), DYCON_ERR_WORKSPACE,
                "FeCL fwd: state %zu < %zu bytes", state_bytes, dycon_fecl_state_bytes(B, N, D, has_teacher, precision));
  DYCON_REQUIRE(workspace_bytes >= dycon_fecl_workspace_bytes(B, N, D, precision), DYCON_ERR_WORKSPACE,
                "FeCL fwd: workspace %zu < %zu bytes", workspace_bytes, dycon_fecl_workspace_bytes(B, N, D, precision));
  FeclProblem p{B, N, D, has_teacher,
                FeclScalars{inv_tau, gamma, cross_thresh, lambda_cross, (use_focal && row_weight == nullptr) ? 1 : 0},
                inv_rows, precision};
  FeclFwdArgs a{feat, f_sb, f_sn, f_sd, teacher, t_sb, t_sn, t_sd, labels, row_weight, state, sums_out, loss_out, workspace};
  a.xc = xc;
  DYCON_REQUIRE(aligned(feat_scale, 4) && aligned(teacher_scale, 4), DYCON_ERR_ARG, "FeCL fwd: misaligned row scale");
  a.feat_scale = feat_scale;
  a.teacher_scale = teacher ? teacher_scale : nullptr;
  return precision != DYCON_FECL_FP32 ? fecl_tc_fwd(p, a, as_stream(stream)) : fecl_simt_fwd(p, a, as_stream(stream));
}


}  // namespace

extern "C" {

size_t dycon_fecl_state_bytes(int B, int N, int D, int has_teacher, int precision) {
  if (B <= 0 || N <= 0 || D <= 0) return 0;
  return precision != DYCON_FECL_FP32 ? fecl_tc_state_bytes(B, N, D, has_teacher, precision == DYCON_FECL_FP16)
                                      : fecl_simt_state_bytes(B, N, D, has_teacher);
}

size_t dycon_fecl_workspace_bytes(int B, int N, int D, int precision) {
  if (B <= 0 || N <= 0 || D <= 0) return 0;
  return precision != DYCON_FECL_FP32 ? fecl_tc_workspace_bytes(B, N, D) : fecl_simt_workspace_bytes(B, N, D);
}

int dycon_fecl_fwd(const float* feat, int64_t f_sb, int64_t f_sn, int64_t f_sd, const float* teacher, int64_t t_sb,
                   int64_t t_sn, int64_t t_sd, const float* labels, const float* row_weight, int B, int N, int D,
                   float inv_tau, float gamma, int use_focal, float cross_thresh, float lambda_cross, double inv_rows,
                   int precision, void* state, size_t state_bytes, double* sums_out, float* loss_out, void* workspace,
                   size_t workspace_bytes, dycon_stream_t stream) {
  return fecl_fwd_checked(feat, f_sb, f_sn, f_sd, teacher, t_sb, t_sn, t_sd, labels, row_weight, B, N, D, inv_tau, gamma,
                          use_focal, cross_thresh, lambda_cross, inv_rows, precision, state, state_bytes, sums_out,
                          loss_out, workspace, workspace_bytes, nullptr, stream);
}

int dycon_fecl_fwd_sharded(const float* feat, int64_t f_sb, int64_t f_sn, int64_t f_sd, const float* teacher,
                           int64_t t_sb, int64_t t_sn, int64_t t_sd, const float* labels, const float* row_weight,
                           int B, int N, int D, float inv_tau, float gamma, int use_focal, float cross_thresh,
                           float lambda_cross, double inv_rows, int precision, void* state, size_t state_bytes,
                           double* sums_out, float* loss_out, void* workspace, size_t workspace_bytes,
                           void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                           double timeout_s, dycon_stream_t stream) {
  ExchangeCtx xc;
  if (int rc = make_exchange_ctx(&xc, peer_inboxes, rank, world, seq_counters, DYCON_CHANNEL_FECL, timeout_s)) return rc;
  if (xc.world == 1 || precision != DYCON_FECL_FP32)      // the tensor-core loss sweep exchanges in its own tail
    return fecl_fwd_checked(feat, f_sb, f_sn, f_sd, teacher, t_sb, t_sn, t_sd, labels, row_weight, B, N, D, inv_tau,
                            gamma, use_focal, cross_thresh, lambda_cross, inv_rows, precision, state, state_bytes,
                            sums_out, loss_out, workspace, workspace_bytes, xc.world > 1 ? &xc : nullptr, stream);
  if (int rc = fecl_fwd_checked(feat, f_sb, f_sn, f_sd, teacher, t_sb, t_sn, t_sd, labels, row_weight, B, N, D, inv_tau,
                                gamma, use_focal, cross_thresh, lambda_cross, inv_rows, precision, state, state_bytes,
                                sums_out, nullptr, workspace, workspace_bytes, nullptr, stream))
    return rc;
  return dycon_exchange_sums(sums_out, 3, sums_out, peer_inboxes, rank, world, seq_counters,
                             teacher ? DYCON_EXCHANGE_FECL_TEACHER : DYCON_EXCHANGE_FECL, inv_rows, lambda_cross, loss_out,
                             timeout_s, stream);
}

int dycon_fecl_bwd(const void* state, size_t state_bytes, const float* labels, int B, int N, int D, int has_teacher,
                   float inv_tau, float gamma, int use_focal, int has_row_weight, float cross_thresh,
                   float lambda_cross, int precision, const double* cross_cnt, const float* grad_out,
                   float* grad_feat, int64_t g_sb, int64_t g_sn, int64_t g_sd, dycon_stream_t stream) {
  if (int rc = check_shape(B, N, D, precision)) return rc;
  DYCON_REQUIRE(state && labels && grad_out && grad_feat, DYCON_ERR_ARG, "FeCL bwd: NULL state/labels/grad_out/grad_feat");
  DYCON_REQUIRE(!has_teacher || cross_cnt, DYCON_ERR_ARG, "FeCL bwd: the teacher term needs cross_cnt");
  DYCON_REQUIRE(aligned(state, 128) && aligned(grad_feat, 4) && aligned(cross_cnt, 8), DYCON_ERR_ARG,
                "FeCL bwd: misaligned pointer");
  DYCON_REQUIRE(g_sn > 0 && g_sd > 0 && g_sb >= 0, DYCON_ERR_ARG, "FeCL bwd: non-positive grad_feat strides");
  DYCON_REQUIRE(state_bytes >= dycon_fecl_state_bytes(B, N, D, has_teacher, precision), DYCON_ERR_WORKSPACE,
                "FeCL bwd: state %zu < %zu bytes", state_bytes, dycon_fecl_state_bytes(B, N, D, has_teacher, precision));
  FeclProblem p{B, N, D, has_teacher ? 1 : 0,
                FeclScalars{inv_tau, gamma, cross_thresh, lambda_cross, (use_focal && !has_row_weight) ? 1 : 0}, 0.0, precision};
  FeclBwdArgs a{state, labels, cross_cnt, grad_out, grad_feat, g_sb, g_sn, g_sd};
  return precision != DYCON_FECL_FP32 ? fecl_tc_bwd(p, a, as_stream(stream)) : fecl_simt_bwd(p, a, as_stream(stream));
}

int dycon_fecl_fwd_scaled(const float* feat, int64_t f_sb, int64_t f_sn, int64_t f_sd, const float* teacher, int64_t t_sb,
                          int64_t t_sn, int64_t t_sd, const float* feat_row_scale, const float* teacher_row_scale,
                          const float* labels, const float* row_weight, int B, int N, int D, float inv_tau, float gamma,
                          int use_focal, float cross_thresh, float lambda_cross, double inv_rows, int precision, void* state,
                          size_t state_bytes, double* sums_out, float* loss_out, void* workspace, size_t workspace_bytes,
                          void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                          double timeout_s, dycon_stream_t stream) {
  ExchangeCtx xc;
  if (int rc = make_exchange_ctx(&xc, peer_inboxes, rank, world, seq_counters, DYCON_CHANNEL_FECL, timeout_s)) return rc;
  const bool fused = xc.world > 1 && precision != DYCON_FECL_FP32;
  if (int rc = fecl_fwd_checked(feat, f_sb, f_sn, f_sd, teacher, t_sb, t_sn, t_sd, labels, row_weight, B, N, D, inv_tau,
                                gamma, use_focal, cross_thresh, lambda_cross, inv_rows, precision, state, state_bytes,
                                sums_out, (xc.world > 1 && !fused) ? nullptr : loss_out, workspace, workspace_bytes,
                                fused ? &xc : nullptr, stream, feat_row_scale, teacher_row_scale))
    return rc;
  if (xc.world > 1 && !fused)
    return dycon_exchange_sums(sums_out, 3, sums_out, peer_inboxes, rank, world, seq_counters,
                               teacher ? DYCON_EXCHANGE_FECL_TEACHER : DYCON_EXCHANGE_FECL, inv_rows, lambda_cross, loss_out,
                               timeout_s, stream);
  return DYCON_OK;
}

size_t dycon_debug_timeline(void* host_out, size_t bytes) { return fecl_tc_debug_timeline(host_out, bytes); }

// ---- global negatives: the merged batch as ONE sample, rows split over ranks (include/dycon_b200.h) -------
size_t dycon_fecl_gn_state_bytes(int B_all, int N, int D, int has_teacher, int precision) {
  if (B_all <= 0 || N <= 0 || D <= 0 || precision == DYCON_FECL_FP32) return 0;
  return fecl_tc_state_bytes(1, B_all * N, D, has_teacher, false);
}

int dycon_fecl_gn_layout(int B_all, int N, int D, int has_teacher, int precision, size_t* out6) {
  DYCON_REQUIRE(out6 && B_all > 0 && N > 0 && D > 0, DYCON_ERR_ARG, "FeCL gn layout: bad arguments");
  DYCON_REQUIRE(precision == DYCON_FECL_BF16 || precision == DYCON_FECL_FP16, DYCON_ERR_UNSUPPORTED,
                "FeCL global negatives run on the tensor-core path only (precision fp16 / bf16)");
  fecl_tc_layout(1, B_all * N, D, has_teacher, out6);
  return DYCON_OK;
}

int dycon_fecl_gn_fwd(int phase_mask, const float* feat_all, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                      const float* teacher_all, int64_t t_sb, int64_t t_sn, int64_t t_sd, const float* labels_all,
                      const float* row_weight_all, int B_all, int N, int D, float inv_tau, float gamma, int use_focal,
                      float cross_thresh, float lambda_cross, int precision, void* state, size_t state_bytes,
                      int row_lo, int row_hi, double* sums_out, void* workspace, size_t workspace_bytes,
                      dycon_stream_t stream) {
  if (int rc = check_shape(B_all, N, D, precision)) return rc;
  DYCON_REQUIRE(precision == DYCON_FECL_BF16 || precision == DYCON_FECL_FP16, DYCON_ERR_UNSUPPORTED,
                "FeCL global negatives run on the tensor-core path only (precision fp16 / bf16)");
  const long long M = (long long)B_all * N;
  DYCON_REQUIRE(M < (1LL << 30), DYCON_ERR_UNSUPPORTED, "FeCL gn: %lld merged rows", M);
  DYCON_REQUIRE(phase_mask > 0 && phase_mask < 16, DYCON_ERR_ARG, "FeCL gn: phase_mask=%d", phase_mask);
  DYCON_REQUIRE(0 <= row_lo && row_lo < row_hi && row_hi <= M, DYCON_ERR_ARG, "FeCL gn: rows [%d, %d) of %lld", row_lo,
                row_hi, M);
  DYCON_REQUIRE(feat_all && labels_all && state && workspace && ((phase_mask & 8) == 0 || sums_out), DYCON_ERR_ARG,
                "FeCL gn fwd: NULL feat/labels/state/workspace/sums_out");
  DYCON_REQUIRE(aligned(state, 128) && aligned(sums_out, 8) && aligned(workspace, 16), DYCON_ERR_ARG,
                "FeCL gn fwd: misaligned pointer");
  const int has_teacher = teacher_all != nullptr;
  DYCON_REQUIRE(state_bytes >= dycon_fecl_gn_state_bytes(B_all, N, D, has_teacher, precision), DYCON_ERR_WORKSPACE,
                "FeCL gn fwd: state %zu < %zu bytes", state_bytes,
                dycon_fecl_gn_state_bytes(B_all, N, D, has_teacher, precision));
  DYCON_REQUIRE(workspace_bytes >= fecl_tc_workspace_bytes(1, (int)M, D), DYCON_ERR_WORKSPACE,
                "FeCL gn fwd: workspace too small");
  FeclProblem p{1, (int)M, D, has_teacher,
                FeclScalars{inv_tau, gamma, cross_thresh, lambda_cross, (use_focal && row_weight_all == nullptr) ? 1 : 0},
                1.0 / (double)M, precision};
  FeclFwdArgs a{feat_all, f_sb, f_sn, f_sd, teacher_all, t_sb, t_sn, t_sd, labels_all, row_weight_all, state, sums_out,
                nullptr, workspace};
  a.merge_B = B_all;
  a.phase_mask = phase_mask;
  a.row_lo = row_lo;
  a.row_hi = row_hi;
  return fecl_tc_fwd(p, a, as_stream(stream));
}

int dycon_fecl_gn_bwd(const void* state, size_t state_bytes, const float* labels_all, int B_all, int N, int D,
                      int has_teacher, float inv_tau, float gamma, int use_focal, int has_row_weight, float cross_thresh,
                      float lambda_cross, int precision, int row_lo, int row_hi, const double* cross_cnt,
                      const float* grad_out, float* grad_feat, int64_t g_sb, int64_t g_sn, int64_t g_sd,
                      dycon_stream_t stream) {
  if (int rc = check_shape(B_all, N, D, precision)) return rc;
  DYCON_REQUIRE(precision == DYCON_FECL_BF16 || precision == DYCON_FECL_FP16, DYCON_ERR_UNSUPPORTED,
                "FeCL global negatives run on the tensor-core path only (precision fp16 / bf16)");
  const long long M = (long long)B_all * N;
  DYCON_REQUIRE(0 <= row_lo && row_lo < row_hi && row_hi <= M && (row_hi - row_lo) % N == 0, DYCON_ERR_ARG,
                "FeCL gn bwd: rows [%d, %d) must be whole samples of %d rows", row_lo, row_hi, N);
  DYCON_REQUIRE(state && labels_all && grad_out && grad_feat && (!has_teacher || cross_cnt), DYCON_ERR_ARG,
                "FeCL gn bwd: NULL argument");
  DYCON_REQUIRE(state_bytes >= dycon_fecl_gn_state_bytes(B_all, N, D, has_teacher, precision), DYCON_ERR_WORKSPACE,
                "FeCL gn bwd: state too small");
  FeclProblem p{1, (int)M, D, has_teacher ? 1 : 0,
                FeclScalars{inv_tau, gamma, cross_thresh, lambda_cross, (use_focal && !has_row_weight) ? 1 : 0}, 0.0, precision};
  FeclBwdArgs a{state, labels_all, cross_cnt, grad_out, grad_feat, g_sb, g_sn, g_sd};
  a.row_lo = row_lo;
  a.row_hi = row_hi;
  a.grad_rows = N;
  return fecl_tc_bwd(p, a, as_stream(stream));
}

}  // extern "C"
