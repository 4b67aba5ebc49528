/*
 * dycon_b200.h -- C ABI of the B200-native DyCON loss hot path.
 *
 * One shared library (dycon_paper_replication_b200/_dycon_b200.so, built by
 * dycon_paper_replication_b200/csrc/build.py for sm_100a) exports exactly these
 * entry points.  Every pointer is a plain device pointer unless stated
 * otherwise; the library allocates no device memory, keeps no device state
 * between calls (host side: per-device kernel attributes and a per-thread cache
 * of TMA descriptors, each a pure function of its key), never synchronises the
 * host with the device and only enqueues work on the stream it is given (all
 * launches are CUDA-graph capturable).
 *
 * Reference interfaces replaced (rogeliorjr/DyCON_Paper_Replication):
 *   dycon_uncl_*   UnCLoss.forward + its autograd backward    code/utils/dycon_losses.py:94-118
 *   dycon_fecl_*   FeCLoss.forward + its autograd backward    code/utils/dycon_losses.py:150-235
 *                  (also legacy losses.FeCLoss                code/utils/losses.py:221-250)
 *   dycon_ema_*    update_ema_variables                       code/train_DyCON_BraTS19.py:155-164
 *
 * Return value: 0 on success, a negative DYCON_ERR_* code for a rejected call
 * (nothing was enqueued), or a positive cudaError_t if a launch failed.
 * dycon_last_error() returns a thread-local, human readable description.
 * Floating-point inputs are IEEE fp32; NaN/Inf in the inputs propagate to the
 * outputs (the caller's isnan/isinf guard, train_DyCON_BraTS19.py:360, keeps
 * working).
 */
#ifndef DYCON_B200_H
#define DYCON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define DYCON_ABI_VERSION 2

#define DYCON_OK 0
#define DYCON_ERR_ARG (-1)         /* NULL / misaligned / out-of-range argument          */
#define DYCON_ERR_UNSUPPORTED (-2) /* shape or option outside what the kernels implement */
#define DYCON_ERR_DEVICE (-3)      /* current device is not sm_100 (B200)                */
#define DYCON_ERR_WORKSPACE (-4)   /* workspace / state buffer too small                 */

/* FeCL similarity arithmetic */
#define DYCON_FECL_FP32 0 /* SIMT fp32 tiles (exact mode, parity 1e-5)                              */
#define DYCON_FECL_BF16 1 /* TMA + tcgen05/TMEM tiles, bf16 operands, fp32 accumulate                  */
#define DYCON_FECL_FP16 2 /* same kernels with fp16 operands (10-bit mantissa): meets 2e-3 on all inputs */

typedef void* dycon_stream_t; /* a cudaStream_t (NULL = legacy default stream) */

int dycon_abi_version(void);
const char* dycon_last_error(void);
/* 0 if the *current* CUDA device can run this library (compute capability 10.x). */
int dycon_device_check(void);
/* Number of CUDA kernels this library has launched in this process (all threads, monotonic). */
uint64_t dycon_launch_count(void);

/* ------------------------------------------------------------------ UnCL
 * s_logits, t_logits: (B, C, V) fp32 contiguous (V = H*W*D voxels of one sample).
 * Per voxel: L_v = sum_c (ps-pt)^2 / (exp(beta*Hs) + exp(beta*Ht)) + beta*(Hs+Ht)
 * (dycon_losses.py:98-116).  inv_count is 1/(B_global*V): the mean denominator of
 * dycon_losses.py:116, passed explicitly so a batch shard produces correctly
 * scaled partial results.
 *
 * workspace: dycon_uncl_workspace_bytes() bytes, 16-byte aligned, ZERO-FILLED once
 * by the caller after allocation; each call leaves it zeroed again.  It must not be
 * shared by calls that may run concurrently (one workspace per stream).
 * stash (C == 2 only, else NULL): B*V floats; receives the unit gradient
 * dL_v/ds[:,1,v] so the backward is a 4-byte read instead of a recompute.
 * sum_out: 1 double, sum of L_v over the local voxels (the quantity to all-reduce).
 * loss_out: 1 float, sum * inv_count (may be NULL).
 */
size_t dycon_uncl_workspace_bytes(void);
int dycon_uncl_fwd(const float* s_logits, const float* t_logits, int64_t B, int C, int64_t V,
                   float beta, double inv_count, float* stash, double* sum_out, float* loss_out,
                   void* workspace, size_t workspace_bytes, dycon_stream_t stream);
/* grad_out: 1 float on the device (upstream gradient of the scalar loss).
 * grad_s: (B, C, V) fp32 contiguous, fully overwritten.
 * C == 2: reads only `stash` (s_logits/t_logits may be NULL).  C != 2: recomputes from
 * s_logits/t_logits (stash ignored). */
int dycon_uncl_bwd(const float* s_logits, const float* t_logits, const float* stash, int64_t B, int C,
                   int64_t V, float beta, double inv_count, const float* grad_out, float* grad_s,
                   dycon_stream_t stream);

/* ------------------------------------------------------------------ FeCL
 * feat / teacher: (B, N, D) fp32 with arbitrary ELEMENT strides (the caller's
 * normalize(transpose(view)) result has strides (D*N, 1, N), train_DyCON_BraTS19.py:316-323).
 * teacher may be NULL (no cross term).  labels: B*N fp32 contiguous (the (B,1,N) mask);
 * pairs are positive iff labels are equal (dycon_losses.py:172).  row_weight: B*N fp32 or
 * NULL -- the reference's gambling_uncertainty; when given, focal weighting is off
 * (dycon_losses.py:209-211).  cross_thresh = sigmoid_rampup(epoch, rampup, 0.3, 0.5)
 * computed by the host (dycon_losses.py:222).  inv_rows = 1/(B_global*N).
 *
 * state: dycon_fecl_state_bytes() bytes, 128-byte aligned; written by fwd, read by bwd
 * (operand copies in kernel layout + per-row statistics m, n, A, kappa).
 * workspace: dycon_fecl_workspace_bytes() bytes, ZERO-FILLED once by the caller, left
 * zeroed by each call, one per stream.
 * sums_out: 3 doubles {student_sum, cross_sum, cross_cnt} over the local samples
 *   (student_sum = sum_rows r_i*c_i*sum_j loss_ij; the loss is
 *    student_sum*inv_rows + lambda_cross*cross_sum/(cross_cnt + 1e-18)).
 * loss_out: 1 float, that expression evaluated on the local sums (may be NULL).
 */
size_t dycon_fecl_state_bytes(int B, int N, int D, int has_teacher, int precision);
size_t dycon_fecl_workspace_bytes(int B, int N, int D, int precision);
int dycon_fecl_fwd(const float* feat, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                   const float* teacher, int64_t t_sb, int64_t t_sn, int64_t t_sd,
                   const float* labels, const float* row_weight, int B, int N, int D,
                   float inv_tau, float gamma, int use_focal, float cross_thresh, float lambda_cross,
                   double inv_rows, int precision, void* state, size_t state_bytes,
                   double* sums_out, float* loss_out, void* workspace, size_t workspace_bytes,
                   dycon_stream_t stream);
/* cross_cnt: 1 double on the device -- the BATCH-GLOBAL hard-negative count
 * (dycon_losses.py:229; the all-reduced sums_out[2] when the batch is sharded).
 * grad_out: 1 float on the device.  grad_feat: (B, N, D) fp32 with ELEMENT strides g_sb/g_sn/g_sd
 * (dense: every element is written exactly once).  Passing feat's own strides -- (D*N, 1, N) for the
 * caller's normalize(transpose(view)) layout -- lets autograd hand the gradient on without a re-layout
 * copy; either g_sd == 1 (rows contiguous) or g_sn == 1 (columns contiguous) is the fast path. */
int dycon_fecl_bwd(const void* state, size_t state_bytes, const float* labels, int B, int N, int D,
                   int has_teacher, float inv_tau, float gamma, int use_focal, int has_row_weight,
                   float cross_thresh, float lambda_cross, int precision, const double* cross_cnt,
                   const float* grad_out, float* grad_feat, int64_t g_sb, int64_t g_sn, int64_t g_sd,
                   dycon_stream_t stream);

/* Measurement aid: a library built with -DDYCON_TIMELINE (DYCON_TIMELINE=1 python -m ...csrc.build) records
 * device-clock stamps of one CTA of the FeCL loss sweep and backward; this copies them to host_out ([2][4][64][8]
 * uint64) and returns the byte count -- 0 in a normal build.  tools/timeline.py prints them. */
size_t dycon_debug_timeline(void* host_out, size_t bytes);

/* ------------------------------------------------------------------ FeCL with global negatives
 * Extension for batches sharded over ranks (BASELINE config 5; not in the reference, whose contrast is per
 * sample): every row is contrasted against the rows of ALL B_all samples of the global batch.  The result is
 * the reference FeCLoss evaluated on feat.reshape(1, B_all*N, D), mask.reshape(1, 1, B_all*N) (that is the
 * oracle).  Every rank all-gathers the embeddings (feat_all / teacher_all: (B_all, N, D) fp32 with element
 * strides, labels_all: B_all*N fp32) and owns the rows [row_lo, row_hi) of the merged batch -- whole samples.
 * dycon_fecl_gn_fwd runs the phases of phase_mask for those rows: 1 pack (+ zero the statistics; the state
 * must have been ZERO-FILLED once by the caller), 2 row max m, 4 negative sums, 8 loss terms.  Between the
 * phases the caller all-gathers the row statistics, which live in `state` at the byte offsets returned by
 * dycon_fecl_gn_layout: out6 = {header, m plane, n plane, kappa plane, A plane 0, plane stride}; a plane holds
 * B_all*N floats.  Protocol: fwd(1|2) -> all-gather m -> fwd(4|8) -> all-reduce sums_out (3 doubles, as in
 * dycon_fecl_fwd; the loss is sums[0]/(B_all*N) + lambda*sums[1]/(sums[2]+1e-18)) -> [for the backward: add the
 * 8 A planes of the own rows into A plane 0, store 1.0f at header float 1, all-gather n, kappa, A plane 0]
 * -> dycon_fecl_gn_bwd, which writes the gradient of the OWN rows: grad_feat is the rank's local
 * ((row_hi-row_lo)/N, N, D) tensor with element strides.  Tensor-core precisions only.
 */
size_t dycon_fecl_gn_state_bytes(int B_all, int N, int D, int has_teacher, int precision);
int dycon_fecl_gn_layout(int B_all, int N, int D, int has_teacher, int precision, size_t* out6);
int dycon_fecl_gn_fwd(int phase_mask, const float* feat_all, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                      const float* teacher_all, int64_t t_sb, int64_t t_sn, int64_t t_sd, const float* labels_all,
                      const float* row_weight_all, int B_all, int N, int D, float inv_tau, float gamma, int use_focal,
                      float cross_thresh, float lambda_cross, int precision, void* state, size_t state_bytes,
                      int row_lo, int row_hi, double* sums_out, void* workspace, size_t workspace_bytes,
                      dycon_stream_t stream);
int dycon_fecl_gn_bwd(const void* state, size_t state_bytes, const float* labels_all, int B_all, int N, int D,
                      int has_teacher, float inv_tau, float gamma, int use_focal, int has_row_weight, float cross_thresh,
                      float lambda_cross, int precision, int row_lo, int row_hi, const double* cross_cnt,
                      const float* grad_out, float* grad_feat, int64_t g_sb, int64_t g_sn, int64_t g_sd,
                      dycon_stream_t stream);

/* ------------------------------------------------------------------ EMA
 * For every tensor k:  ema[k] = fma(one_minus_alpha, param[k], rn(ema[k]*alpha))  -- the
 * rounding order of ema.mul_(alpha).add_(param, alpha=1-alpha)
 * (train_DyCON_BraTS19.py:164).  alpha / one_minus_alpha are the fp32 roundings of the
 * host doubles alpha and 1-alpha with alpha = min(1 - 1/(step+1), ema_decay).
 * ema_ptrs / param_ptrs / numels are HOST arrays of n_tensors entries (copied into the
 * kernel's parameter space: no host->device copy, graph capturable).
 */
int dycon_ema_multi(float* const* ema_ptrs, const float* const* param_ptrs, const int64_t* numels,
                    int n_tensors, float alpha, float one_minus_alpha, dycon_stream_t stream);

/* ------------------------------------------------------------------ fused step losses (two classes)
 * The four voxel-wise losses the step loop computes from the same logits (code/train_DyCON_BraTS19.py:308-314,
 * 351-352), from ONE pass over s, t (B, 2, V) and the int64 labels of the first labeled_bs samples (labeled_bs, V):
 *   losses_out[0] = UnCLoss(s, t, beta)                                       code/utils/dycon_losses.py:94-118
 *   losses_out[1] = F.cross_entropy(s[:lb], label[:lb])                       (label == 1 -> class 1, else class 0)
 *   losses_out[2] = dice_loss(softmax(s)[:lb, 1], label[:lb] == 1)            code/utils/losses.py:8-16
 *   losses_out[3] = softmax_mse_loss(softmax(s)[lb:], softmax(t)[lb:]).mean() code/utils/losses.py:65-82 (the
 *                   reference hands it probabilities, so it is a softmax of a softmax: kept)
 * sums_out: 6 doubles {uncl, ce, dice intersect, dice z, dice y, consistency} kept for the backward.
 * dycon_segcons_bwd: grad_s (B, 2, V) = d(sum_k grad_out4[k] * losses_out[k]) / ds, grad_out4 = 4 floats on the
 * device.  C must be 2 (DYCON_ERR_UNSUPPORTED otherwise: compose the unfused entry points).
 */
size_t dycon_segcons_workspace_bytes(void);
int dycon_segcons_fwd(const float* s, const float* t, const long long* label, int B, int labeled_bs, int C, int64_t V,
                      float beta, double* sums_out, float* losses_out, void* workspace, size_t workspace_bytes,
                      dycon_stream_t stream);
int dycon_segcons_bwd(const float* s, const float* t, const long long* label, int B, int labeled_bs, int C, int64_t V,
                      float beta, const double* sums, const float* grad_out4, float* grad_s, dycon_stream_t stream);

/* ------------------------------------------------------------------ caller-side preparation at the FeCL boundary
 * What the step loop runs between the network and FeCLoss (code/train_DyCON_BraTS19.py:316-330):
 *   emb  = F.normalize(features.view(B, C, -1).transpose(1, 2), dim=-1)   -> dycon_row_inv_norm + a row scale folded
 *          into FeCL's operand staging (dycon_fecl_fwd_scaled): the normalised embeddings are never materialised
 *   mask = (F.avg_pool3d(label.float(), k, k) > 0.5).float()               -> dycon_pool_mask (per-axis kernels)
 * and, for the backward, the Jacobian of the normalisation applied to FeCL's gradient -> dycon_normalize_bwd:
 *   dx = (g - f (f . g)) * inv,  f = x * inv,  inv = 1/max(|x|_2, 1e-12)   (rows below eps: dx = g * inv).
 * label: (B, H, W, Dz) contiguous, dtype DYCON_LABEL_*; mask_out: (B, (H/kh)*(W/kw)*(Dz/kd)) floats in {0, 1}.
 * x / g / dx: (B, N, D) fp32 with ELEMENT strides; inv: B*N floats.
 * dycon_fecl_fwd_scaled: dycon_fecl_fwd with the two row-scale vectors (either may be NULL) and, optionally, the
 * sharded exchange of dycon_fecl_fwd_sharded (peer_inboxes NULL: not sharded).
 */
#define DYCON_LABEL_INT64 0
#define DYCON_LABEL_FLOAT32 1
#define DYCON_LABEL_UINT8 2
int dycon_pool_mask(const void* label, int label_dtype, int B, int H, int W, int Dz, int kh, int kw, int kd,
                    float* mask_out, dycon_stream_t stream);
int dycon_row_inv_norm(const float* x, int64_t sb, int64_t sn, int64_t sd, int B, int N, int D, float* inv_out,
                       dycon_stream_t stream);
int dycon_normalize_bwd(const float* x, int64_t x_sb, int64_t x_sn, int64_t x_sd, const float* g, int64_t g_sb, int64_t g_sn,
                        int64_t g_sd, const float* inv_norm, int B, int N, int D, float* dx, int64_t d_sb, int64_t d_sn,
                        int64_t d_sd, dycon_stream_t stream);
int dycon_fecl_fwd_scaled(const float* feat, int64_t f_sb, int64_t f_sn, int64_t f_sd, const float* teacher, int64_t t_sb,
                          int64_t t_sn, int64_t t_sd, const float* feat_row_scale, const float* teacher_row_scale,
                          const float* labels, const float* row_weight, int B, int N, int D, float inv_tau, float gamma,
                          int use_focal, float cross_thresh, float lambda_cross, double inv_rows, int precision, void* state,
                          size_t state_bytes, double* sums_out, float* loss_out, void* workspace, size_t workspace_bytes,
                          void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                          double timeout_s, dycon_stream_t stream);

/* ------------------------------------------------------------------ clip + SGD + EMA, finite check
 * Replaces, with two launches over ALL parameter tensors, the per-tensor launches of
 *   torch.nn.utils.clip_grad_norm_(params, max_norm); optimizer.step()  [torch.optim.SGD: lr, momentum,
 *   weight_decay, nesterov; dampening 0]; update_ema_variables(model, ema_model, ema_decay, iter_num)
 * (code/train_DyCON_BraTS19.py:268,366-372) and, through skip_flag, the host-side
 *   if torch.isnan(loss) or torch.isinf(loss): continue                                   (:360-362).
 * All pointer tables are HOST arrays of n_tensors entries (copied into the kernel parameters: graph capturable).
 * dycon_grad_norm: out2 = {||g||_2 over all tensors, min(1, max_norm / (norm + 1e-6))} on the device (max_norm <= 0:
 *   coefficient 1); NULL gradient entries are skipped; at most 160 tensors.
 * dycon_sgd_ema_step, per element, with the rounding order of the PyTorch ops:
 *   g' = rn(g * clip2[1]); g' = fma(wd, p, g'); buf = first_step ? g' : rn(rn(buf * momentum) + g');
 *   p = fma(-lr, nesterov ? fma(momentum, buf, g') : buf, p); ema = fma(one_minus_alpha, p, rn(ema * alpha)).
 *   grad_ptrs[k] == NULL: the parameter is not stepped (its teacher copy still follows it); ema_ptrs == NULL or
 *   ema_ptrs[k] == NULL: no teacher copy; clip2 == NULL: coefficient 1; scale_grads: also write g' = g * coefficient
 *   back (what clip_grad_norm_ leaves in .grad); skip_flag (1 int on the device, may be NULL) != 0: nothing is
 *   written at all.
 * dycon_finite_check: flag_out = 1 if any of the n <= 16 device scalars is NaN / +-inf, else 0; skipped_count
 *   (device uint64, may be NULL) is incremented when the flag is raised.
 */
size_t dycon_grad_norm_workspace_bytes(void);
int dycon_grad_norm(const float* const* grad_ptrs, const int64_t* numels, int n_tensors, float max_norm, float* out2,
                    void* workspace, size_t workspace_bytes, dycon_stream_t stream);
int dycon_sgd_ema_step(float* const* param_ptrs, const float* const* grad_ptrs, float* const* buf_ptrs,
                       float* const* ema_ptrs, const int64_t* numels, int n_tensors, float lr, float momentum,
                       float weight_decay, int nesterov, int first_step, float alpha, float one_minus_alpha,
                       const float* clip2, const int* skip_flag, int scale_grads, dycon_stream_t stream);
int dycon_finite_check(const float* const* value_ptrs, int n, int* flag_out, unsigned long long* skipped_count,
                       dycon_stream_t stream);

/* ------------------------------------------------------------------ sharded batches
 * The path shards over ranks along the batch; the only exchange is the all-reduce of the partial sums above
 * (dycon_uncl_fwd: sum_out; dycon_fecl_fwd: sums_out; the reference has no multi-process mode, its
 * DataParallel gather at train_DyCON_BraTS19.py:180-193 is what this replaces).  It runs over NVLink peer
 * memory: every rank owns an inbox of dycon_exchange_inbox_bytes() bytes (ZERO-FILLED once, 16-byte aligned,
 * mapped into every peer, e.g. through CUDA IPC); peer_inboxes is a HOST array of `world` device pointers (entry
 * r = rank r's inbox as mapped in THIS process, entry `rank` the local one).  seq_counters: DYCON_EXCHANGE_CHANNELS
 * device uint64, zero-initialised, private to this exchange object (one counter per channel: UnCL, FeCL and the
 * stand-alone call are independent exchanges); all ranks must issue the same sequence of calls per channel.
 * The totals are added in rank order, so every rank gets bit-identical results.  world <= 16.
 * timeout_s: how long a rank waits for its peers before it gives up -- the sums (and the loss) then become NaN
 * and the error word of the channel is raised in the local inbox (dycon_exchange_error_offset()), the launch
 * itself still completes.  0: wait for ever (a blocking collective); negative: the environment variable
 * DYCON_EXCHANGE_TIMEOUT_S, default 600.  Ranks may lag behind each other by up to that long (rank-0-only
 * validation, a data-loader stall) -- the waiting ranks simply spin inside the kernel.
 *
 * dycon_uncl_fwd_sharded / dycon_fecl_fwd_sharded: the forward of one rank's shard WITH the exchange in the tail
 * of the kernel that produces the sums (the last block of the UnCL forward / of the FeCL loss sweep pushes its
 * sums to the peers and waits for theirs): sum_out / sums_out / loss_out are the GLOBAL sums and loss, and the
 * step has no extra launch in front of the backward.  inv_count / inv_rows must be the global denominators.
 * dycon_exchange_sums(): the same exchange as a stand-alone launch (local / out: n <= 7 doubles on the device,
 * out may alias local), used by the paths that do not fuse it (generic-C UnCL, fp32 FeCL, global negatives).
 */
#define DYCON_EXCHANGE_CHANNELS 3
#define DYCON_CHANNEL_PLAIN 0
#define DYCON_CHANNEL_UNCL 1
#define DYCON_CHANNEL_FECL 2
size_t dycon_exchange_inbox_bytes(void);
/* Byte offset inside an inbox of the DYCON_EXCHANGE_CHANNELS uint64 error words (0: no time-out so far). */
size_t dycon_exchange_error_offset(void);
/* Lets kernels of the CURRENT device store into memory of `peer_device` (cudaDeviceEnablePeerAccess; a no-op
 * if already enabled).  Call once per peer before the first exchange. */
int dycon_exchange_enable_peer(int peer_device);
/* kind / scale / lambda_cross / loss_out: optionally evaluate the loss from the totals in the same launch
 * (loss_out: 1 float on the device, may be NULL): UNCL: out[0]*scale (scale = 1/(B_global V));
 * FECL: out[0]*scale (scale = 1/(B_global N)); FECL_TEACHER: out[0]*scale + lambda_cross*out[1]/(out[2]+1e-18). */
#define DYCON_EXCHANGE_NONE 0
#define DYCON_EXCHANGE_UNCL 1
#define DYCON_EXCHANGE_FECL 2
#define DYCON_EXCHANGE_FECL_TEACHER 3
int dycon_exchange_sums(const double* local, int n, double* out, void* const* peer_inboxes, int rank, int world,
                        unsigned long long* seq_counters, int kind, double scale, double lambda_cross, float* loss_out,
                        double timeout_s, dycon_stream_t stream);
int dycon_uncl_fwd_sharded(const float* s, const float* t, int64_t B, int C, int64_t V, float beta, double inv_count,
                           float* stash, double* sum_out, float* loss_out, void* workspace, size_t workspace_bytes,
                           void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                           double timeout_s, dycon_stream_t stream);
int dycon_fecl_fwd_sharded(const float* feat, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                           const float* teacher, int64_t t_sb, int64_t t_sn, int64_t t_sd,
                           const float* labels, const float* row_weight, int B, int N, int D,
                           float inv_tau, float gamma, int use_focal, float cross_thresh, float lambda_cross,
                           double inv_rows, int precision, void* state, size_t state_bytes,
                           double* sums_out, float* loss_out, void* workspace, size_t workspace_bytes,
                           void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                           double timeout_s, dycon_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DYCON_B200_H */
