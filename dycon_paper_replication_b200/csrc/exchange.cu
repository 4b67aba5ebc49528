// Partial-sum exchange over NVLink peer memory (the one exchange step of the sharded
// path is a 1- or 3-double all-reduce per loss).  A NCCL all-reduce of 32 bytes costs ~25 us per call once the
// stream hand-offs around it are counted; this kernel does the same exchange in a few microseconds on the
// caller's own stream: every rank stores its partial sums straight into each peer's inbox (P2P stores over
// NVLink / NVSwitch), publishes a sequence flag with system-scope release, spins until the flags of all peers
// have arrived in its own inbox and adds the partials in RANK order -- so every rank computes the bit-identical
// total.  No host involvement: the sequence number lives in device memory and is advanced by the kernel,
// which keeps the exchange CUDA-graph replayable.
//
// The protocol (inbox layout, slots, channels, time-out) is in exchange.cuh; the UnCL forward and the FeCL P2
// sweep run it in their own tail (dycon_uncl_fwd_sharded / dycon_fecl_fwd_sharded), this file is the
// stand-alone launch for everything else.
#include "exchange.cuh"

#include <cstdlib>

namespace dycon {
namespace {

struct ExchangeParams {
  ExchangeCtx x;
  const double* local;              // n partial sums of this rank
  double* out;                      // n totals
  int n;
  int kind;                         // DYCON_EXCHANGE_*: which loss to evaluate from the totals (0: none)
  double scale, lambda;             // 1/(B_global V) or 1/(B_global N); lambda_cross
  float* loss_out;
};

__global__ void __launch_bounds__(32)
exchange_sums_kernel(const __grid_constant__ ExchangeParams p) {
  const int lane = threadIdx.x;
  const double acc = exchange_warp(p.x, lane < p.n ? p.local[lane] : 0.0, p.n);
  if (lane < p.n) p.out[lane] = acc;
  // the loss from the reduced sums, in the same launch (dycon_losses.py:116-118 / :193,229-234)
  const double t1 = __shfl_sync(0xffffffffu, acc, 1), t2 = __shfl_sync(0xffffffffu, acc, 2);
  if (lane == 0 && p.loss_out) {
    if (p.kind == DYCON_EXCHANGE_UNCL) *p.loss_out = (float)(acc * p.scale);
    if (p.kind == DYCON_EXCHANGE_FECL) *p.loss_out = (float)(acc * p.scale);
    if (p.kind == DYCON_EXCHANGE_FECL_TEACHER) *p.loss_out = (float)(acc * p.scale + p.lambda * (t1 / (t2 + 1e-18)));
  }
}

}  // namespace

int make_exchange_ctx(ExchangeCtx* ctx, void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                      int channel, double timeout_s) {
  for (int r = 0; r < kXMaxRanks; ++r) ctx->inbox[r] = nullptr;
  ctx->seq = seq_counters;
  ctx->rank = 0; ctx->world = 1; ctx->channel = channel;
  ctx->timeout_ns = 0;
  if (world <= 1 && peer_inboxes == nullptr) return DYCON_OK;
  DYCON_REQUIRE(peer_inboxes && seq_counters, DYCON_ERR_ARG, "exchange: NULL inbox table / sequence counters");
  DYCON_REQUIRE(world >= 1 && world <= kXMaxRanks && rank >= 0 && rank < world, DYCON_ERR_ARG,
                "exchange: rank %d / world %d (at most %d ranks)", rank, world, kXMaxRanks);
  DYCON_REQUIRE(channel >= 0 && channel < kXChannels, DYCON_ERR_ARG, "exchange: channel %d", channel);
  for (int r = 0; r < world; ++r) {
    DYCON_REQUIRE(peer_inboxes[r] && aligned(peer_inboxes[r], 16), DYCON_ERR_ARG, "exchange: inbox of rank %d is NULL / misaligned", r);
    ctx->inbox[r] = reinterpret_cast<double*>(peer_inboxes[r]);
  }
  ctx->rank = rank; ctx->world = world;
  // A rank that lags by more than the timeout (rank-0-only validation, a data-loader stall) makes its peers give
  // up: they get NaN sums and an error word instead of a destroyed context.  Negative: DYCON_EXCHANGE_TIMEOUT_S
  // from the environment, default 600 s; 0: wait for ever, like a blocking collective.
  if (timeout_s < 0) {
    const char* e = getenv("DYCON_EXCHANGE_TIMEOUT_S");
    timeout_s = e ? atof(e) : 600.0;
    if (timeout_s < 0) timeout_s = 0;
  }
  ctx->timeout_ns = (unsigned long long)(timeout_s * 1e9);
  return DYCON_OK;
}

}  // namespace dycon

using namespace dycon;

extern "C" {

size_t dycon_exchange_inbox_bytes(void) { return sizeof(double) * kXInboxDoubles; }

size_t dycon_exchange_error_offset(void) { return sizeof(double) * kXChannels * kXSlots * kXMaxRanks * kXEntry; }

int dycon_exchange_enable_peer(int peer_device) {
  int dev = 0, can = 0;
  DYCON_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev) return DYCON_OK;
  DYCON_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  DYCON_REQUIRE(can, DYCON_ERR_DEVICE, "exchange: device %d cannot access device %d (no P2P path)", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return DYCON_OK;
  }
  DYCON_CUDA(e);
  return DYCON_OK;
}

int dycon_exchange_sums(const double* local, int n, double* out, void* const* peer_inboxes, int rank, int world,
                        unsigned long long* seq_counters, int kind, double scale, double lambda_cross, float* loss_out,
                        double timeout_s, dycon_stream_t stream) {
  DYCON_REQUIRE(local && out && peer_inboxes && seq_counters, DYCON_ERR_ARG, "exchange: NULL argument");
  DYCON_REQUIRE(n >= 1 && n <= kXMaxPayload, DYCON_ERR_ARG, "exchange: n=%d outside [1, %d]", n, kXMaxPayload);
  DYCON_REQUIRE(kind >= 0 && kind <= DYCON_EXCHANGE_FECL_TEACHER && (kind != DYCON_EXCHANGE_FECL_TEACHER || n >= 3),
                DYCON_ERR_ARG, "exchange: kind=%d with n=%d", kind, n);
  ExchangeParams p;
  if (int rc = make_exchange_ctx(&p.x, peer_inboxes, rank, world, seq_counters, DYCON_CHANNEL_PLAIN, timeout_s)) return rc;
  p.local = local; p.out = out; p.n = n;
  p.kind = kind; p.scale = scale; p.lambda = lambda_cross; p.loss_out = loss_out;
  exchange_sums_kernel<<<1, 32, 0, as_stream(stream)>>>(p);
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // extern "C"
