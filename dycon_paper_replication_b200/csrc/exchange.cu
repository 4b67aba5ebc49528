// Partial-sum exchange over NVLink peer memory (the one exchange step of the sharded
// path is a 1- or 3-double all-reduce per loss).  A NCCL all-reduce of 32 bytes costs ~25 us per call once the
// stream hand-offs around it are counted; this kernel does the same exchange in a few microseconds on the
// caller's own stream: every rank stores its partial sums straight into each peer's inbox (P2P stores over
// NVLink / NVSwitch), publishes a sequence flag with system-scope release, spins until the flags of all peers
// have arrived in its own inbox and adds the partials in RANK order -- so every rank computes the bit-identical
// total.  No host involvement: the sequence number lives in device memory and is advanced by the kernel,
// which keeps the exchange CUDA-graph replayable.
//
// Inbox layout (per rank, allocated by the host and shared through CUDA IPC):
//   [kSlots][kMaxRanks] entries of kEntry doubles: payload[0..6], then the sequence flag (as int64 bits).
// Two slots alternate with the sequence parity: a peer can only be one call ahead of this rank (it needs this
// rank's message of call k to finish call k), so the slot of call k is no longer read when call k+2 writes it.
#include "common.cuh"

namespace dycon {
namespace {

constexpr int kSlots = 2;
constexpr int kMaxRanks = 16;
constexpr int kEntry = 8;           // doubles per entry: 7 payload + 1 flag
constexpr int kMaxPayload = 7;

struct ExchangeParams {
  double* inbox[kMaxRanks];         // inbox[r] = base of rank r's inbox (peer-mapped device pointers)
  const double* local;              // n partial sums of this rank
  double* out;                      // n totals
  unsigned long long* seq;          // device counter, advanced once per call
  int n, rank, world;
  int kind;                         // DYCON_EXCHANGE_*: which loss to evaluate from the totals (0: none)
  double scale, lambda;             // 1/(B_global V) or 1/(B_global N); lambda_cross
  float* loss_out;
};

__global__ void __launch_bounds__(32)
exchange_sums_kernel(const __grid_constant__ ExchangeParams p) {
  const int lane = threadIdx.x;
  const unsigned long long seq = *p.seq + 1;                 // every lane reads the same value
  const int slot = (int)(seq & (kSlots - 1));
  if (lane < p.world) {
    // ---- send: my partials into slot[seq][my rank] of peer `lane` (my own inbox included) ----
    double* dst = p.inbox[lane] + ((size_t)slot * kMaxRanks + p.rank) * kEntry;
    for (int k = 0; k < p.n; ++k) dst[k] = p.local[k];
    __threadfence_system();                                  // payload before flag, at system scope
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(dst + kMaxPayload)),
                 "l"(seq)
                 : "memory");
    // ---- receive: wait for peer `lane`'s entry in MY inbox ----
    const double* src = p.inbox[p.rank] + ((size_t)slot * kMaxRanks + lane) * kEntry;
    unsigned long long got = 0;
    unsigned int spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got)
                   : "l"(reinterpret_cast<const unsigned long long*>(src + kMaxPayload))
                   : "memory");
      if (got != seq) {
        __nanosleep(100);
        if (++spins > (1u << 27)) __trap();                  // a peer never arrived: fail the launch, do not hang
      }
    } while (got != seq);
  }
  __syncwarp();
  if (lane < p.n) {                                          // fixed rank order: identical result on every rank
    double acc = 0.0;
    for (int r = 0; r < p.world; ++r) {
      const volatile double* src = p.inbox[p.rank] + ((size_t)slot * kMaxRanks + r) * kEntry;
      acc += src[lane];
    }
    p.out[lane] = acc;
    // the loss from the reduced sums, in the same launch (dycon_losses.py:116-118 / :193,229-234)
    const double t1 = __shfl_sync((1u << p.n) - 1u, acc, 1 < p.n ? 1 : 0);
    const double t2 = __shfl_sync((1u << p.n) - 1u, acc, 2 < p.n ? 2 : 0);
    if (lane == 0 && p.loss_out) {
      if (p.kind == DYCON_EXCHANGE_UNCL) *p.loss_out = (float)(acc * p.scale);
      if (p.kind == DYCON_EXCHANGE_FECL) *p.loss_out = (float)(acc * p.scale);
      if (p.kind == DYCON_EXCHANGE_FECL_TEACHER) *p.loss_out = (float)(acc * p.scale + p.lambda * (t1 / (t2 + 1e-18)));
    }
  }
  __syncwarp();
  if (lane == 0) *p.seq = seq;
}

}  // namespace
}  // namespace dycon

using namespace dycon;

extern "C" {

size_t dycon_exchange_inbox_bytes(void) { return sizeof(double) * kSlots * kMaxRanks * kEntry; }

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
                        unsigned long long* seq_counter, int kind, double scale, double lambda_cross, float* loss_out,
                        dycon_stream_t stream) {
  DYCON_REQUIRE(local && out && peer_inboxes && seq_counter, DYCON_ERR_ARG, "exchange: NULL argument");
  DYCON_REQUIRE(n >= 1 && n <= kMaxPayload, DYCON_ERR_ARG, "exchange: n=%d outside [1, %d]", n, kMaxPayload);
  DYCON_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, DYCON_ERR_ARG,
                "exchange: rank %d / world %d (at most %d ranks)", rank, world, kMaxRanks);
  ExchangeParams p;
  for (int r = 0; r < kMaxRanks; ++r) p.inbox[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    DYCON_REQUIRE(peer_inboxes[r] && aligned(peer_inboxes[r], 16), DYCON_ERR_ARG, "exchange: inbox of rank %d is NULL / misaligned", r);
    p.inbox[r] = reinterpret_cast<double*>(peer_inboxes[r]);
  }
  DYCON_REQUIRE(kind >= 0 && kind <= DYCON_EXCHANGE_FECL_TEACHER && (kind != DYCON_EXCHANGE_FECL_TEACHER || n >= 3),
                DYCON_ERR_ARG, "exchange: kind=%d with n=%d", kind, n);
  p.local = local; p.out = out; p.seq = seq_counter; p.n = n; p.rank = rank; p.world = world;
  p.kind = kind; p.scale = scale; p.lambda = lambda_cross; p.loss_out = loss_out;
  exchange_sums_kernel<<<1, 32, 0, as_stream(stream)>>>(p);
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // extern "C"
