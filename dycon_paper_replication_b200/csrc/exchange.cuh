// Device side of the partial-sum exchange over NVLink peer memory, shared by the
// stand-alone exchange kernel (exchange.cu) and by the kernels that run the exchange in their own tail:
// the last block of the UnCL forward and of the FeCL P2 sweep pushes its sums to the peers the moment they
// exist, so the sharded step has no extra launch in front of the backward.
//
// Inbox layout (per rank; allocated and ZERO-FILLED by the host, mapped into every peer through CUDA IPC):
//   [kChannels][kSlots][kMaxRanks] entries of 16 64-bit words, then one 64-bit error word per channel
//   (non-zero: a wait timed out).
// A double travels as TWO self-validating 64-bit words {32 bits of payload, 32-bit sequence tag} (the "LL"
// scheme): a 64-bit store is single-copy atomic, so the receiver simply polls every word until its tag equals
// the sequence number -- no payload / flag ordering, hence no system-scope fence and no second NVLink round trip
// on the sender's side (the fenced payload-then-flag version cost ~8 us per exchange; this one costs about the
// one-way latency).
// Channels are independent exchanges with their own sequence counters (UnCL, FeCL, stand-alone), because the
// kernels that use them may run in any order or concurrently.  Two slots alternate with the sequence parity: a
// peer can only be one call ahead of this rank (it needs this rank's message of call k to finish call k), so
// the slot of call k is no longer read when call k + 2 writes it.
#pragma once

#include "common.cuh"

namespace dycon {

constexpr int kXChannels = 3;        // DYCON_CHANNEL_*
constexpr int kXSlots = 2;
constexpr int kXMaxRanks = 16;
constexpr int kXEntry = 16;          // 64-bit words per entry: two per double
constexpr int kXMaxPayload = 7;
constexpr size_t kXInboxDoubles = (size_t)kXChannels * kXSlots * kXMaxRanks * kXEntry + kXChannels;

struct ExchangeCtx {
  double* inbox[kXMaxRanks];         // inbox[r] = base of rank r's inbox (peer-mapped device pointers)
  unsigned long long* seq;           // kXChannels device counters of THIS rank, advanced once per call
  unsigned long long timeout_ns;     // 0: wait for ever (like a blocking collective)
  int rank, world, channel;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All-reduce (sum, in RANK order: bit-identical on every rank) of n <= 7 doubles.  Called by ONE full warp; the
// totals are returned in lanes 0..n-1 (`mine` = this lane's local partial for lane < n).  A peer that does not
// show up within timeout_ns turns the totals into NaN and raises the channel's error word in the local inbox --
// the launch itself completes, so the CUDA context survives and the caller's NaN guard fires.
__device__ __forceinline__ double exchange_warp(const ExchangeCtx& x, double mine, int n) {
  const int lane = threadIdx.x & 31;
  unsigned long long* seqp = x.seq + x.channel;
  const unsigned long long seq = *seqp + 1;                     // every lane reads the same value
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const int slot = (int)(seq & (kXSlots - 1));
  const size_t base = ((size_t)x.channel * kXSlots + slot) * kXMaxRanks;
  bool late = false;
  double acc = 0.0;
  if (lane < n) {
    // ---- send: words 2 lane, 2 lane + 1 of entry [channel][slot][my rank] of every peer (my own inbox included) ----
    const unsigned long long bits = (unsigned long long)__double_as_longlong(mine);
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    for (int r = 0; r < x.world; ++r) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(x.inbox[r]) + (base + x.rank) * kXEntry + 2 * lane;
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w0) : "memory");
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(w1) : "memory");
    }
    // ---- receive, in rank order: identical result on every rank ----
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    for (int r = 0; r < x.world && !late; ++r) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(x.inbox[x.rank]) + (base + r) * kXEntry + 2 * lane;
      unsigned long long a, b;
      while (true) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(src + 1) : "memory");
        if ((a & 0xffffffff00000000ull) == tag && (b & 0xffffffff00000000ull) == tag) break;
        __nanosleep(32);
        if ((++spins & 1023u) == 0 && x.timeout_ns) {
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > x.timeout_ns) { late = true; break; }
        }
      }
      acc += __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
    }
  }
  late = __any_sync(0xffffffffu, late);
  if (late) acc = __longlong_as_double(0x7ff8000000000000ll);
  __syncwarp();
  if (lane == 0) {
    *seqp = seq;
    if (late) reinterpret_cast<unsigned long long*>(x.inbox[x.rank])[(size_t)kXChannels * kXSlots * kXMaxRanks * kXEntry + x.channel] = seq;
  }
  return acc;
}
#endif  // __CUDACC__

// Host side: validates the ABI arguments and fills the context; world <= 1 (or a NULL table) leaves
// ctx.world = 1, which the kernels treat as "no exchange".
int make_exchange_ctx(ExchangeCtx* ctx, void* const* peer_inboxes, int rank, int world, unsigned long long* seq_counters,
                      int channel, double timeout_s);

}  // namespace dycon
