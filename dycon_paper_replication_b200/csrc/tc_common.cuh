// Blackwell (sm_100a) primitives used by the FeCL tensor-core kernels: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and UMMA descriptors, as thin
// inline-PTX wrappers.  Encodings follow the PTX ISA "tcgen05" chapter; field layouts were
// cross-checked against the CuTe sm100 headers (mma_sm100_desc.hpp) shipped in this image.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dycon {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) __trap();
  }
}

// Same, for the single-lane producer / MMA-issuer warps: back off between polls so that their spin loop
// does not take issue slots from the epilogue warps of the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    __nanosleep(32);
    if (spin > (1u << 24)) __trap();
  }
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream still runs; pdl_wait() blocks until that predecessor has completed and its
// writes are visible, pdl_trigger() lets this kernel's own successor start early.  Always trigger AFTER
// waiting, so that "my successor runs" implies "everything before my predecessor is complete".
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- proxy / tcgen05 fences ----------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA -------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tile load: coordinates are (c0 = innermost element index, c1 = row); completes tx bytes on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2D tile store shared -> global (bulk async-group completion).  The writers of the tile fence.proxy.async and
// synchronise first; the issuing thread commits the group and, before the tile is overwritten, waits for the reads.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 1D bulk copy global -> shared (16-byte aligned addresses, size a multiple of 16); completes tx bytes on `bar`.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------------
// One full warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread `lane` of the warp receives row (lane base of taddr + lane),
// columns [col, col+32).  The warp may only touch lanes 32*(warp_id % 4) .. +31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// Same shape, 16 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64-bit):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100) | [49,52) base offset = 0 (tiles are 1024-B aligned) | [61,64) layout type
// Layout type 2 = SWIZZLE_128B.
//
// K-major operand tile [rows][64 bf16] (128-B rows, 8-row / 1024-B swizzle atoms, what TMA
// SWIZZLE_128B writes): SBO = 1024 (next 8-row group), LBO unused (1).  A 16-element K step
// advances the start address by 32 B inside the atom.
//
// MN-major operand (the same bytes read "transposed": MN = the 64 contiguous elements of a row,
// K = the rows): SBO = 1024 (next 8 rows of K), LBO = byte distance between consecutive
// 64-element MN blocks.
constexpr uint64_t kDescSw128 = (uint64_t(1) << 46) | (uint64_t(2) << 61);
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return kDescSw128 | uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32);
}
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) { return umma_desc(smem_addr, 16, 1024); }
// Descriptor of the same layout `bytes` further into shared memory (a multiple of 16; the sum must stay below
// 256 KB so that only the 14-bit start-address field changes): one 32-bit add in the MMA issue loop.
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) {
  return (desc & 0xffffffff00000000ull) | (uint32_t)((uint32_t)desc + (bytes >> 4));
}

// Instruction descriptor, kind::f16: D = f32 (bits [4,6) = 1), A / B format at bits [7,10) / [10,13)
// (0 = fp16, 1 = bf16), bit 15 / 16 = A / B major (0 = K, 1 = MN), [17,23) = N >> 3, [24,29) = M >> 4.
__host__ __device__ constexpr uint32_t umma_idesc_16(int m, int n, bool a_mn_major, bool b_mn_major, bool bf16) {
  return (1u << 4) | (uint32_t(bf16) << 7) | (uint32_t(bf16) << 10) | (uint32_t(a_mn_major) << 15) |
         (uint32_t(b_mn_major) << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
  return umma_idesc_16(m, n, a_mn_major, b_mn_major, true);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// Arrive on `bar` when every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Byte offset of element (row, col) inside a [rows][64 bf16] SWIZZLE_128B tile whose base is
// 1024-B aligned: the 16-byte unit index is XORed with (row % 8)  (Swizzle<3,4,3>).
__device__ __forceinline__ uint32_t sw128_offset(int row, int col) {
  return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)col & 7u) << 1);
}

}  // namespace tc

// ---- host side: tensor maps ------------------------------------------------------------------------
// 2D 16-bit tensor [rows][cols] (cols contiguous), box = [box_rows][64 cols], SWIZZLE_128B.
int make_tmap_16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool bf16);
// Store side of the stored-pairs matrices (tc_host.cu): 16-bit [rows][cols], box = [128 rows][16 cols] = a dense
// 4 KB tile of 32-byte rows in shared memory, no swizzle.
int make_tmap_16_store(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, bool bf16, uint32_t box_rows = 128);
inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_16_2d(out, base, rows, cols, box_rows, true);
}

}  // namespace dycon
