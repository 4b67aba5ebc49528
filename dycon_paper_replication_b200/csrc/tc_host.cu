// Host-side TMA descriptor construction (driver entry point resolved through the runtime, so the
// library does not link against libcuda).
#include "common.cuh"
#include "tc_common.cuh"

#include <cudaTypedefs.h>

namespace dycon {

namespace {
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}
}  // namespace

// A tensor map is a pure function of (base, rows, cols, box rows, element type): a small per-thread cache of host-side
// descriptors saves the driver call on every launch of a training loop that reuses its buffers (the caching allocator
// hands the same state buffer back step after step) -- and keeps driver calls out of ncu's range replays.
struct TmapKey {
  const void* base;
  uint64_t rows, cols;
  uint32_t box_rows;
  bool bf16;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && box_rows == o.box_rows && bf16 == o.bf16;
  }
};
constexpr int kTmapCache = 64;
struct TmapCache {
  TmapKey key[kTmapCache];
  CUtensorMap map[kTmapCache];
  int n = 0, next = 0;
};

int make_tmap_16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool bf16) {
  static thread_local TmapCache cache;
  const TmapKey k{base, rows, cols, box_rows, bf16};
  for (int i = 0; i < cache.n; ++i) {
    if (cache.key[i] == k) {
      *out = cache.map[i];
      return DYCON_OK;
    }
  }
  auto fn = encode_fn();
  DYCON_REQUIRE(fn != nullptr, DYCON_ERR_DEVICE, "cuTensorMapEncodeTiled is not available from this driver");
  DYCON_REQUIRE(aligned(base, 128) && cols % 64 == 0 && box_rows >= 1 && box_rows <= 256, DYCON_ERR_ARG,
                "tensor map: base must be 128-B aligned, cols a multiple of 64, box_rows in [1, 256]");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};   // bytes, dimension 1
  const cuuint32_t box[2] = {64, box_rows};   // 64 bf16 = 128 B = one swizzle row
  const cuuint32_t elem_strides[2] = {1, 1};
  CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DYCON_REQUIRE(r == CUDA_SUCCESS, DYCON_ERR_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  const int slot = cache.n < kTmapCache ? cache.n++ : (cache.next = (cache.next + 1) % kTmapCache);
  cache.key[slot] = k;
  cache.map[slot] = *out;
  return DYCON_OK;
}

// Store side of the pair matrices of the FeCL sweeps: box = [box_rows][16 cols], no swizzle (the staging tile in shared
// memory is dense, 32-byte rows): 128 rows for the team-wide staging tiles of the stored-pairs loss sweep, 32 rows for the
// warp-private tiles of the similarity sweep.
int make_tmap_16_store(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, bool bf16, uint32_t box_rows) {
  static thread_local TmapCache cache;
  const TmapKey k{base, rows, cols, box_rows, bf16};
  for (int i = 0; i < cache.n; ++i) {
    if (cache.key[i] == k) {
      *out = cache.map[i];
      return DYCON_OK;
    }
  }
  auto fn = encode_fn();
  DYCON_REQUIRE(fn != nullptr, DYCON_ERR_DEVICE, "cuTensorMapEncodeTiled is not available from this driver");
  DYCON_REQUIRE(aligned(base, 128) && cols % 16 == 0 && rows > 0 && box_rows >= 1 && box_rows <= 256, DYCON_ERR_ARG,
                "store tensor map: bad base / cols / box");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {16, box_rows};
  const cuuint32_t elem_strides[2] = {1, 1};
  CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DYCON_REQUIRE(r == CUDA_SUCCESS, DYCON_ERR_ARG, "cuTensorMapEncodeTiled (store) failed with CUresult %d", (int)r);
  const int slot = cache.n < kTmapCache ? cache.n++ : (cache.next = (cache.next + 1) % kTmapCache);
  cache.key[slot] = k;
  cache.map[slot] = *out;
  return DYCON_OK;
}

}  // namespace dycon
