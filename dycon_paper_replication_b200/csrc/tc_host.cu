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

int make_tmap_16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool bf16) {
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
  return DYCON_OK;
}

}  // namespace dycon
