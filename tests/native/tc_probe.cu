// Hardware probe for the tcgen05 / TMA encodings in dycon_paper_replication_b200/csrc/tc_common.cuh.
// Stand-alone (no torch):  nvcc -gencode arch=compute_100a,code=sm_100a -I../../dycon_paper_replication_b200/csrc
//                               -I../../include tc_probe.cu ../../dycon_paper_replication_b200/csrc/tc_host.cu -o tc_probe
// Test 1: S = X * Y^T   (both operands K-major, SWIZZLE_128B, filled by TMA)          -> UMMA M=128,N=128
// Test 2: O = P * Y     (A = P written by threads with sw128_offset(); B = the same K-major
//                        bytes of Y read as an MN-major operand)                       -> UMMA M=128,N=Dp
// Prints max |error| against a double-precision host product of the bf16-rounded inputs.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dycon_b200.h"
#include "tc_common.cuh"

using namespace dycon;
using namespace dycon::tc;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)

// ---- test 1 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_kmajor(const __grid_constant__ CUtensorMap mx,
                                                    const __grid_constant__ CUtensorMap my, int kc, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + kc * 16384;
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, kc * 32768);
    for (int c = 0; c < kc; ++c) {
      tma_load_2d(sa + c * 16384, &mx, c * 64, 0, &bar_full);
      tma_load_2d(sb + c * 16384, &my, c * 64, 0, &bar_full);
    }
    mbar_wait(&bar_full, 0);
    tcgen05_after_sync();
    const uint32_t idesc = umma_idesc_bf16(128, 128, false, false);
    for (int c = 0; c < kc; ++c)
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, umma_desc_kmajor(smem_u32(sa + c * 16384) + k * 32),
                  umma_desc_kmajor(smem_u32(sb + c * 16384) + k * 32), idesc, (c | k) != 0);
    umma_commit(&bar_done);
  }
  mbar_wait(&bar_done, 0);
  tcgen05_after_sync();
  for (int ch = 0; ch < 4; ++ch) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 128 + ch * 32 + i] = v[i];
  }
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- test 2 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_mnmajor(const __grid_constant__ CUtensorMap my, const float* P, int kc,
                                                     uint32_t lbo, uint32_t sbo, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sp = smem;           // P: [128 i][64 j] bf16, K-major SW128 (16 KB)
  uint8_t* sy = smem + 16384;   // Y: kc chunks of [64 j][64 d] (8 KB each)
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = kc * 64;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  // thread r writes row r of P
  for (int j = 0; j < 64; ++j)
    *reinterpret_cast<__nv_bfloat16*>(sp + sw128_offset(threadIdx.x, j)) = __float2bfloat16(P[threadIdx.x * 64 + j]);
  fence_async_smem();
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, kc * 8192);
    for (int c = 0; c < kc; ++c) tma_load_2d(sy + c * 8192, &my, c * 64, 0, &bar_full);
    mbar_wait(&bar_full, 0);
    tcgen05_after_sync();
    const uint32_t idesc = umma_idesc_bf16(128, n, false, true);
    for (int k = 0; k < 4; ++k)   // 16 j-rows of Y per step = 2048 B inside each chunk
      umma_bf16(tmem, umma_desc_kmajor(smem_u32(sp) + k * 32), umma_desc(smem_u32(sy) + k * 2048, lbo, sbo), idesc,
                k != 0);
    umma_commit(&bar_done);
  }
  mbar_wait(&bar_done, 0);
  tcgen05_after_sync();
  for (int ch = 0; ch < n / 32; ++ch) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * n + ch * 32 + i] = v[i];
  }
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  const int kc = argc > 1 ? atoi(argv[1]) : 4, dp = kc * 64;
  std::vector<float> X(128 * dp), Y(128 * dp), P(128 * 64);
  srand(1337);
  for (auto& v : X) v = bf((rand() / (float)RAND_MAX - 0.5f));
  for (auto& v : Y) v = bf((rand() / (float)RAND_MAX - 0.5f));
  for (auto& v : P) v = bf((rand() / (float)RAND_MAX - 0.5f));
  std::vector<__nv_bfloat16> Xb(X.size()), Yb(Y.size());
  for (size_t i = 0; i < X.size(); ++i) Xb[i] = __float2bfloat16(X[i]), Yb[i] = __float2bfloat16(Y[i]);
  __nv_bfloat16 *dX, *dY;
  float *dP, *dO;
  CK(cudaMalloc(&dX, Xb.size() * 2));
  CK(cudaMalloc(&dY, Yb.size() * 2));
  CK(cudaMalloc(&dP, P.size() * 4));
  CK(cudaMalloc(&dO, 128 * 256 * 4));
  CK(cudaMemcpy(dX, Xb.data(), Xb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dY, Yb.data(), Yb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
  int fails = 0;
  {
    CUtensorMap mx, my;
    if (make_tmap_bf16_2d(&mx, dX, 128, dp, 128) || make_tmap_bf16_2d(&my, dY, 128, dp, 128)) {
      printf("tensor map creation failed: %s\n", dycon_last_error());
      return 2;
    }
    const int smem = kc * 32768 + 1024;
    CK(cudaFuncSetAttribute(probe_kmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(dO, 0, 128 * 256 * 4));
    probe_kmajor<<<1, 128, smem>>>(mx, my, kc, dO);
    CK(cudaDeviceSynchronize());
    std::vector<float> S(128 * 128);
    CK(cudaMemcpy(S.data(), dO, S.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        double r = 0;
        for (int k = 0; k < dp; ++k) r += (double)X[i * dp + k] * Y[j * dp + k];
        worst = fmax(worst, fabs(r - S[i * 128 + j]));
      }
    printf("probe_kmajor  kc=%d  max|err|=%.3e  %s\n", kc, worst, worst < 1e-3 ? "OK" : "FAIL");
    fails += worst >= 1e-3;
  }
  {
    CUtensorMap my;
    if (make_tmap_bf16_2d(&my, dY, 128, dp, 64)) return 2;
    const int smem = 16384 + kc * 8192 + 1024;
    CK(cudaFuncSetAttribute(probe_mnmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // LBO = distance between 64-element MN blocks (one 8 KB chunk), SBO = 8 rows of K (1024 B)
    const uint32_t variants[1][2] = {{8192, 1024}};
    for (int vsel = 0; vsel < 1; ++vsel) {
      CK(cudaMemset(dO, 0, 128 * 256 * 4));
      probe_mnmajor<<<1, 128, smem>>>(my, dP, kc, variants[vsel][0], variants[vsel][1], dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("probe_mnmajor variant %d: %s\n", vsel, cudaGetErrorString(e));
        return 3;
      }
      std::vector<float> O(128 * dp);
      CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
      double worst = 0;
      for (int i = 0; i < 128; ++i)
        for (int d = 0; d < dp; ++d) {
          double r = 0;
          for (int j = 0; j < 64; ++j) r += (double)P[i * 64 + j] * Y[j * dp + d];
          worst = fmax(worst, fabs(r - O[i * dp + d]));
        }
      printf("probe_mnmajor kc=%d lbo=%u sbo=%u  max|err|=%.3e  %s\n", kc, variants[vsel][0], variants[vsel][1], worst,
             worst < 1e-3 ? "OK" : "FAIL");
      if (vsel == 0) fails += worst >= 1e-3;
    }
  }
  return fails ? 1 : 0;
}
