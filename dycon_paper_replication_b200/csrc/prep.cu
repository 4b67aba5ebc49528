// K6 -- the caller-side tensor preparation at the FeCL boundary.
//
// Between the network and FeCLoss the reference step runs (code/train_DyCON_BraTS19.py:316-330, twins
// train_DyCON_Pancreas.py:219-232, train_DyCON_ISLES22.py:251-288):
//   emb  = F.normalize(features.view(B, C, -1).transpose(1, 2), dim=-1)      for student and teacher
//   mask = (F.avg_pool3d(label.float(), k, k) > 0.5).float().reshape(B, -1).unsqueeze(1),  k = 4 * feature_scaler
//          (ISLES22: per-axis kernels = input extent // feature extent)
// and autograd replays the normalisation backwards.  The FeCL pack kernel converts and transposes the embeddings
// anyway, so the normalisation only needs the per-row 1/max(|x|, 1e-12) (dycon_row_inv_norm) handed to it as a row
// scale -- the normalised fp32 embeddings are never materialised -- plus one kernel for the pooled mask
// (dycon_pool_mask) and one for the Jacobian of the normalisation applied to FeCL's gradient
// (dycon_normalize_bwd):  dx = (g - f (f . g)) / max(|x|, eps),  f = x / max(|x|, eps).
#include "common.cuh"

namespace dycon {
namespace {

constexpr float kNormEps = 1e-12f;     // F.normalize's eps

// ---- mask = avg_pool3d(label, (kh, kw, kd), stride = kernel) > 0.5 --------------------------------------------
// One block per (b, output y, output x): threads run along the contiguous z axis of the kh x kw input rows, so every
// load is coalesced; the kd values of an output cell are then added through shared memory.
template <typename T>
__global__ void __launch_bounds__(128)
pool_mask_kernel(const T* __restrict__ label, int H, int W, int Dz, int kh, int kw, int kd, float* __restrict__ mask) {
  extern __shared__ float zsum[];                  // od * kd partial sums
  const int oh = H / kh, ow = W / kw, od = Dz / kd;
  const int bid = blockIdx.x, ox = bid % ow, oy = (bid / ow) % oh, b = bid / (ow * oh);
  const T* base = label + ((size_t)b * H + (size_t)oy * kh) * W * Dz + (size_t)ox * kw * Dz;
  for (int z = threadIdx.x; z < od * kd; z += blockDim.x) {
    float acc = 0.f;
    for (int dy = 0; dy < kh; ++dy)
      for (int dx = 0; dx < kw; ++dx) acc += (float)base[((size_t)dy * W + dx) * Dz + z];
    zsum[z] = acc;
  }
  __syncthreads();
  const float cells = (float)kh * (float)kw * (float)kd;      // avg_pool3d divides the window sum by its size
  for (int oz = threadIdx.x; oz < od; oz += blockDim.x) {
    float acc = 0.f;
    for (int dz = 0; dz < kd; ++dz) acc += zsum[oz * kd + dz];
    mask[((size_t)b * oh + oy) * ow * od + (size_t)ox * od + oz] = acc / cells > 0.5f ? 1.f : 0.f;
  }
}

// ---- inv[b][n] = 1 / max(|x[b][n][:]|_2, eps) -----------------------------------------------------------------
// Rows interleaved (the caller's (D*N, 1, N) view, sn <= sd): a block owns 32 rows, lanes run along n (coalesced),
// its 8 warps split the D columns and meet in shared memory.  Rows contiguous: a warp owns a row.
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const float* __restrict__ x, int64_t sb, int64_t sn, int64_t sd, int N, int D, float* __restrict__ inv) {
  __shared__ float part[8][33];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (int64_t)b * sb;
  if (sn <= sd) {
    const int n = blockIdx.x * 32 + lane;
    float acc = 0.f;
    if (n < N) {
#pragma unroll 8
      for (int d = warp; d < D; d += 8) {
        const float v = __ldg(xb + (int64_t)n * sn + (int64_t)d * sd);
        acc = fmaf(v, v, acc);
      }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && n < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][lane];
      inv[(size_t)b * N + n] = 1.f / fmaxf(sqrtf(t), kNormEps);
    }
  } else {
    const int n = blockIdx.x * 8 + warp;
    if (n >= N) return;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = __ldg(xb + (int64_t)n * sn + (int64_t)d * sd);
      acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) inv[(size_t)b * N + n] = 1.f / fmaxf(sqrtf(acc), kNormEps);
  }
}

// ---- dx = (g - f (f . g)) inv,  f = x inv   (rows whose norm is below eps: dx = g inv, F.normalize's clamp) ----
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const float* __restrict__ x, int64_t xb_, int64_t xn, int64_t xd, const float* __restrict__ g, int64_t gb,
                     int64_t gn, int64_t gd, const float* __restrict__ inv, int N, int D, float* __restrict__ dx, int64_t db,
                     int64_t dn, int64_t dd) {
  __shared__ float part[8][33];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xp = x + (int64_t)b * xb_;
  const float* gp = g + (int64_t)b * gb;
  float* dp = dx + (int64_t)b * db;
  if (xn <= xd) {          // 32 rows per block, lanes along n, the 8 warps split the columns
    const int n = blockIdx.x * 32 + lane;
    float dot = 0.f;
    if (n < N) {
#pragma unroll 8
      for (int d = warp; d < D; d += 8)
        dot = fmaf(__ldg(xp + (int64_t)n * xn + (int64_t)d * xd), __ldg(gp + (int64_t)n * gn + (int64_t)d * gd), dot);
    }
    part[warp][lane] = dot;
    __syncthreads();
    if (n >= N) return;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    const float r = __ldg(inv + (size_t)b * N + n);
    const float sc = r >= 1.f / kNormEps ? 0.f : t * r * r;          // (f . g) / |x| = (x . g) inv^2
#pragma unroll 8
    for (int d = warp; d < D; d += 8) {                               // (second read of x, g: L1 / L2 hits)
      const float xv = __ldg(xp + (int64_t)n * xn + (int64_t)d * xd), gv = __ldg(gp + (int64_t)n * gn + (int64_t)d * gd);
      dp[(int64_t)n * dn + (int64_t)d * dd] = fmaf(-xv, sc, gv) * r;  // (g - x (x.g) inv^2) inv
    }
  } else {                 // rows contiguous: a warp owns a row
    const int n = blockIdx.x * 8 + warp;
    if (n >= N) return;
    const float r = __ldg(inv + (size_t)b * N + n);
    float dot = 0.f;
    for (int d = lane; d < D; d += 32)
      dot = fmaf(__ldg(xp + (int64_t)n * xn + (int64_t)d * xd), __ldg(gp + (int64_t)n * gn + (int64_t)d * gd), dot);
    dot = warp_sum(dot);
    const float sc = r >= 1.f / kNormEps ? 0.f : dot * r * r;
    for (int d = lane; d < D; d += 32) {
      const float xv = __ldg(xp + (int64_t)n * xn + (int64_t)d * xd), gv = __ldg(gp + (int64_t)n * gn + (int64_t)d * gd);
      dp[(int64_t)n * dn + (int64_t)d * dd] = fmaf(-xv, sc, gv) * r;
    }
  }
}

}  // namespace
}  // namespace dycon

using namespace dycon;

extern "C" {

int dycon_pool_mask(const void* label, int label_dtype, int B, int H, int W, int Dz, int kh, int kw, int kd, float* mask_out,
                    dycon_stream_t stream) {
  DYCON_REQUIRE(label && mask_out, DYCON_ERR_ARG, "pool mask: NULL label / mask_out");
  DYCON_REQUIRE(B > 0 && kh > 0 && kw > 0 && kd > 0 && H >= kh && W >= kw && Dz >= kd, DYCON_ERR_ARG,
                "pool mask: B=%d volume %dx%dx%d kernel %dx%dx%d", B, H, W, Dz, kh, kw, kd);
  DYCON_REQUIRE(Dz <= 8192, DYCON_ERR_UNSUPPORTED, "pool mask: innermost extent %d > 8192", Dz);
  const int oh = H / kh, ow = W / kw;
  const unsigned grid = (unsigned)((size_t)B * oh * ow);
  const size_t smem = sizeof(float) * (size_t)(Dz / kd) * kd;
  cudaStream_t st = as_stream(stream);
  switch (label_dtype) {
    case DYCON_LABEL_INT64:
      pool_mask_kernel<long long><<<grid, 128, smem, st>>>((const long long*)label, H, W, Dz, kh, kw, kd, mask_out);
      break;
    case DYCON_LABEL_FLOAT32:
      pool_mask_kernel<float><<<grid, 128, smem, st>>>((const float*)label, H, W, Dz, kh, kw, kd, mask_out);
      break;
    case DYCON_LABEL_UINT8:
      pool_mask_kernel<unsigned char><<<grid, 128, smem, st>>>((const unsigned char*)label, H, W, Dz, kh, kw, kd, mask_out);
      break;
    default: return fail(DYCON_ERR_ARG, "pool mask: unknown label dtype %d", label_dtype);
  }
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

int dycon_row_inv_norm(const float* x, int64_t sb, int64_t sn, int64_t sd, int B, int N, int D, float* inv_out,
                       dycon_stream_t stream) {
  DYCON_REQUIRE(x && inv_out && B > 0 && N > 0 && D > 0 && sn > 0 && sd > 0, DYCON_ERR_ARG, "row inv norm: bad arguments");
  const dim3 grid(sn <= sd ? (N + 31) / 32 : (N + 7) / 8, B);
  row_inv_norm_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, sb, sn, sd, N, D, inv_out);
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

int dycon_normalize_bwd(const float* x, int64_t x_sb, int64_t x_sn, int64_t x_sd, const float* g, int64_t g_sb, int64_t g_sn,
                        int64_t g_sd, const float* inv_norm, int B, int N, int D, float* dx, int64_t d_sb, int64_t d_sn,
                        int64_t d_sd, dycon_stream_t stream) {
  DYCON_REQUIRE(x && g && inv_norm && dx && B > 0 && N > 0 && D > 0, DYCON_ERR_ARG, "normalize bwd: bad arguments");
  DYCON_REQUIRE(x_sn > 0 && x_sd > 0 && g_sn > 0 && g_sd > 0 && d_sn > 0 && d_sd > 0, DYCON_ERR_ARG, "normalize bwd: non-positive strides");
  const dim3 grid(x_sn <= x_sd ? (N + 31) / 32 : (N + 7) / 8, B);
  normalize_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_sb, x_sn, x_sd, g, g_sb, g_sn, g_sd, inv_norm, N, D, dx, d_sb,
                                                            d_sn, d_sd);
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // extern "C"
