// K2 (fp32 exact mode) -- FeCL forward / backward with SIMT fp32 similarity tiles.
// Reference: code/utils/dycon_losses.py:150-235.  Same phase structure as the tcgen05 path
// (fecl_tc.cu) and the same per-pair arithmetic (fecl_math.cuh), but S = F F^T and CS = F T^T are
// computed with fp32 FFMA tiles, so this mode meets the fp32 tolerance (1e-5) and serves as the
// on-device check of the bf16 tensor-core mode.  No (N,N) tensor is ever written to HBM.
//
// Phases (a CTA owns 64 rows of one sample and sweeps every 64-column tile):
//   P0  rowmax   : m_i = max(0, max_{j != i} l_ij), P_i = #same-label columns -> kappa_i
//   P1  fwd      : sweep 1 -> n_i ; sweep 2 -> row loss, A_i ; sweep 3 -> cross sum / count
//   P3  bwd      : per column tile: H = G + G^T (registers -> smem), dF += H F_J ; Gc, dF += Gc T_J
#include "fecl_internal.h"

namespace dycon {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
constexpr int kThreads = 256;

struct TileSmem {
  float a[2][BK][BM + PAD];
  float b[2][BK][BN + PAD];
};

// acc[r][c] = sum_k X[i0 + ty*4 + r][k] * Y[j0 + tx*4 + c][k];  X, Y row-major [N][D], D % 4 == 0.
__device__ __forceinline__ void tile_gemm(const float* __restrict__ X, const float* __restrict__ Y, int i0, int j0,
                                          int N, int D, float (&acc)[4][4], TileSmem& sm) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  auto gload = [&](const float* M, int r0, int k0) -> float4 {
    const int r = r0 + lrow, k = k0 + lk;
    if (r < N && k < D) return __ldg(reinterpret_cast<const float4*>(M + (size_t)r * D + k));
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](float (*dst)[BM + PAD], const float4& v) {
    dst[lk + 0][lrow] = v.x;
    dst[lk + 1][lrow] = v.y;
    dst[lk + 2][lrow] = v.z;
    dst[lk + 3][lrow] = v.w;
  };

  float4 ra = gload(X, i0, 0), rb = gload(Y, j0, 0);
  sstore(sm.a[0], ra);
  sstore(sm.b[0], rb);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < D; k0 += BK) {
    const bool more = k0 + BK < D;
    if (more) {
      ra = gload(X, i0, k0 + BK);
      rb = gload(Y, j0, k0 + BK);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sm.a[buf][kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sm.b[buf][kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    if (more) {
      sstore(sm.a[buf ^ 1], ra);
      sstore(sm.b[buf ^ 1], rb);
    }
    __syncthreads();
    buf ^= 1;
  }
}

// sum / max across the 16 lanes that share a row (tx = lane & 15)
__device__ __forceinline__ float row_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float row_max16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- pack: (B,N,D) with element strides -> contiguous [B][N][D] -------------------------------
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ src, int64_t sb, int64_t sn, int64_t sd, int N, int D,
                 float* __restrict__ dst, const float* __restrict__ row_scale /* B*N factors or nullptr */) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* s = src + (int64_t)b * sb;
  if (sn <= sd) {  // n is the faster dimension in memory (the caller's layout): lanes along n
    for (int k = ty; k < 32; k += 8) {
      const int n = n0 + tx, d = d0 + k;
      tile[k][tx] = (n < N && d < D) ? s[(int64_t)n * sn + (int64_t)d * sd] : 0.f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      const int n = n0 + k, d = d0 + tx;
      if (n < N && d < D) dst[((size_t)b * N + n) * D + d] = tile[tx][k] * (row_scale ? __ldg(row_scale + (size_t)b * N + n) : 1.f);
    }
  } else {  // d is the faster dimension: straight copy
    for (int k = ty; k < 32; k += 8) {
      const int n = n0 + k, d = d0 + tx;
      if (n < N && d < D)
        dst[((size_t)b * N + n) * D + d] = s[(int64_t)n * sn + (int64_t)d * sd] * (row_scale ? __ldg(row_scale + (size_t)b * N + n) : 1.f);
    }
  }
}

// ---- P0 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
fecl_simt_rowmax_kernel(const float* __restrict__ F, const float* __restrict__ labels,
                        const float* __restrict__ row_weight, int N, int D, float inv_tau, float inv_rows,
                        float* __restrict__ stat_m, float* __restrict__ stat_kappa) {
  __shared__ TileSmem sm;
  const int b = blockIdx.y, i0 = blockIdx.x * BM;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float* Fb = F + (size_t)b * N * D;
  const float* yb = labels + (size_t)b * N;
  float yi[4], mx[4], pc[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    yi[r] = i < N ? __ldg(yb + i) : 0.f;
    mx[r] = 0.f;  // the zeroed diagonal always takes part in the max (dycon_losses.py:178-180)
    pc[r] = 0.f;
  }
  for (int j0 = 0; j0 < N; j0 += BN) {
    float acc[4][4];
    tile_gemm(Fb, Fb, i0, j0, N, D, acc, sm);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= N) continue;
      const float yj = __ldg(yb + j);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        if (i != j) mx[r] = fmaxf(mx[r], acc[r][c] * inv_tau);
        pc[r] += (yi[r] == yj) ? 1.f : 0.f;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float m = row_max16(mx[r]);
    const float P = row_sum16(pc[r]);
    const int i = i0 + ty * 4 + r;
    if (tx == 0 && i < N) {
      const float rw = row_weight ? __ldg(row_weight + (size_t)b * N + i) : 1.f;
      stat_m[(size_t)b * N + i] = m;
      stat_kappa[(size_t)b * N + i] = rw / ((P - 1.f) + kTiny) * inv_rows;
    }
  }
}

// ---- P1 + P2 + cross ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
fecl_simt_fwd_kernel(const float* __restrict__ F, const float* __restrict__ T, const float* __restrict__ labels,
                     int N, int D, FeclScalars sc, double inv_rows, const float* __restrict__ stat_m,
                     const float* __restrict__ stat_kappa, float* __restrict__ stat_n, float* __restrict__ stat_a,
                     unsigned int* ticket, double* partials, double* __restrict__ sums_out,
                     float* __restrict__ loss_out) {
  __shared__ TileSmem sm;
  __shared__ double scratch[3 * 32];
  const int b = blockIdx.y, i0 = blockIdx.x * BM;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float* Fb = F + (size_t)b * N * D;
  const float* yb = labels + (size_t)b * N;
  const float* mb = stat_m + (size_t)b * N;
  float yi[4], nsum[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    yi[r] = i < N ? __ldg(yb + i) : 0.f;
    nsum[r] = 0.f;
  }
  // sweep 1: n_i = sum_k neg_ik exp(l_ik - m_k)      (dycon_losses.py:183-184)
  for (int j0 = 0; j0 < N; j0 += BN) {
    float acc[4][4];
    tile_gemm(Fb, Fb, i0, j0, N, D, acc, sm);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= N) continue;
      const float yj = __ldg(yb + j), mj = __ldg(mb + j);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (yi[r] != yj) nsum[r] += expf(acc[r][c] * sc.inv_tau - mj);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) nsum[r] = row_sum16(nsum[r]);

  // sweep 2: positives -> row loss and A_i             (dycon_losses.py:186-206)
  float lsum[4] = {0.f, 0.f, 0.f, 0.f}, asum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < N; j0 += BN) {
    float acc[4][4];
    tile_gemm(Fb, Fb, i0, j0, N, D, acc, sm);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= N) continue;
      const float yj = __ldg(yb + j), mj = __ldg(mb + j);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        if (yi[r] == yj && i != j) {
          float phi, at;
          fecl_pos_fwd<false>(expf(acc[r][c] * sc.inv_tau - mj), nsum[r], sc, phi, at);
          lsum[r] += phi;
          asum[r] += at;
        }
      }
    }
  }
  double red[3] = {0.0, 0.0, 0.0};  // student_sum, cross_sum, cross_cnt
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float ls = row_sum16(lsum[r]), as = row_sum16(asum[r]);
    const int i = i0 + ty * 4 + r;
    if (tx == 0 && i < N) {
      stat_n[(size_t)b * N + i] = nsum[r];
      stat_a[(size_t)b * N + i] = as;
      // kappa_i = r_i c_i inv_rows  ->  r_i c_i loss_i = kappa_i loss_i / inv_rows
      red[0] += (double)(__ldg(stat_kappa + (size_t)b * N + i) * ls);
    }
  }
  // sweep 3: cross term                                  (dycon_losses.py:213-229)
  if (T != nullptr) {
    const float* Tb = T + (size_t)b * N * D;
    float csum = 0.f, ccnt = 0.f;
    for (int j0 = 0; j0 < N; j0 += BN) {
      float acc[4][4];
      tile_gemm(Fb, Tb, i0, j0, N, D, acc, sm);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c;
        if (j >= N) continue;
        const float yj = __ldg(yb + j);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = i0 + ty * 4 + r;
          const float cs = acc[r][c];
          if (i < N && yi[r] != yj && cs > sc.cross_thresh) {
            csum += fecl_cross_term<false>(cs);
            ccnt += 1.f;
          }
        }
      }
    }
    red[1] = (double)csum;
    red[2] = (double)ccnt;
  }
  double total[3];
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (grid_sum_last_block<3>(red, total, ticket, partials, nblocks, bid, scratch) && threadIdx.x == 0) {
    const double student = total[0] / inv_rows;  // undo the inv_rows folded into kappa
    sums_out[0] = student;
    sums_out[1] = total[1];
    sums_out[2] = total[2];
    if (loss_out) {
      const double cross = (T != nullptr) ? total[1] / (total[2] + 1e-18) : 0.0;
      *loss_out = (float)(student * inv_rows + (double)sc.lambda_cross * cross);
    }
  }
}

// ---- P3: backward ------------------------------------------------------------------------------
template <int kGroups>  // D <= 64 * kGroups
__global__ void __launch_bounds__(kThreads)
fecl_simt_bwd_kernel(const float* __restrict__ F, const float* __restrict__ T, const float* __restrict__ labels,
                     int N, int D, FeclScalars sc, const float* __restrict__ stat_m,
                     const float* __restrict__ stat_n, const float* __restrict__ stat_a,
                     const float* __restrict__ stat_kappa, const double* __restrict__ cross_cnt,
                     const float* __restrict__ grad_out, float* __restrict__ grad_feat, int64_t g_sb, int64_t g_sn,
                     int64_t g_sd) {
  __shared__ TileSmem sm;
  __shared__ float hT[BN][BM + PAD];  // hT[j][i]
  const int b = blockIdx.y, i0 = blockIdx.x * BM;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const size_t off = (size_t)b * N;
  const float* Fb = F + off * D;
  const float* Tb = T ? T + off * D : nullptr;
  const float* yb = labels + off;
  float yi[4], mi[4], ni[4], ai[4], ki[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r, ic = i < N ? i : N - 1;
    yi[r] = __ldg(yb + ic);
    mi[r] = __ldg(stat_m + off + ic);
    ni[r] = __ldg(stat_n + off + ic);
    ai[r] = __ldg(stat_a + off + ic);
    ki[r] = i < N ? __ldg(stat_kappa + off + ic) : 0.f;
  }
  const float gc_scale = T ? sc.lambda_cross / ((float)(*cross_cnt) + kTiny) : 0.f;
  float dF[4][kGroups][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
#pragma unroll
      for (int c = 0; c < 4; ++c) dF[r][g][c] = 0.f;

  auto accumulate = [&](const float* __restrict__ Y, int j0) {
    // dF[i][d] += sum_j hT[j][i] * Y[j0 + j][d]
    __syncthreads();
    for (int j = 0; j < BN; ++j) {
      const int jj = j0 + j < N ? j0 + j : N - 1;  // rows past N carry hT == 0
      const float4 hv = *reinterpret_cast<const float4*>(&hT[j][ty * 4]);
      const float h[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        const int d = g * 64 + tx * 4;
        if (d < D) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(Y + (size_t)jj * D + d));
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            dF[r][g][0] = fmaf(h[r], yv.x, dF[r][g][0]);
            dF[r][g][1] = fmaf(h[r], yv.y, dF[r][g][1]);
            dF[r][g][2] = fmaf(h[r], yv.z, dF[r][g][2]);
            dF[r][g][3] = fmaf(h[r], yv.w, dF[r][g][3]);
          }
        }
      }
    }
    __syncthreads();
  };

  for (int j0 = 0; j0 < N; j0 += BN) {
    float acc[4][4];
    tile_gemm(Fb, Fb, i0, j0, N, D, acc, sm);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c, jc = j < N ? j : N - 1;
      const float yj = __ldg(yb + jc), mj = __ldg(stat_m + off + jc), nj = __ldg(stat_n + off + jc);
      const float aj = __ldg(stat_a + off + jc), kj = __ldg(stat_kappa + off + jc);
      float h[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        const float l = acc[r][c] * sc.inv_tau;
        const float eij = expf(l - mj), eji = expf(l - mi[r]);
        float g;
        if (yi[r] == yj) {
          g = ki[r] * fecl_pos_bwd<false>(eij, ni[r], sc) + kj * fecl_pos_bwd<false>(eji, nj, sc);
        } else {
          g = -(ki[r] * eij * ai[r] + kj * eji * aj);
        }
        h[r] = (i < N && j < N && i != j) ? g * sc.inv_tau : 0.f;
      }
      *reinterpret_cast<float4*>(&hT[tx * 4 + c][ty * 4]) = make_float4(h[0], h[1], h[2], h[3]);
    }
    accumulate(Fb, j0);
    if (Tb) {
      tile_gemm(Fb, Tb, i0, j0, N, D, acc, sm);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c, jc = j < N ? j : N - 1;
        const float yj = __ldg(yb + jc);
        float h[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = i0 + ty * 4 + r;
          const float cs = acc[r][c];
          const bool hard = i < N && j < N && yi[r] != yj && cs > sc.cross_thresh;
          h[r] = hard ? gc_scale / (1.f - cs + kTiny) : 0.f;
        }
        *reinterpret_cast<float4*>(&hT[tx * 4 + c][ty * 4]) = make_float4(h[0], h[1], h[2], h[3]);
      }
      accumulate(Tb, j0);
    }
  }
  const float go = __ldg(grad_out);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= N) continue;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
      const int d = g * 64 + tx * 4;
      if (d < D) {
        float* dst = grad_feat + (int64_t)b * g_sb + (int64_t)i * g_sn + (int64_t)d * g_sd;
        if (g_sd == 1 && ((g_sb | g_sn) & 3) == 0) {
          *reinterpret_cast<float4*>(dst) = make_float4(go * dF[r][g][0], go * dF[r][g][1], go * dF[r][g][2], go * dF[r][g][3]);
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[(int64_t)c * g_sd] = go * dF[r][g][c];
        }
      }
    }
  }
}

struct SimtState {
  float* F;
  float* T;
  float* stats;  // kNumStats planes of B*N
};
size_t operand_bytes(int B, int N, int D) { return align_up((size_t)B * N * D * sizeof(float), 128); }
size_t stats_bytes(int B, int N) { return align_up((size_t)kNumStats * B * N * sizeof(float), 128); }
SimtState carve(void* state, int B, int N, int D, int has_teacher) {
  char* p = reinterpret_cast<char*>(state);
  SimtState s;
  s.F = reinterpret_cast<float*>(p);
  p += operand_bytes(B, N, D);
  s.T = has_teacher ? reinterpret_cast<float*>(p) : nullptr;
  if (has_teacher) p += operand_bytes(B, N, D);
  s.stats = reinterpret_cast<float*>(p);
  return s;
}

}  // namespace

size_t fecl_simt_state_bytes(int B, int N, int D, int has_teacher) {
  return operand_bytes(B, N, D) * (has_teacher ? 2 : 1) + stats_bytes(B, N);
}
size_t fecl_simt_workspace_bytes(int, int, int) { return 16 + sizeof(double) * 3 * kMaxPartials; }

int fecl_simt_fwd(const FeclProblem& p, const FeclFwdArgs& a, cudaStream_t st) {
  const int B = p.B, N = p.N, D = p.D;
  DYCON_REQUIRE(D % 4 == 0 && D <= 256, DYCON_ERR_UNSUPPORTED, "FeCL fp32: D=%d must be a multiple of 4 and <= 256", D);
  const int row_blocks = (N + BM - 1) / BM;
  DYCON_REQUIRE((long long)row_blocks * B <= kMaxPartials && B <= 65535, DYCON_ERR_UNSUPPORTED,
                "FeCL fp32: %d row blocks x B=%d exceeds %d CTAs", row_blocks, B, kMaxPartials);
  SimtState s = carve(a.state, B, N, D, p.has_teacher);
  const size_t plane = (size_t)B * N;
  dim3 pgrid((N + 31) / 32, (D + 31) / 32, B);
  pack_rows_kernel<<<pgrid, 256, 0, st>>>(a.feat, a.f_sb, a.f_sn, a.f_sd, N, D, s.F, a.feat_scale);
  if (p.has_teacher) pack_rows_kernel<<<pgrid, 256, 0, st>>>(a.teacher, a.t_sb, a.t_sn, a.t_sd, N, D, s.T, a.teacher_scale);
  dim3 grid(row_blocks, B);
  fecl_simt_rowmax_kernel<<<grid, kThreads, 0, st>>>(s.F, a.labels, a.row_weight, N, D, p.sc.inv_tau,
                                                     (float)p.inv_rows, s.stats + kStatM * plane,
                                                     s.stats + kStatKappa * plane);
  ReduceWorkspace ws = carve_reduce_workspace(a.workspace);
  fecl_simt_fwd_kernel<<<grid, kThreads, 0, st>>>(s.F, s.T, a.labels, N, D, p.sc, p.inv_rows, s.stats + kStatM * plane,
                                                  s.stats + kStatKappa * plane, s.stats + kStatN * plane,
                                                  s.stats + kStatA * plane, ws.ticket, ws.partials, a.sums_out,
                                                  a.loss_out);
  DYCON_CUDA(cudaGetLastError());
  count_launches(p.has_teacher ? 4 : 3);
  return DYCON_OK;
}

int fecl_simt_bwd(const FeclProblem& p, const FeclBwdArgs& a, cudaStream_t st) {
  const int B = p.B, N = p.N, D = p.D;
  DYCON_REQUIRE(D % 4 == 0 && D <= 256, DYCON_ERR_UNSUPPORTED, "FeCL fp32: D=%d must be a multiple of 4 and <= 256", D);
  SimtState s = carve(const_cast<void*>(a.state), B, N, D, p.has_teacher);
  const size_t plane = (size_t)B * N;
  dim3 grid((N + BM - 1) / BM, B);
#define DYCON_LAUNCH_BWD(G)                                                                                      \
  fecl_simt_bwd_kernel<G><<<grid, kThreads, 0, st>>>(s.F, s.T, a.labels, N, D, p.sc, s.stats + kStatM * plane,   \
                                                     s.stats + kStatN * plane, s.stats + kStatA * plane,         \
                                                     s.stats + kStatKappa * plane, a.cross_cnt, a.grad_out,      \
                                                     a.grad_feat, a.g_sb, a.g_sn, a.g_sd)
  if (D <= 64) DYCON_LAUNCH_BWD(1);
  else if (D <= 128) DYCON_LAUNCH_BWD(2);
  else DYCON_LAUNCH_BWD(4);
#undef DYCON_LAUNCH_BWD
  DYCON_CUDA(cudaGetLastError());
  count_launches(1);
  return DYCON_OK;
}

}  // namespace dycon
