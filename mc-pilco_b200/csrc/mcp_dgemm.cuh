// FP64 tensor-core GEMM building block:  C[M,N] = alpha * A[M,K] * B[N,K]^T + beta * C   ("NT").
// Both operands are row-major with the contraction index contiguous, which is the natural layout of
// every product on this path: K* (particles x training points) times the symmetric K^-1, and all the
// blocked-Cholesky / triangular-inverse updates of the precompute.
//
// CTA tile BM x BN x 16, WARPS_M x WARPS_N warps, each warp an (8*MI) x (8*NJ) grid of DMMA.8x8x4 tiles,
// accumulators in registers (FP64 has no tcgen05/TMEM kind).  Operand tiles are staged with 16-byte
// cp.async into a 3-stage ring; rows are padded to 20 doubles so the 64-bit fragment loads of a half-warp
// hit 16 distinct 8-byte bank pairs.
#pragma once
#include "mcp_common.cuh"

namespace mcp {

constexpr int GEMM_BK = 16;
constexpr int GEMM_LDS = 20;  // padded row length (doubles) of a staged tile
constexpr int GEMM_STAGES = 3;

template <int BM, int BN>
constexpr size_t gemm_smem_bytes() {
  return (size_t)GEMM_STAGES * (BM + BN) * GEMM_LDS * sizeof(double);
}

// stage one [ROWS x 16] tile of a row-major matrix (rows r0.., columns k0..) into smem with zero fill
template <int ROWS, int NT>
__device__ __forceinline__ void gemm_load_tile(double* __restrict__ s, const double* __restrict__ G, int ld, int r0, int nrows,
                                               int k0, int K, int tid) {
#pragma unroll
  for (int c = tid; c < ROWS * 8; c += NT) {
    int row = c >> 3, kc = (c & 7) * 2;
    int gr = r0 + row, gk = k0 + kc;
    int rem = K - gk;
    int bytes = (gr < nrows && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
    const double* src = bytes ? (G + (size_t)gr * ld + gk) : G;
    cp_async16(s + row * GEMM_LDS + kc, src, bytes);
  }
}

// Accumulate acc += A[m0:m0+BM, kb:ke] * B[n0:n0+BN, kb:ke]^T.   acc[i][j][2] per the DMMA C layout.
template <int BM, int BN, int WARPS_M, int WARPS_N>
__device__ __forceinline__ void gemm_mainloop(const double* __restrict__ A, int lda, int M, int m0, const double* __restrict__ B,
                                              int ldb, int N, int n0, int kb, int ke, double* smem,
                                              double (&acc)[BM / WARPS_M / 8][BN / WARPS_N / 8][2]) {
  constexpr int NT = 32 * WARPS_M * WARPS_N;
  constexpr int MI = BM / WARPS_M / 8, NJ = BN / WARPS_N / 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % WARPS_M) * (BM / WARPS_M), wn0 = (warp / WARPS_M) * (BN / WARPS_N);
  const int g = lane >> 2, q = lane & 3;
  double* As = smem;
  double* Bs = smem + GEMM_STAGES * BM * GEMM_LDS;
  const int KT = (ke - kb + GEMM_BK - 1) / GEMM_BK;

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; s++) {
    if (s < KT) {
      gemm_load_tile<BM, NT>(As + s * BM * GEMM_LDS, A, lda, m0, M, kb + s * GEMM_BK, ke, tid);
      gemm_load_tile<BN, NT>(Bs + s * BN * GEMM_LDS, B, ldb, n0, N, kb + s * GEMM_BK, ke, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + GEMM_STAGES - 1;
      if (nk < KT) {
        int st = nk % GEMM_STAGES;
        gemm_load_tile<BM, NT>(As + st * BM * GEMM_LDS, A, lda, m0, M, kb + nk * GEMM_BK, ke, tid);
        gemm_load_tile<BN, NT>(Bs + st * BN * GEMM_LDS, B, ldb, n0, N, kb + nk * GEMM_BK, ke, tid);
      }
      cp_async_commit();
    }
    const double* as = As + (kt % GEMM_STAGES) * BM * GEMM_LDS + (wm0 + g) * GEMM_LDS + q;
    const double* bs = Bs + (kt % GEMM_STAGES) * BN * GEMM_LDS + (wn0 + g) * GEMM_LDS + q;
#pragma unroll
    for (int ks = 0; ks < GEMM_BK / 4; ks++) {
      double a[MI], b[NJ];
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = as[i * 8 * GEMM_LDS + ks * 4];
#pragma unroll
      for (int j = 0; j < NJ; j++) b[j] = bs[j * 8 * GEMM_LDS + ks * 4];
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// tri: 0 full; 1 only tiles touching the lower triangle (m >= n) are computed (SYRK-style).
// kflags trim the contraction range when an operand is known to be triangular in (row, k):
//   1: A[r][k] == 0 for k < r      2: B[r][k] == 0 for k < r      (upper-triangular rows)
//   4: A[r][k] == 0 for k > r      8: B[r][k] == 0 for k > r      (lower-triangular rows)
enum { KF_A_UPPER = 1, KF_B_UPPER = 2, KF_A_LOWER = 4, KF_B_LOWER = 8 };
template <int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(32 * WARPS_M * WARPS_N)
dgemm_nt_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                double beta, double* __restrict__ C, int ldc, int tri, int kflags) {
  extern __shared__ __align__(16) double smem[];
  constexpr int MI = BM / WARPS_M / 8, NJ = BN / WARPS_N / 8;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (tri == 1 && n0 > m0 + BM - 1) return;
  int kb = 0, ke = K;
  if (kflags & KF_A_UPPER) kb = max(kb, m0);
  if (kflags & KF_B_UPPER) kb = max(kb, n0);
  if (kflags & KF_A_LOWER) ke = min(ke, m0 + BM);
  if (kflags & KF_B_LOWER) ke = min(ke, n0 + BN);
  kb = (kb / GEMM_BK) * GEMM_BK;
  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  gemm_mainloop<BM, BN, WARPS_M, WARPS_N>(A, lda, M, m0, B, ldb, N, n0, kb, ke, smem, acc);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm0 = (warp % WARPS_M) * (BM / WARPS_M), wn0 = (warp / WARPS_M) * (BN / WARPS_N);
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int i = 0; i < MI; i++) {
    int r = m0 + wm0 + 8 * i + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      int c = n0 + wn0 + 8 * j + 2 * q;
      double* p = C + (size_t)r * ldc + c;
      if (c < N) p[0] = alpha * acc[i][j][0] + (beta != 0.0 ? beta * p[0] : 0.0);
      if (c + 1 < N) p[1] = alpha * acc[i][j][1] + (beta != 0.0 ? beta * p[1] : 0.0);
    }
  }
}

// host launcher (defined in mcp_dgemm.cu)
int dgemm_nt(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C,
             int ldc, int tri, int kflags, cudaStream_t st);

// TMA + mbarrier pipelined variant for full (non-triangular) products (mcp_dgemm_tma.cu)
bool dgemm_tma_usable(const double* A, int lda, const double* B, int ldb, const double* C, int ldc);
int dgemm_nt_tma(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc, cudaStream_t st);
// the same pipeline with beta, the lower-tile mask and the contraction-range flags (the precompute's large triangular products)
int dgemm_nt_tma_trim(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc,
                      int tri, int kflags, cudaStream_t st);

}  // namespace mcp
