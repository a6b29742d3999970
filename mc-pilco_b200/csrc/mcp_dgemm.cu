// Host launcher of the FP64 tensor-core NT GEMM (see mcp_dgemm.cuh).
#include "mcp_dgemm.cuh"

namespace mcp {

template <int BM, int BN, int WM, int WN>
static int launch(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C,
                  int ldc, int tri, int kflags, cudaStream_t st) {
  auto kern = dgemm_nt_kernel<BM, BN, WM, WN>;
  static bool configured[MCP_MAX_DEVICES] = {};
  constexpr size_t smem = gemm_smem_bytes<BM, BN>();
  MCP_CUDA(ensure_dynamic_smem(configured, kern, (int)smem));
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
  kern<<<grid, 32 * WM * WN, smem, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kflags);
  MCP_LAUNCH_CHECK();
  return MCP_OK;
}

int dgemm_nt(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C,
             int ldc, int tri, int kflags, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MCP_OK;
  MCP_CHECK_ARG((lda % 2 == 0) && (ldb % 2 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0),
                "dgemm_nt: operands must be 16-byte aligned with even leading dimensions (lda=%d ldb=%d)", lda, ldb);
  // 128 x 128 tiles once there are enough of them to occupy most of the 148 SMs; below that 64 x 64 tiles (three CTAs per SM), and
  // for the small products of the real MC-PILCO shapes (M = 200..400 particles, N = 60..400 points) 32 x 32 tiles, which spread the
  // work over ~100 SMs instead of a dozen: one CTA's time per k-step is fixed by its SM's FP64 rate
  if (N >= 128 && (size_t)cdiv(M, 128) * cdiv(N, 128) >= 120) return launch<128, 128, 4, 2>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kflags, st);
  if ((size_t)cdiv(M, 64) * cdiv(N, 64) < 96) return launch<32, 32, 2, 2>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kflags, st);
  return launch<64, 64, 2, 2>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kflags, st);
}

}  // namespace mcp
