// FP64 tensor-core GEMM for the posterior contraction  V[M,N] = A[M,K] * B[N,K]^T  (A = K*, B = K^-1, both row-major with
// the contraction index contiguous), Blackwell-style data movement:
//
//   * operand tiles ([128 rows x 16 doubles] = 128 rows x 128 B) are fetched by TMA (cp.async.bulk.tensor.2d) with the
//     hardware 128-byte swizzle; out-of-range rows / columns (M, N, K tails) are zero-filled by the TMA unit, so the
//     kernel has no predicates and no per-thread address arithmetic for loads;
//   * a 6-stage ring of {A tile, B tile} is handed from one producer lane (thread 0, between its own MMA work) to the 8
//     consumer warps through full/empty mbarriers — no __syncthreads in the main loop;
//   * consumers run DMMA.8x8x4 (FP64 has no tcgen05/TMEM kind; accumulators live in registers), each warp a 32 x 64 tile.
//
// Fragment <-> tile mapping.  A DMMA A-fragment gives lane (g = lane/4, q = lane%4) element A[g][q].  The tile row fed to
// fragment slot g is chosen as  r(i, g) = 16 (i / 2) + 2 g + (i % 2)  (i = fragment index): with the 128-byte swizzle
// (16-byte chunk index XOR row % 8) the four rows a half-warp touches then have row % 8 in {b, 2+b, 4+b, 6+b}, which sends
// its eight 32-byte pieces to eight distinct bank groups — conflict-free 64-bit fragment loads straight from the dense
// TMA layout.  The same permutation is applied to B rows (output columns); the epilogue undoes it: lane (g, q) of
// fragments (i, 2j') and (i, 2j'+1) owns four consecutive output columns 16 j' + 4 q .. + 3 of row r(i, g).
#include <cuda.h>

#include "mcp_dgemm.cuh"

namespace mcp {

constexpr int T_BM = 128, T_BN = 128, T_BK = 16, T_STAGES = 6, T_GROUP = 12;
constexpr int T_TILE_BYTES = T_BM * T_BK * 8;           // 16 KB per operand tile
constexpr int T_STAGE_BYTES = 2 * T_TILE_BYTES;         // A + B
constexpr int T_CONSUMER_WARPS = 8;
constexpr int T_THREADS = 32 * T_CONSUMER_WARPS;  // 256: lane 0 of warp 0 also drives the TMA unit (a 9th warp would cap
                                                  // ptxas at 168 registers/thread and spill the accumulators)
constexpr int T_LOOKAHEAD = T_STAGES - 2;         // tiles in flight ahead of the consumers; the slot being refilled was
                                                  // released two iterations ago, so the producer's wait is (almost) never taken
constexpr size_t T_SMEM_BYTES = (size_t)T_STAGES * T_STAGE_BYTES + 2 * T_STAGES * 8 + 1024;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// TRIM = false: the posterior contraction (full product, C = alpha A B^T).  TRIM = true: the precompute's triangular products
// (C = alpha A B^T + beta C with the lower-tile mask `tri` and the contraction-range flags `kflags` of mcp_dgemm.cuh).
template <bool TRIM>
__global__ void __launch_bounds__(T_THREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, double alpha,
                 double* __restrict__ C, int ldc, double beta, int tri, int kflags) {
  extern __shared__ unsigned char smem_raw[];
  const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-byte aligned tiles
  const unsigned char* tiles = smem_raw + (base - smem_u32(smem_raw));  // same address for ordinary (compiler-scheduled) loads
  const unsigned bars = base + T_STAGES * T_STAGE_BYTES;        // full[s] at bars + 8 s, empty[s] at bars + 8 (STAGES + s)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grouped rasterisation: consecutive CTAs (one wave = 148 of them) cover T_GROUP row blocks x ~148/T_GROUP column blocks, so a
  // wave pulls ~(12 + 12) operand panels through L2 instead of (2 + 64) with a row-major tile order
  const int tiles_m = (M + T_BM - 1) / T_BM, tiles_n = (N + T_BN - 1) / T_BN;
  const int per_group = T_GROUP * tiles_n, group = blockIdx.x / per_group, first_m = group * T_GROUP;
  const int rows_here = min(tiles_m - first_m, T_GROUP), in_group = blockIdx.x - group * per_group;
  int mt = first_m + in_group % rows_here;
  // contraction trimmed to k < m0 + tile (A lower triangular): the heavy row blocks are the last ones, schedule them first
  if (TRIM && (kflags & KF_A_LOWER) && !(kflags & (KF_A_UPPER | KF_B_UPPER | KF_B_LOWER))) mt = tiles_m - 1 - mt;
  const int m0 = mt * T_BM;
  int nt = in_group / rows_here;
  // contraction trimmed to k < n0 + tile (B lower triangular): the work of a tile grows with its column index, so the heavy
  // columns are scheduled first and the short ones fill the tail of the last wave
  if (TRIM && (kflags & KF_B_LOWER) && !(kflags & (KF_A_UPPER | KF_B_UPPER))) nt = tiles_n - 1 - nt;
  const int n0 = nt * T_BN;
  int kb = 0, ke = K;
  if (TRIM) {
    if (tri == 1 && n0 > m0 + T_BM - 1) return;
    if (kflags & KF_A_UPPER) kb = max(kb, m0);
    if (kflags & KF_B_UPPER) kb = max(kb, n0);
    if (kflags & KF_A_LOWER) ke = min(ke, m0 + T_BM);
    if (kflags & KF_B_LOWER) ke = min(ke, n0 + T_BN);
    kb = (kb / T_BK) * T_BK;
  }
  const int KT = ke > kb ? (ke - kb + T_BK - 1) / T_BK : 0;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < T_STAGES; s++) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (T_STAGES + s), T_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---------------- producer role (thread 0): issue one {A, B} stage ----------------
  auto produce = [&](int kn) {
    const int s = kn % T_STAGES;
    if (kn >= T_STAGES) mbar_wait(bars + 8 * (T_STAGES + s), ((kn / T_STAGES) - 1) & 1);
    const unsigned full = bars + 8 * s, dst = base + s * T_STAGE_BYTES;
    mbar_expect_tx(full, T_STAGE_BYTES);
    tma_load_2d(dst, &tmA, kb + kn * T_BK, m0, full);
    tma_load_2d(dst + T_TILE_BYTES, &tmB, kb + kn * T_BK, n0, full);
  };
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int kn = 0; kn < T_LOOKAHEAD && kn < KT; kn++) produce(kn);
  }

  // ---------------- consumers: 4 x 2 warps, each 32 x 64 of the 128 x 128 tile ----------------
  constexpr int MI = 4, NJ = 8;
  const int g = lane >> 2, q = lane & 3;
  const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 64;
  // byte offsets inside a stage; see the header comment for the row permutation and the swizzle algebra
  const unsigned kx = (unsigned)(g & 3) << 5;                                    // ((2 (g&3)) << 4): XORed with (ks*2) << 4
  const unsigned laneE = ((unsigned)(q >> 1) << 4) | ((unsigned)(q & 1) << 3);   // even fragments: row % 8 has bit 0 clear
  const unsigned laneO = ((unsigned)((q >> 1) ^ 1) << 4) | ((unsigned)(q & 1) << 3);
  const unsigned aE = (unsigned)(wm0 + 2 * g) * 128u + laneE, aO = (unsigned)(wm0 + 2 * g + 1) * 128u + laneO;
  const unsigned bE = T_TILE_BYTES + (unsigned)(wn0 + 2 * g) * 128u + laneE, bO = T_TILE_BYTES + (unsigned)(wn0 + 2 * g + 1) * 128u + laneO;

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < KT; kt++) {
    const int s = kt % T_STAGES;
    if (threadIdx.x == 0 && kt + T_LOOKAHEAD < KT) produce(kt + T_LOOKAHEAD);
    mbar_wait(bars + 8 * s, (kt / T_STAGES) & 1);
    const unsigned char* st = tiles + s * T_STAGE_BYTES;
#pragma unroll
    for (int ks = 0; ks < T_BK / 4; ks++) {
      const unsigned ko = ((unsigned)(ks * 2) << 4) ^ kx;
      double a[MI], b[NJ];
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = *reinterpret_cast<const double*>(st + (((i & 1) ? aO : aE) + ko + (unsigned)(i >> 1) * 2048u));
#pragma unroll
      for (int j = 0; j < NJ; j++) b[j] = *reinterpret_cast<const double*>(st + (((j & 1) ? bO : bE) + ko + (unsigned)(j >> 1) * 2048u));
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (T_STAGES + s));
  }

  // ---------------- epilogue: lane owns 4 consecutive columns per (i, j-pair) ----------------
#pragma unroll
  for (int i = 0; i < MI; i++) {
    const int r = m0 + wm0 + 16 * (i >> 1) + 2 * g + (i & 1);
    if (r >= M) continue;
#pragma unroll
    for (int jp = 0; jp < NJ / 2; jp++) {
      const int c = n0 + wn0 + 16 * jp + 4 * q;
      double* p = C + (size_t)r * ldc + c;
      double v0 = alpha * acc[i][2 * jp][0], v1 = alpha * acc[i][2 * jp + 1][0];
      double v2 = alpha * acc[i][2 * jp][1], v3 = alpha * acc[i][2 * jp + 1][1];
      if (TRIM && beta != 0.0) {
        if (c < N) v0 = fma(beta, p[0], v0);
        if (c + 1 < N) v1 = fma(beta, p[1], v1);
        if (c + 2 < N) v2 = fma(beta, p[2], v2);
        if (c + 3 < N) v3 = fma(beta, p[3], v3);
      }
      if (c + 3 < N) {
        *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
        *reinterpret_cast<double2*>(p + 2) = make_double2(v2, v3);
      } else {
        if (c < N) p[0] = v0;
        if (c + 1 < N) p[1] = v1;
        if (c + 2 < N) p[2] = v2;
        if (c + 3 < N) p[3] = v3;
      }
    }
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lock(init_mutex());
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// [rows x K] row-major fp64 matrix with leading dimension ld -> tiles of [128 rows x 16 doubles], 128-byte swizzle
static int make_map(CUtensorMap* map, const double* ptr, int rows, int K, int ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return MCP_E_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)T_BK, (cuuint32_t)T_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a [%d x %d] matrix with ld %d", (int)r, rows, K, ld);
    return MCP_E_CUDA;
  }
  return MCP_OK;
}

bool dgemm_tma_usable(const double* A, int lda, const double* B, int ldb, const double* C, int ldc) {
  return ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) && lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0 &&
         encode_fn() != nullptr;
}

template <bool TRIM>
static int launch_tma(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc,
                      int tri, int kflags, cudaStream_t st) {
  static bool configured[MCP_MAX_DEVICES] = {};
  MCP_CUDA(ensure_dynamic_smem(configured, dgemm_tma_kernel<TRIM>, (int)T_SMEM_BYTES));
  CUtensorMap tmA, tmB;
  if (int e = make_map(&tmA, A, M, K, lda)) return e;
  if (int e = make_map(&tmB, B, N, K, ldb)) return e;
  dim3 grid((unsigned)cdiv(N, T_BN) * (unsigned)cdiv(M, T_BM));
  dgemm_tma_kernel<TRIM><<<grid, T_THREADS, T_SMEM_BYTES, st>>>(tmA, tmB, M, N, K, alpha, C, ldc, beta, tri, kflags);
  MCP_LAUNCH_CHECK();
  return MCP_OK;
}

int dgemm_nt_tma(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MCP_OK;
  MCP_CHECK_ARG(dgemm_tma_usable(A, lda, B, ldb, C, ldc), "dgemm_nt_tma: operands must be 16-byte aligned with even leading dimensions");
  return launch_tma<false>(M, N, K, alpha, A, lda, B, ldb, 0.0, C, ldc, 0, 0, st);
}

int dgemm_nt_tma_trim(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc,
                      int tri, int kflags, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MCP_OK;
  MCP_CHECK_ARG(dgemm_tma_usable(A, lda, B, ldb, C, ldc), "dgemm_nt_tma_trim: operands must be 16-byte aligned with even leading dimensions");
  return launch_tma<true>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kflags, st);
}

}  // namespace mcp
