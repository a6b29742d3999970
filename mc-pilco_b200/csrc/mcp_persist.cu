// Whole-horizon PERSISTENT forward rollout for the reference's real cart-pole-sized configurations:
// M = 200..2000 particles, N <= ~300 training points, D <= 6 gp inputs, SE + Volterra kernels (BASELINE configs 1-3).
//
// At these sizes a rollout is ~1 GFLOP and the per-step kernel chain (small_step_kernel + small_gemm_kernel, mcp_small.cu) is pure
// latency: ~37 us per time step, although the arithmetic is worth ~5.  Particles are independent over the WHOLE horizon, so one
// kernel can run all H steps without ever synchronising the grid — if K^-1 (720 KB per output at N = 300) does not have to be
// streamed from L2 every step.  It stays resident by splitting it over a thread-block CLUSTER of 8 CTAs:
//
//   * a cluster owns a batch of up to 27 particles for the whole horizon; CTA r of the cluster owns the column slice
//     [r W, (r+1) W) of K^-1 (rows of the symmetric K^-1, so the slice is contiguous in memory) — 96 KB of shared memory at N = 300 —
//     the training inputs / alpha of those columns, and up to 4 of the batch's particles ("owner": integration, policy, checkpoints);
//   * per time step and output: every CTA evaluates the K* entries of ITS columns for all 27 particles and stores them into all
//     eight CTAs' K* buffers through distributed shared memory (all-gather by remote stores); after a cluster barrier each CTA
//     contracts the full K* rows with its K^-1 slice on FP64 DMMA (27 x W x N), folds its columns' share of the factored posterior
//     sums (mean, variance, the Jacobian channels of posterior_reduce_fast_kernel) — again as DMMA products: an 8 x 8 tile
//     [weight channel x feature] per particle — and sends each particle's 64 partial sums to the particle's owner (remote stores);
//   * after the next cluster barrier the owners add the eight partial tiles in rank order, finish mean / variance / Jacobians,
//     draw the reparameterised sample, integrate, apply the measurement model, evaluate the policy with dropout and broadcast the new
//     gp-input features to the cluster.  1 + 2 E cluster barriers per time step, no global-memory round trip on the step's
//     critical path, K^-1 never re-read from L2 when E = 1 (with several outputs the slice of the next output streams in behind
//     the current output's reduce).
//
// Same formulas as the per-step kernels (cov_fast_kernel, posterior_reduce_fast_kernel, integrate_warp_kernel,
// policy_fwd_block_kernel); sums are taken in a different order, so results agree to rounding, and — the slice geometry depending
// on N only and the batches being independent — a sharded rollout stays bit-identical to the unsharded one.
// Reference: MC_PILCO.apply_policy (policy_learning/MC_PILCO.py:615-674, :808-906), get_next_state (Model_learning.py:210-229),
// get_estimate_from_alpha (gpr_lib/GP_prior/GP_prior.py:137-155), Sum_of_gaussians.forward (Policy.py:242-265).
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "mcp_gpdev.cuh"
#include "mcp_kfn.cuh"
#include "mcp_rollout_dev.cuh"

namespace mcp {

constexpr int PK_CL = 8;                // CTAs per cluster
constexpr int PK_P = 27;                // particles per cluster batch: only 15 clusters of 8 CTAs with > 113 KB of shared memory each are
                                        // co-resident on a B200 (scripts/probe/cluster_occ.cu), and 15 x 27 covers the reference's 400
constexpr int PK_MT = (PK_P + 7) / 8;   // DMMA row tiles (rows past PK_P are clamped reads, their results dropped)
constexpr int PK_OWN = (PK_P + PK_CL - 1) / PK_CL;  // particles owned per CTA (64 threads each)
constexpr int PK_THREADS = 256;
constexpr int PK_WARPS = PK_THREADS / 32;
constexpr int PK_MAX_E = 4;
constexpr int PK_NV = 64;               // one 8 x 8 tile of partial sums per (particle, output)

struct PkGeom {
  int Kc;   // contraction length: N_max rounded up to 16
  int W;    // K^-1 columns per CTA: 8 W >= Kc
  int Wt;   // W rounded up to the DMMA tile (8)
  int LD;   // row stride of the K* rows and of the K^-1 slice rows: = 4 (mod 16) doubles -> conflict-free DMMA fragment loads
  int WS;   // row stride of the per-warp channel-weight rows
  size_t o_kinv, o_ks, o_v, o_vx, o_wt, o_yf, o_al, o_feat, o_part, o_sum, o_own, o_spec, o_bar, doubles;
};
__host__ __device__ inline PkGeom pk_geom(int Nmax, int E) {
  PkGeom g;
  g.Kc = (Nmax + 15) / 16 * 16;
  g.W = (g.Kc + PK_CL - 1) / PK_CL;
  g.Wt = (g.W + 7) / 8 * 8;
  g.LD = g.Kc + 4;
  g.WS = g.Wt + 4;
  size_t o = 0;
  g.o_kinv = o; o += (size_t)g.Wt * g.LD;
  g.o_ks = o;   o += (size_t)PK_P * g.LD;
  g.o_v = o;    o += (size_t)PK_P * g.Wt;
  g.o_vx = o;   o += (size_t)(PK_WARPS / 2) * 64;                    // second K-halves of the matvec's last, partial round of tiles
  g.o_wt = o;   o += (size_t)PK_WARPS * 8 * g.WS;
  g.o_yf = o;   o += (size_t)E * 8 * g.WS;
  g.o_al = o;   o += (size_t)E * g.Wt;
  g.o_feat = o; o += (size_t)PK_P * 8;
  g.o_part = o; o += (size_t)PK_CL * PK_OWN * PK_NV;         // partial tiles of the current output received from the eight CTAs
  g.o_sum = o;  o += (size_t)PK_OWN * E * PK_NV;             // their sums, per output
  g.o_own = o;  o += (size_t)PK_OWN * 136;                   // owner-side particle state
  g.o_spec = o; o += (size_t)E * 40;                         // kernel hyper-parameters of each output (PkSpec)
  g.o_bar = o;  o += 2;                                      // mbarrier of the K^-1 slice copies
  g.doubles = o;
  return g;
}

namespace {

__device__ __forceinline__ uint32_t pk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pk_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t pk_cluster_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a double at the same shared-memory location of CTA `rank` of this cluster
__device__ __forceinline__ void pk_st_remote(const double* local, uint32_t rank, double v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(pk_smem_u32(local)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
// K^-1 slice rows arrive by bulk copies that complete on one mbarrier
__device__ __forceinline__ void pk_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void pk_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pk_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void pk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pk_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// barrier among the 64 threads that work on one owned particle (ids 1..PK_OWN; 0 is __syncthreads)
__device__ __forceinline__ void pk_pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// owner-side state of one particle (doubles inside the o_own block)
constexpr int PK_MAX_DS = 8, PK_MAX_DU = 4, PK_MAX_DP = 16, PK_MAX_NPOS = 4;  // this path's limits (persist_path_ok)
struct PkOwn {
  double x[PK_MAX_DS], pin[PK_MAX_DS], u[PK_MAX_DU], nv[PK_MAX_NPOS], mean[PK_MAX_E], var[PK_MAX_E], jm[PK_MAX_E][8], jv[PK_MAX_E][8];
  double il[PK_MAX_DP], z[PK_MAX_DP], part[2][PK_MAX_DU];
};
static_assert(sizeof(PkOwn) == 136 * sizeof(double), "owner state block size");
// what the step's inner loops need of an output's McpGpSpec, staged once in shared memory (the descriptor table lives in global memory)
struct PkSpec {
  double ils[8], w1[8], w2a[8], w2b[8];  // inverse lengthscales; Volterra weights (degree 1; the two factors of degree 2)
  double lambda, o1, o2a, o2b, var_scale, mean0;
  int N, has_se;
};
static_assert(sizeof(PkSpec) <= 40 * sizeof(double), "spec block size");

}  // namespace

template <int DT, int NP, bool JAC>
__global__ void __cluster_dims__(PK_CL, 1, 1) __launch_bounds__(PK_THREADS, 1)
persist_rollout_kernel(const __grid_constant__ McpRollout r, const McpGpDev* __restrict__ gps, int Nmax, const double* __restrict__ nv0) {
  extern __shared__ __align__(16) double pk_smem[];
  const McpModel& mdl = r.model;
  const McpPolicy& pol = r.policy;
  const McpMeas& ms = r.meas;
  const McpNoise& nz = r.noise;
  const int M = r.M, H = r.H, Ds = mdl.Ds, Du = mdl.Du, E = mdl.E, D = mdl.D;
  const PkGeom G = pk_geom(Nmax, E);
  double* sKinv = pk_smem + G.o_kinv;   // [Wt][LD]  rows c0 .. c0 + Wt of K^-1 (= its columns)
  double* sKs = pk_smem + G.o_ks;       // [P][LD]   full K* rows of the batch for the current output
  double* sV = pk_smem + G.o_v;         // [P][Wt]
  double* sVx = pk_smem + G.o_vx;       // [warps / 2][64]  upper-K halves of the tiles of the matvec's partial last round
  double* sWt = pk_smem + G.o_wt;       // [warps][8][WS]
  double* sYt = pk_smem + G.o_yf;       // [E][8][WS]  features of own columns, feature-major: y_0 .. y_5, 1, 0
  double* sAl = pk_smem + G.o_al;       // [E][Wt]
  double* sFeat = pk_smem + G.o_feat;   // [P][8]    gp-input features of the batch's particles
  double* sPart = pk_smem + G.o_part;   // [CL][OWN][64]   (current output)
  double* sSum = pk_smem + G.o_sum;     // [OWN][E][64]
  PkOwn* own = reinterpret_cast<PkOwn*>(pk_smem + G.o_own);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, q = lane & 3;
  const int rank = (int)pk_rank(), c0 = rank * G.W;
  const int nclusters = gridDim.x / PK_CL, cid = (int)pk_cluster_id();
  const int items = PK_MT * (G.Wt / 8), items_full = items / PK_WARPS * PK_WARPS;
  const bool split_tail = items > items_full && 2 * (items - items_full) <= PK_WARPS;
  const bool drop = dropout_active(pol, nz);
  const double keep_scale = drop ? 1.0 / (1.0 - nz.p_dropout) : 1.0;
  const uint64_t seed = noise_seed(nz);

  PkSpec* sSpec = reinterpret_cast<PkSpec*>(pk_smem + G.o_spec);
  if (tid < E * 8) {
    const int e = tid >> 3, j = tid & 7;
    const McpGpSpec& s0 = gps[e].spec;
    PkSpec& sp = sSpec[e];
    sp.ils[j] = j < D ? s0.inv_ls[j] : 0.0;
    sp.w1[j] = (NP >= 1 && j < D) ? s0.poly_w2[0][0][j] : 0.0;
    sp.w2a[j] = (NP >= 2 && j < D) ? s0.poly_w2[1][0][j] : 0.0;
    sp.w2b[j] = (NP >= 2 && j < D) ? s0.poly_w2[1][1][j] : 0.0;
    if (j == 0) {
      sp.lambda = s0.has_se ? s0.lambda : 0.0;
      sp.o1 = NP >= 1 ? s0.poly_w2[0][0][MCP_MAX_D] : 0.0;
      sp.o2a = NP >= 2 ? s0.poly_w2[1][0][MCP_MAX_D] : 0.0;
      sp.o2b = NP >= 2 ? s0.poly_w2[1][1][MCP_MAX_D] : 0.0;
      sp.var_scale = gps[e].var_scale;
      sp.mean0 = s0.mean0;
      sp.N = gps[e].N;
      sp.has_se = s0.has_se;
    }
  }
  // ---- resident per-column data of every output: features [y, 1, 0] and alpha of own columns ----
  for (int i = tid; i < E * G.Wt; i += PK_THREADS) {
    const int e = i / G.Wt, c = i - e * G.Wt, cg = c0 + c;
    const McpGpDev& g = gps[e];
    const bool ok = c < G.W && cg < g.N;
#pragma unroll
    for (int j = 0; j < 8; j++) sYt[((size_t)e * 8 + j) * G.WS + c] = (ok && j < D) ? g.Xtr[(size_t)cg * D + j] : (ok && j == 6 ? 1.0 : 0.0);
    sAl[i] = ok ? g.alpha[cg] : 0.0;
  }
  // K^-1 slice of output e -> sKinv: one bulk copy per slice row, all completing on one mbarrier.  Every output has the same N
  // (persist_path_ok), so the zero padding (columns past N, rows past the slice) is written once here and never overwritten.
  const uint32_t bar = pk_smem_u32(pk_smem + G.o_bar);
  const int Nall = gps[0].N, Ncopy = Nall & ~1;   // bulk copies move multiples of 16 bytes; an odd last element goes by hand
  if (tid == 0) pk_mbar_init(bar, 1);
  for (int rr = warp; rr < G.Wt; rr += PK_WARPS) {
    const bool row_ok = rr < G.W && c0 + rr < Nall;
    for (int k = (row_ok ? Nall : 0) + lane; k < G.LD; k += 32) sKinv[(size_t)rr * G.LD + k] = 0.0;
  }
  __syncthreads();
  int rows_ok = G.W < Nall - c0 ? G.W : Nall - c0;
  rows_ok = rows_ok < 0 ? 0 : rows_ok;
  uint32_t slices_issued = 0;
  auto load_slice = [&](int e) {   // called by all threads, issued by warp 0
    slices_issued++;
    if (warp == 0) {
      const McpGpDev& g = gps[e];
      if (lane == 0) pk_mbar_expect_tx(bar, (uint32_t)rows_ok * (uint32_t)Ncopy * 8u);
      __syncwarp();
      for (int rr = lane; rr < rows_ok; rr += 32) {
        const double* src = g.Kinv + (size_t)(c0 + rr) * g.ld;
        double* dst = sKinv + (size_t)rr * G.LD;
        if (Ncopy > 0) pk_bulk_g2s(dst, src, (uint32_t)Ncopy * 8u, bar);
        if (Ncopy != Nall) dst[Nall - 1] = src[Nall - 1];
      }
    }
  };
  load_slice(0);

  for (int batch = cid; batch * PK_P < M; batch += nclusters) {
    const int base = batch * PK_P, cnt = min(PK_P, M - base);
    for (int i = tid; i < PK_P * G.LD; i += PK_THREADS) sKs[i] = 0.0;   // padded particles / columns stay zero
    for (int i = tid; i < PK_P * 8; i += PK_THREADS) sFeat[i] = 0.0;
    __syncthreads();
    pk_cluster_sync();  // nobody writes into a peer's buffers before the peer has cleared them

    for (int t = 0; t < H; t++) {
      // =============================================================== owner phase: x_t, u_t, features (64 threads per owned particle)
      {
        const int lp = warp >> 1, ltid = tid & 63, ml = rank + PK_CL * lp;   // local particle index in the batch
        if (lp < PK_OWN && ml < cnt) {
          const int m = base + ml;
          PkOwn& o = own[lp];
          const uint64_t pid = nz.particle_offset + (uint64_t)m;
          if (t > 0) {
            const int tp = t - 1;
            // ---- finish mean / variance / Jacobians: thread (e, j): j < D the Jacobian entries, j == 7 mean and variance ----
            if (ltid < E * 8) {
              const int e = ltid >> 3, j = ltid & 7;
              const PkSpec& sp = sSpec[e];
              const double* S = sSum + ((size_t)lp * E + e) * PK_NV;   // S[ch * 8 + feat]
              double x[DT];
#pragma unroll
              for (int jj = 0; jj < DT; jj++) x[jj] = sFeat[ml * 8 + jj];
              // Volterra factors at (x, x): L1 = o1 + sum w1 x^2, L2a, L2b likewise
              double L1 = sp.o1, La = sp.o2a, Lb = sp.o2b;
#pragma unroll
              for (int jj = 0; jj < DT; jj++) {
                L1 = fma(sp.w1[jj] * x[jj], x[jj], L1);
                La = fma(sp.w2a[jj] * x[jj], x[jj], La);
                Lb = fma(sp.w2b[jj] * x[jj], x[jj], Lb);
              }
              if (j == 7) {
                double kd = sp.lambda;
                if (NP >= 1) kd += L1;
                if (NP >= 2) kd += La * Lb;
                o.mean[e] = sp.mean0 + S[2 * 8 + 7];
                o.var[e] = sp.var_scale * (kd - S[3 * 8 + 7]);
              } else if (JAC && j < D) {
                const double xj = x[j];
                double dkd = 0.0;
                if (NP >= 1) dkd = 2.0 * sp.w1[j] * xj;
                if (NP >= 2) dkd = fma(2.0 * xj, sp.w2a[j] * Lb + sp.w2b[j] * La, dkd);
                const double il2 = -2.0 * sp.ils[j] * sp.ils[j];
                double ga = il2 * (xj * S[0 * 8 + 6] - S[0 * 8 + j]), gv = il2 * (xj * S[1 * 8 + 6] - S[1 * 8 + j]);
                if (NP >= 1) {
                  ga = fma(sp.w1[j], S[2 * 8 + j], ga);
                  gv = fma(sp.w1[j], S[3 * 8 + j], gv);
                }
                if (NP >= 2) {
                  ga = fma(sp.w2a[j], S[4 * 8 + j], ga);
                  gv = fma(sp.w2a[j], S[5 * 8 + j], gv);
                  ga = fma(sp.w2b[j], S[6 * 8 + j], ga);
                  gv = fma(sp.w2b[j], S[7 * 8 + j], gv);
                }
                o.jm[e][j] = ga;
                o.jv[e][j] = sp.var_scale * (dkd - 2.0 * gv);
              }
            }
            pk_pair_sync(1 + lp);
            // ---- reparameterised sample, integration, checkpoint, measurement model (first warp of the pair) ----
            if ((warp & 1) == 0) {
              double* xn = r.states + ((size_t)t * M + m) * Ds;
              double delta = 0.0, coef = 0.0;
              if (lane < E) {
                const double mu = o.mean[lane], v = o.var[lane];
                delta = mu;
                if (mdl.particle_pred) {
                  const double eps = nz.eps ? nz.eps[((size_t)tp * M + m) * E + lane] : rng_normal(seed, pid, tp, RNG_EPS, lane);
                  const double sd = sqrt(v);
                  delta = fma(sd, eps, mu);
                  coef = eps / (2.0 * sd);
                }
              }
              if (JAC && r.jac != nullptr) {
                double* jo = r.jac + ((size_t)tp * M + m) * E * D;
                const int n = E * D;
                for (int i0 = 0; i0 < n; i0 += 32) {
                  const int idx = i0 + lane, e = min(idx, n - 1) / D, d = min(idx, n - 1) - e * D;
                  const double ce = __shfl_sync(0xffffffffu, coef, e);
                  if (idx < n) jo[idx] = fma(ce, o.jv[e][d], o.jm[e][d]);
                }
              }
              double xnew = 0.0;  // lane j < Ds holds x_t[j]
              if (mdl.kind == 1) {
                for (int e = 0; e < E; e++) {
                  const double de = __shfl_sync(0xffffffffu, delta, e);
                  const int iv = mdl.vel_idx[e], ip = mdl.pos_idx[e];
                  if (lane == iv) xnew = o.x[iv] + de;
                  if (lane == ip) xnew = o.x[ip] + mdl.T * o.x[iv] + 0.5 * mdl.T * de;
                }
              } else {
                const double de = __shfl_sync(0xffffffffu, delta, min(lane, E - 1));
                if (lane < E) xnew = o.x[lane] + de;
              }
              if (lane < Ds) xn[lane] = xnew;
              double pin = xnew;
              if (ms.enabled) {
                double* pn = r.pol_in + ((size_t)t * M + m) * Ds;
                for (int i = 0; i < ms.n_pos; i++) {
                  const int ip = ms.pos_idx[i], iv = ms.vel_idx[i];
                  const double e = nz.meas_eps ? nz.meas_eps[((size_t)tp * M + m) * ms.n_pos + i] : rng_normal(seed, pid, tp, RNG_MEAS, i);
                  const double xp = __shfl_sync(0xffffffffu, xnew, ip);
                  const double np_new = fma(ms.std_pos[i], e, xp);
                  const double nv_old = o.nv[i];
                  const double nv_new = (np_new - o.pin[ip]) / ms.T;
                  const double mv_new = (ms.b0 * nv_new + ms.b1 * nv_old - ms.a1 * o.pin[iv]) / ms.a0;
                  __syncwarp();
                  if (lane == 0) o.nv[i] = nv_new;
                  if (lane == ip) pin = np_new;
                  if (lane == iv) pin = mv_new;
                }
                if (lane < Ds) pn[lane] = pin;
              }
              __syncwarp();
              if (lane < Ds) {
                o.x[lane] = xnew;
                o.pin[lane] = pin;
              }
            }
          } else {
            if (ltid < Ds) {
              o.x[ltid] = r.states[(size_t)m * Ds + ltid];
              o.pin[ltid] = ms.enabled ? r.pol_in[(size_t)m * Ds + ltid] : o.x[ltid];
            }
            if (ms.enabled && ltid < ms.n_pos) o.nv[ltid] = nv0[(size_t)m * ms.n_pos + ltid];
            if (ltid < pol.Dp) o.il[ltid] = exp(-pol.log_ls[ltid]);
          }
          pk_pair_sync(1 + lp);
          // ---- policy (64 threads over the basis functions), cf. policy_fwd_block_kernel ----
          if (ltid < pol.Dp) o.z[ltid] = policy_feature(pol, o.pin, t, ltid);
          pk_pair_sync(1 + lp);
          {
            double a[PK_MAX_DU];
#pragma unroll
            for (int k = 0; k < PK_MAX_DU; k++) a[k] = 0.0;
            // four consecutive basis functions per thread: their dropout draws are the four words of one Philox block
            for (int b0 = 4 * ltid; b0 < pol.nb; b0 += 256) {
              bool keep[4] = {true, true, true, true};
              if (drop) {
                if (nz.masks) {
#pragma unroll
                  for (int i = 0; i < 4; i++) keep[i] = b0 + i < pol.nb && nz.masks[((size_t)t * M + m) * pol.nb + b0 + i] != 0;
                } else {
                  const Philox4 pq = philox4x32_10(seed, (uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)t | ((uint32_t)RNG_MASK << 24), (uint32_t)(b0 >> 2));
#pragma unroll
                  for (int i = 0; i < 4; i++) keep[i] = (double)pq.v[i] * (1.0 / 4294967296.0) >= nz.p_dropout;
                }
              }
              // clamped indices instead of branches: all the (L2-latency) loads of the four basis functions issue together
              const double* cb[4];
              double d[4], wv[4][PK_MAX_DU];
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const int b = min(b0 + i, pol.nb - 1);
                cb[i] = pol.centers + (size_t)b * pol.Dp;
                d[i] = 0.0;
#pragma unroll
                for (int k = 0; k < PK_MAX_DU; k++) wv[i][k] = pol.W[(size_t)min(k, pol.Du - 1) * pol.nb + b];
              }
#pragma unroll 4
              for (int j = 0; j < pol.Dp; j++) {
                const double zj = o.z[j], ilj = o.il[j];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                  const double rr = (zj - cb[i][j]) * ilj;
                  d[i] = fma(rr, rr, d[i]);
                }
              }
#pragma unroll
              for (int i = 0; i < 4; i++) {
                double h = (b0 + i < pol.nb && keep[i]) ? exp(-d[i]) * keep_scale : 0.0;
#pragma unroll
                for (int k = 0; k < PK_MAX_DU; k++)
                  if (k < pol.Du) a[k] = fma(wv[i][k], h, a[k]);
              }
            }
#pragma unroll
            for (int k = 0; k < PK_MAX_DU; k++)
              if (k < pol.Du) {
                const double vv = warp_sum(a[k]);
                if (lane == 0) o.part[warp & 1][k] = vv;
              }
          }
          pk_pair_sync(1 + lp);
          if (ltid < pol.Du) {
            double vv = o.part[0][ltid] + o.part[1][ltid];
            if (pol.has_bias) vv += pol.bias[ltid];
            if (pol.squash) vv = pol.u_max[ltid] * tanh(vv / pol.u_max[ltid]);
            r.inputs[((size_t)t * M + m) * Du + ltid] = vv;
            o.u[ltid] = vv;
          }
          pk_pair_sync(1 + lp);
          // ---- gp-input features of (x_t, u_t), broadcast to the cluster ----
          if (t < H - 1 && ltid < 8 * PK_CL) {
            const int j = ltid & 7, dst = ltid >> 3;
            double f = 0.0;
            if (j < D) {
              if (mdl.use_trig) {
                if (j < mdl.n_na) f = o.x[mdl.na_idx[j]];
                else if (j < mdl.n_na + mdl.n_a) f = sin(o.x[mdl.a_idx[j - mdl.n_na]]);
                else if (j < mdl.n_na + 2 * mdl.n_a) f = cos(o.x[mdl.a_idx[j - mdl.n_na - mdl.n_a]]);
                else f = o.u[j - mdl.n_na - 2 * mdl.n_a];
              } else {
                f = (j < mdl.Ds) ? o.x[j] : o.u[j - mdl.Ds];
              }
            }
            pk_st_remote(&sFeat[ml * 8 + j], (uint32_t)dst, f);
          }
        }
      }
      if (t == H - 1) break;
      pk_cluster_sync();  // features of step t visible everywhere

      for (int e = 0; e < E; e++) {
        // =========================================================== K* entries of own columns for all particles -> every CTA's sKs
        const PkSpec& sp = sSpec[e];
        const int Ne = sp.N;
        // (particle, own column) pairs spread over all threads, a warp's lanes on consecutive columns; columns past N keep the zero
        // the batch-start clear gave them
        {
          const int Wv = max(0, min(G.W, Ne - c0)), nitems = Wv * cnt;
          for (int i = tid; i < nitems; i += PK_THREADS) {
            const int ml = i / Wv, c = i - ml * Wv;
            const double* x = sFeat + ml * 8;
            double y[DT], d2 = 0.0;
#pragma unroll
            for (int j = 0; j < DT; j++) {
              y[j] = sYt[((size_t)e * 8 + j) * G.WS + c];
              const double tt = (x[j] - y[j]) * sp.ils[j];
              d2 = fma(tt, tt, d2);
            }
            double kv = sp.has_se ? sp.lambda * exp(-d2) : 0.0;
            if (NP >= 1) {
              double L1 = sp.o1;
#pragma unroll
              for (int j = 0; j < DT; j++) L1 = fma(sp.w1[j] * x[j], y[j], L1);
              kv += L1;
            }
            if (NP >= 2) {
              double La = sp.o2a, Lb = sp.o2b;
#pragma unroll
              for (int j = 0; j < DT; j++) {
                La = fma(sp.w2a[j] * x[j], y[j], La);
                Lb = fma(sp.w2b[j] * x[j], y[j], Lb);
              }
              kv = fma(La, Lb, kv);
            }
            const double* dstp = &sKs[(size_t)ml * G.LD + c0 + c];
#pragma unroll
            for (int dst = 0; dst < PK_CL; dst++) pk_st_remote(dstp, (uint32_t)dst, kv);
          }
        }
        pk_cluster_sync();  // K* rows complete in every CTA
        pk_mbar_wait(bar, (slices_issued - 1) & 1);   // this output's K^-1 slice has landed (a no-op once E = 1 has its only slice)

        // =========================================================== V[:, own columns] = K* K^-1[:, own columns] on FP64 DMMA
        {
          // tile (mt, ct) of V: full rounds of eight tiles, one per warp; a last partial round of <= 4 tiles is split in two K halves
          // over warp pairs (the upper halves go to sVx and are added by the consumer) so that no warp idles through a whole round
          auto tile = [&](int it, int k_lo, int k_hi, double* upper) {   // upper: the 8 x 8 block of sVx for an upper K half, else nullptr
            const int mt = it % PK_MT, ct = it / PK_MT, row = mt * 8 + gq;
            const double* A = sKs + (size_t)min(row, PK_P - 1) * G.LD + q;
            const double* B = sKinv + (size_t)(ct * 8 + gq) * G.LD + q;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            for (int k = k_lo; k < k_hi; k += 8) {   // two independent accumulation chains
              dmma884(a0, a1, A[k], B[k]);
              dmma884(b0, b1, A[k + 4], B[k + 4]);
            }
            if (upper != nullptr || row < PK_P) {
              double* o2 = upper != nullptr ? upper + gq * 8 + 2 * q : sV + (size_t)row * G.Wt + ct * 8 + 2 * q;
              o2[0] = a0 + b0;
              o2[1] = a1 + b1;
            }
          };
          for (int it = warp; it < items_full; it += PK_WARPS) tile(it, 0, G.Kc, nullptr);
          if (split_tail) {
            const int it = items_full + (warp >> 1);
            if (it < items) {
              if (warp & 1) tile(it, G.Kc / 2, G.Kc, sVx + (size_t)(warp >> 1) * 64);
              else tile(it, 0, G.Kc / 2, nullptr);
            }
          } else if (items_full + warp < items) {
            tile(items_full + warp, 0, G.Kc, nullptr);
          }
        }
        __syncthreads();
        if (E > 1) load_slice((e + 1) % E);  // the slice buffer is free: stream the next output's slice in behind the reduce

        // =========================================================== own columns' share of the factored posterior sums
        for (int ml = warp; ml < cnt; ml += PK_WARPS) {
          double* wt = sWt + (size_t)warp * 8 * G.WS;
          double x[DT], xw1[DT], xw2a[DT], xw2b[DT];
#pragma unroll
          for (int j = 0; j < DT; j++) {
            x[j] = sFeat[ml * 8 + j];
            xw1[j] = sp.w1[j] * x[j];
            xw2a[j] = sp.w2a[j] * x[j];
            xw2b[j] = sp.w2b[j] * x[j];
          }
          // channel weights per column: [a e, v e, a, v, a L2b, v L2b, a L2a, v L2a]
          for (int c = lane; c < G.Wt; c += 32) {
            double y[DT];
#pragma unroll
            for (int j = 0; j < DT; j++) y[j] = sYt[((size_t)e * 8 + j) * G.WS + c];
            const double a = sAl[e * G.Wt + c];
            double vn = sV[(size_t)ml * G.Wt + c];
            if (split_tail) {
              const int it = (c >> 3) * PK_MT + (ml >> 3);
              if (it >= items_full) vn += sVx[(size_t)(it - items_full) * 64 + (ml & 7) * 8 + (c & 7)];
            }
            const double kv = (c < G.W && c0 + c < G.Kc) ? sKs[(size_t)ml * G.LD + c0 + c] : 0.0;
            double poly = 0.0, L2a = 0.0, L2b = 0.0;
            if (NP >= 1) {
              double L1 = sp.o1;
#pragma unroll
              for (int j = 0; j < DT; j++) L1 = fma(xw1[j], y[j], L1);
              poly = L1;
            }
            if (NP >= 2) {
              L2a = sp.o2a;
              L2b = sp.o2b;
#pragma unroll
              for (int j = 0; j < DT; j++) {
                L2a = fma(xw2a[j], y[j], L2a);
                L2b = fma(xw2b[j], y[j], L2b);
              }
              poly = fma(L2a, L2b, poly);
            }
            const double ev = kv - poly;
            const bool live = c < G.W && c0 + c < Ne;   // padded columns carry no weight (their alpha, V and K* are zero anyway)
            wt[0 * G.WS + c] = live ? a * ev : 0.0;
            wt[1 * G.WS + c] = live ? vn * ev : 0.0;
            wt[2 * G.WS + c] = live ? a : 0.0;
            wt[3 * G.WS + c] = live ? vn : 0.0;
            wt[4 * G.WS + c] = live ? a * L2b : 0.0;
            wt[5 * G.WS + c] = live ? vn * L2b : 0.0;
            wt[6 * G.WS + c] = live ? a * L2a : 0.0;
            wt[7 * G.WS + c] = live ? vn * L2a : 0.0;
          }
          __syncwarp();
          // tile[channel][feature] = sum_c weight[channel][c] feature[c][.]  with features [y_0..y_5, 1, k_c]
          // (feature 7 is read from the K* row itself: columns past this CTA's slice carry zero weight, the row padding is zero)
          double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
          const double* bbase = (gq == 7) ? sKs + (size_t)ml * G.LD + c0 + q : sYt + ((size_t)e * 8 + gq) * G.WS + q;
          const double* abase = wt + gq * G.WS + q;
          for (int k = 0; k < G.Wt; k += 8) {   // Wt is a multiple of 8: two accumulation chains
            dmma884(t0, t1, abase[k], bbase[k]);
            dmma884(u0, u1, abase[k + 4], bbase[k + 4]);
          }
          t0 += u0;
          t1 += u1;
          __syncwarp();
          // lane (gq, q) holds tile[gq][2q], tile[gq][2q+1]: send to the particle's owner
          const int owner = ml % PK_CL, lp = ml / PK_CL;
          double* dstp = sPart + ((size_t)(rank * PK_OWN + lp) * PK_NV) + gq * 8 + 2 * q;
          pk_st_remote(dstp, (uint32_t)owner, t0);
          pk_st_remote(dstp + 1, (uint32_t)owner, t1);
        }
        pk_cluster_sync();  // partial tiles delivered; every CTA is done with the K* rows of this output
        // owners: the eight partial tiles of each owned particle, added in rank order (sPart is rewritten two cluster barriers later)
        for (int i = tid; i < PK_OWN * PK_NV; i += PK_THREADS) {
          const int lp = i / PK_NV, idx = i - lp * PK_NV;
          double a = 0.0;
#pragma unroll
          for (int src = 0; src < PK_CL; src++) a += sPart[(size_t)(src * PK_OWN + lp) * PK_NV + idx];
          sSum[((size_t)lp * E + e) * PK_NV + idx] = a;
        }
      }
      __syncthreads();  // the owners' threads read sums other threads wrote
    }
    __syncthreads();
  }
  pk_mbar_wait(bar, (slices_issued - 1) & 1);  // no bulk copy in flight at exit
  pk_cluster_sync();  // no CTA leaves while a peer may still store into its shared memory
}

// ---- host side --------------------------------------------------------------------------------------------------------------
// co-resident clusters of this kernel at a given shared-memory size (asked of the driver once per device and size; all template
// instances use the same registers and shared memory).  scripts/probe/cluster_occ.cu: 15 above 113 KB with this register count.
static int pk_max_clusters(size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, size_t>, int> cache;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return 15;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({dev, smem});
  if (it != cache.end()) return it->second;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int n = 15 * sms / 148;
  static bool cfg_[MCP_MAX_DEVICES] = {};
  if (ensure_dynamic_smem(cfg_, persist_rollout_kernel<6, 0, true>, 227 * 1024) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PK_CL * 64);
    cfg.blockDim = dim3(PK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = PK_CL;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, persist_rollout_kernel<6, 0, true>, &cfg) == cudaSuccess && q >= 1) n = q;
    else (void)cudaGetLastError();
  }
  if (n < 1) n = 1;
  cache[{dev, smem}] = n;
  return n;
}

bool persist_path_ok(const McpRollout* r) {
  const char* off = getenv("MCPILCO_NO_PERSIST");
  if (off != nullptr && off[0] == '1') return false;
  off = getenv("MCPILCO_NO_SMALL_PATH");  // "per-step kernels only" switches both small-shape paths off
  if (off != nullptr && off[0] == '1') return false;
  const int Mg = r->M_global > 0 ? r->M_global : r->M;
  if (Mg > 2048 || r->H < 2 || r->model.D > 6 || r->model.E > PK_MAX_E || r->model.Ds > PK_MAX_DS || r->model.Du > PK_MAX_DU ||
      r->policy.Dp > PK_MAX_DP || (r->meas.enabled && r->meas.n_pos > PK_MAX_NPOS))
    return false;
  int np = -1, nmax = 1;
  for (int e = 0; e < r->model.E; e++) {
    const McpGpSpec& s = r->gps[e].spec;
    if (!s.has_se || s.n_poly > 2 || r->gps[e].ozaki_slices != 0) return false;
    for (int p = 0; p < s.n_poly; p++)
      if (s.poly_deg[p] != p + 1) return false;
    if (np >= 0 && s.n_poly != np) return false;
    np = s.n_poly;
    if (r->gps[e].ld_kinv % 2 != 0 || ((uintptr_t)r->gps[e].Kinv % 16) != 0) return false;
    if (r->gps[e].N != r->gps[0].N) return false;  // one zero padding of the slice buffer serves every output
    nmax = r->gps[e].N > nmax ? r->gps[e].N : nmax;
  }
  const size_t smem = pk_geom(nmax, r->model.E).doubles * sizeof(double);
  if (smem > 227 * 1024) return false;
  // Eligible.  Measured on B200 against the fused two-launch-per-step path (profiles/r02_real_shape_paths.txt): with Volterra terms
  // in the kernel (C1) this path is faster at every N that fits; with SE-only outputs (C2, C3) it is faster up to N ~ 250 and a few
  // per cent slower at N = 300, where the per-step DMMA work outgrows the launch latency it saves.  MCPILCO_PERSIST=1 forces it.
  const char* force = getenv("MCPILCO_PERSIST");
  if (force != nullptr && force[0] == '1') return true;
  // ... and only while one pass covers the rollout: a cluster that has to walk several particle batches one after the other loses to
  // the per-step path, which spreads all particles over the GPU at every step (scripts/path_vs_particles.py: 0.88-0.94 of the fused
  // path's time at 400 particles, 1.0-1.2 at 800, 1.3-1.8 at 2048)
  if (cdiv(Mg, PK_P) > pk_max_clusters(smem)) return false;  // the GLOBAL count: shards must take the unsharded rollout's path
  return np >= 1 || nmax <= 240;
}

size_t persist_path_doubles(int E) { return (sizeof(McpGpDev) * (size_t)E + 7) / 8 + 64; }

int rollout_fwd_persist(const McpRollout* r, const double* nv0, double* scratch, size_t scratch_doubles, cudaStream_t st) {
  const int M = r->M, E = r->model.E, D = r->model.D;
  MCP_CHECK_ARG(scratch_doubles >= persist_path_doubles(E), "rollout (persistent path): workspace too small");
  int nmax = 1;
  for (int e = 0; e < E; e++) nmax = r->gps[e].N > nmax ? r->gps[e].N : nmax;
  McpGpDev* tab = reinterpret_cast<McpGpDev*>(scratch);
  MCP_CUDA(gpdev_upload(tab, r->gps, E, st));
  const PkGeom G = pk_geom(nmax, E);
  const size_t smem = G.doubles * sizeof(double);
  int clusters = pk_max_clusters(smem);  // co-resident clusters of 8 CTAs at this shared-memory size; more would only queue
  const int batches = cdiv(M, PK_P);
  if (clusters > batches) clusters = batches;
  const int np = r->gps[0].spec.n_poly;
  const bool jac = r->need_grad != 0;
#define MCP_PK(DT_, NP_, JAC_)                                                                                              \
  do {                                                                                                                      \
    static bool cfg_[MCP_MAX_DEVICES] = {};                                                                                 \
    MCP_CUDA(ensure_dynamic_smem(cfg_, persist_rollout_kernel<DT_, NP_, JAC_>, 227 * 1024)); /* once per device: the size varies with N */                                 \
    persist_rollout_kernel<DT_, NP_, JAC_><<<clusters * PK_CL, PK_THREADS, smem, st>>>(*r, tab, nmax, nv0);                 \
  } while (0)
#define MCP_PK2(DT_, NP_) do { if (jac) MCP_PK(DT_, NP_, true); else MCP_PK(DT_, NP_, false); } while (0)
  if (D <= 4) { if (np == 0) MCP_PK2(4, 0); else if (np == 1) MCP_PK2(4, 1); else MCP_PK2(4, 2); }
  else { if (np == 0) MCP_PK2(6, 0); else if (np == 1) MCP_PK2(6, 1); else MCP_PK2(6, 2); }
#undef MCP_PK2
#undef MCP_PK
  MCP_LAUNCH_CHECK();
  count_launch();
  return MCP_OK;
}

}  // namespace mcp
