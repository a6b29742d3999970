// Fused forward for SMALL rollouts — the reference's real configurations (M = 200..400 particles, N = 60..500 training points,
// cart-pole-sized inputs): ~1 GFLOP per rollout, so the time is the length of the dependent-kernel chain, not arithmetic.
// Everything of a time step that belongs to ONE particle is one kernel; only the contraction with K^-1, which shares K^-1
// across particles, stays a (batched) tile GEMM:
//
//   small_step_kernel(t):   [t > 0]  reduce of step t-1 for all E outputs (factored Jacobian sums, cf. posterior_reduce_fast_kernel),
//                                    reparameterised sample, integration, 4PMS measurement model  ->  x_t, checkpoint J_{t-1}
//                           policy(x_t) with dropout/squashing -> u_t;  [t < H-1] gp-input features and the K* rows of all outputs
//   small_gemm_kernel(t):   V_e = K*_e K_e^-1  for all outputs in one launch (32 x 32 DMMA tiles, blockIdx.z = output)
//
// i.e. 2 dependent launches per step instead of 5, no per-output stream fan-out.  One 128-thread block per particle.
// Same arithmetic as the per-step kernels (policy_fwd_block_kernel, cov_fast_kernel, posterior_reduce_fast_kernel,
// integrate_warp_kernel); selected from the GLOBAL particle count, so shards stay bit-identical to the unsharded rollout.
// Covers kernels of the "fast" family: D <= 6 with up to two Volterra polynomial terms, or D <= 8 with at most one.
#include <stdlib.h>

#include "mcp_dgemm.cuh"
#include "mcp_gpdev.cuh"
#include "mcp_kfn.cuh"
#include "mcp_rollout_dev.cuh"

namespace mcp {

__global__ void gpdev_store_kernel(const __grid_constant__ McpGpDev g, McpGpDev* __restrict__ dst) {
  const int n = (int)(sizeof(McpGpDev) / 8);  // the struct is a multiple of 8 bytes (doubles and pointers)
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&g);
  unsigned long long* out = reinterpret_cast<unsigned long long*>(dst);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = src[i];
}

// Programmatic dependent launch: the two kernels of a step are launched with programmatic stream serialisation, so the next grid's
// launch latency overlaps the tail of the current one.  Every kernel of the chain first waits for its predecessor to complete and
// flush (pdl_wait; a no-op without the launch attribute).  The successor is released implicitly when the blocks exit: an explicit
// early griddepcontrol.launch_dependents (at the top, or before the last phase) was measured 25-40 % SLOWER at C1/C3 — the
// early-resident successor blocks compete with the running grid — while the implicit form gains 8-12 %.

// ---- batched V_e = K*_e Kinv_e^T (Kinv symmetric), 32 x 32 tiles, 128 threads ----
// The contraction is short (K = N = 300..400) and there are only a few hundred tiles, so what matters is that the L2 round trip of the
// operand tiles is covered: k-blocks of 32 in a 3-stage cp.async ring (two blocks = 1000 DMMA-pipe cycles in flight; the generic
// 16-wide 3-stage ring of mcp_dgemm.cuh keeps ~130 in flight and made this kernel wait for L2 at every k-block: 13.7 us per step at
// C2 against a 4 us FP64-pipe floor).  Same accumulation order over k as gemm_mainloop: bit-identical results.
constexpr int SG_BK = 32, SG_LDS = SG_BK + 4, SG_STAGES = 3;   // row stride = 4 (mod 16) doubles: conflict-free fragment loads; 55 KB: four CTAs per SM (the 546 tiles of the UR5 step are one wave)
constexpr size_t SG_SMEM_BYTES = (size_t)SG_STAGES * 64 * SG_LDS * sizeof(double);
__device__ __forceinline__ void sg_load_tile(double* __restrict__ s, const double* __restrict__ G, int ld, int r0, int nrows, int k0, int K,
                                             int tid) {
#pragma unroll
  for (int c = tid; c < 32 * (SG_BK / 2); c += 128) {
    const int row = c / (SG_BK / 2), kc = (c % (SG_BK / 2)) * 2;
    const int gr = r0 + row, rem = K - (k0 + kc);
    const int bytes = (gr < nrows && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
    cp_async16(s + row * SG_LDS + kc, bytes ? (G + (size_t)gr * ld + k0 + kc) : G, bytes);
  }
}
__global__ void __launch_bounds__(128) small_gemm_kernel(const McpGpDev* __restrict__ gps, int M, const double* __restrict__ Ks,
                                                         double* __restrict__ V, int ldk, size_t gp_stride) {
  extern __shared__ __align__(16) double smem[];
  pdl_wait();
  const McpGpDev& g = gps[blockIdx.z];
  const int N = g.N, m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  if (n0 >= N) return;
  double acc[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* A = Ks + blockIdx.z * gp_stride;
  {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, q = lane & 3;
    const int wm0 = (warp % 2) * 16, wn0 = (warp / 2) * 16;
    double* As = smem;
    double* Bs = smem + SG_STAGES * 32 * SG_LDS;
    const int KT = (N + SG_BK - 1) / SG_BK;
#pragma unroll
    for (int st = 0; st < SG_STAGES - 1; st++) {
      if (st < KT) {
        sg_load_tile(As + st * 32 * SG_LDS, A, ldk, m0, M, st * SG_BK, N, tid);
        sg_load_tile(Bs + st * 32 * SG_LDS, g.Kinv, g.ld, n0, N, st * SG_BK, N, tid);
      }
      cp_async_commit();
    }
    for (int kt = 0; kt < KT; kt++) {
      cp_async_wait<SG_STAGES - 2>();
      __syncthreads();
      {
        const int nk = kt + SG_STAGES - 1;
        if (nk < KT) {
          const int st = nk % SG_STAGES;
          sg_load_tile(As + st * 32 * SG_LDS, A, ldk, m0, M, nk * SG_BK, N, tid);
          sg_load_tile(Bs + st * 32 * SG_LDS, g.Kinv, g.ld, n0, N, nk * SG_BK, N, tid);
        }
        cp_async_commit();
      }
      const double* as = As + (kt % SG_STAGES) * 32 * SG_LDS + (wm0 + gq) * SG_LDS + q;
      const double* bs = Bs + (kt % SG_STAGES) * 32 * SG_LDS + (wn0 + gq) * SG_LDS + q;
#pragma unroll
      for (int ks = 0; ks < SG_BK / 4; ks++) {
        double a[2], b[2];
#pragma unroll
        for (int i = 0; i < 2; i++) a[i] = as[i * 8 * SG_LDS + ks * 4];
#pragma unroll
        for (int j = 0; j < 2; j++) b[j] = bs[j * 8 * SG_LDS + ks * 4];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    cp_async_wait<0>();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm0 = (warp % 2) * 16, wn0 = (warp / 2) * 16, gq = lane >> 2, q = lane & 3;
  double* C = V + blockIdx.z * gp_stride;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int r = m0 + wm0 + 8 * i + gq;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int c = n0 + wn0 + 8 * j + 2 * q;
      if (c < N) C[(size_t)r * ldk + c] = acc[i][j][0];
      if (c + 1 < N) C[(size_t)r * ldk + c + 1] = acc[i][j][1];
    }
  }
}

constexpr int SS_THREADS = 128;

template <int DT, int NP, bool JAC>
__global__ void __launch_bounds__(SS_THREADS, 3) small_step_kernel(const __grid_constant__ McpRollout r, const McpGpDev* __restrict__ gps, int t,
                                                                double* __restrict__ Ks, const double* __restrict__ V, int ldk,
                                                                size_t gp_stride, double* __restrict__ Xs, double* __restrict__ nv) {
  const McpModel& mdl = r.model;
  const McpPolicy& pol = r.policy;
  const McpMeas& ms = r.meas;
  const McpNoise& nz = r.noise;
  const int M = r.M, H = r.H, Ds = mdl.Ds, Du = mdl.Du, E = mdl.E, D = mdl.D;
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double s_x[MCP_MAX_DS], s_pin[MCP_MAX_DS], s_feat[DT], s_u[MCP_MAX_DU];
  __shared__ double s_w[4][8 * DT + 4], s_sum[8 * DT + 4];
  __shared__ double s_mean[MCP_MAX_E], s_var[MCP_MAX_E], s_jm[MCP_MAX_E][DT], s_jv[MCP_MAX_E][DT];
  __shared__ double s_il[MCP_MAX_DP], s_z[MCP_MAX_DP], s_part[4][MCP_MAX_DU];

  pdl_wait();
  if (t > 0) {
    // ------------------------------------------------------------------ post(t-1): posterior of step t-1 -> x_t
    const int tp = t - 1;
    if (tid < DT) s_feat[tid] = tid < D ? Xs[(size_t)m * D + tid] : 0.0;
    if (tid < Ds) s_x[tid] = r.states[((size_t)tp * M + m) * Ds + tid];
    __syncthreads();
    double x[DT];
#pragma unroll
    for (int j = 0; j < DT; j++) x[j] = s_feat[j];
    for (int e = 0; e < E; e++) {
      const McpGpDev& g = gps[e];
      const McpGpSpec& s = g.spec;
      const int N = g.N;
      double xs[DT], xw1[DT], xw2a[DT], xw2b[DT];
#pragma unroll
      for (int j = 0; j < DT; j++) {
        xs[j] = x[j] * s.inv_ls[j];
        xw1[j] = NP >= 1 ? s.poly_w2[0][0][j] * x[j] : 0.0;
        xw2a[j] = NP >= 2 ? s.poly_w2[1][0][j] * x[j] : 0.0;
        xw2b[j] = NP >= 2 ? s.poly_w2[1][1][j] * x[j] : 0.0;
      }
      double mu = 0.0, q = 0.0, E0a = 0.0, E0v = 0.0;
      double E1a[DT], E1v[DT], C1a[DT], C1v[DT], C2a0[DT], C2v0[DT], C2a1[DT], C2v1[DT];
#pragma unroll
      for (int j = 0; j < DT; j++) E1a[j] = E1v[j] = C1a[j] = C1v[j] = C2a0[j] = C2v0[j] = C2a1[j] = C2v1[j] = 0.0;
      const double* v = V + e * gp_stride + (size_t)m * ldk;
      const double* kr = Ks + e * gp_stride + (size_t)m * ldk;  // the K* row of step t-1 (written by this block's pre(t-1) phase)
      for (int n = tid; n < N; n += SS_THREADS) {
        double y[DT];
        KFn<DT>::load(y, g.Xtr + (size_t)n * D, D);
        const double a = g.alpha[n], vn = v[n], kv = kr[n];
        // the kernel value is not evaluated again (cf. posterior_reduce_fast_kernel): polynomial factors from the particle-scaled
        // weights, squared-exponential part e_n = k_n - poly_n
        double poly = 0.0, L2a = 0.0, L2b = 0.0;
        if (NP >= 1) {
          double L1 = s.poly_w2[0][0][MCP_MAX_D];
#pragma unroll
          for (int j = 0; j < DT; j++) L1 = fma(xw1[j], y[j], L1);
          poly = L1;
        }
        if (NP >= 2) {
          L2a = s.poly_w2[1][0][MCP_MAX_D];
          L2b = s.poly_w2[1][1][MCP_MAX_D];
#pragma unroll
          for (int j = 0; j < DT; j++) {
            L2a = fma(xw2a[j], y[j], L2a);
            L2b = fma(xw2b[j], y[j], L2b);
          }
          poly = fma(L2a, L2b, poly);
        }
        const double ev = kv - poly;
        mu = fma(a, kv, mu);
        q = fma(vn, kv, q);
        if (JAC) {
          const double ta = a * ev, tv = vn * ev;
          E0a += ta;
          E0v += tv;
          const double ua0 = a * L2b, uv0 = vn * L2b, ua1 = a * L2a, uv1 = vn * L2a;
#pragma unroll
          for (int j = 0; j < DT; j++) {
            E1a[j] = fma(ta, y[j], E1a[j]);
            E1v[j] = fma(tv, y[j], E1v[j]);
            if (NP >= 1) { C1a[j] = fma(a, y[j], C1a[j]); C1v[j] = fma(vn, y[j], C1v[j]); }
            if (NP >= 2) {
              C2a0[j] = fma(ua0, y[j], C2a0[j]); C2v0[j] = fma(uv0, y[j], C2v0[j]);
              C2a1[j] = fma(ua1, y[j], C2a1[j]); C2v1[j] = fma(uv1, y[j], C2v1[j]);
            }
          }
        }
      }
      // block sums: [mu, q, E0a, E0v | E1a | E1v | C1a | C1v | C2a0 | C2v0 | C2a1 | C2v1] — warp shuffles, then one pass through
      // shared memory (two barriers for all 4 + 8 DT values)
      {
        constexpr int NV = 4 + 8 * DT;
        auto put = [&](int k, double val) {
          val = warp_sum(val);
          if (lane == 0) s_w[warp][k] = val;
        };
        put(0, mu);
        put(1, q);
        if (JAC) {
          put(2, E0a);
          put(3, E0v);
#pragma unroll
          for (int j = 0; j < DT; j++) {
            put(4 + j, E1a[j]);
            put(4 + DT + j, E1v[j]);
            if (NP >= 1) { put(4 + 2 * DT + j, C1a[j]); put(4 + 3 * DT + j, C1v[j]); }
            if (NP >= 2) {
              put(4 + 4 * DT + j, C2a0[j]); put(4 + 5 * DT + j, C2v0[j]);
              put(4 + 6 * DT + j, C2a1[j]); put(4 + 7 * DT + j, C2v1[j]);
            }
          }
        }
        __syncthreads();
        if (tid < NV) s_sum[tid] = (s_w[0][tid] + s_w[1][tid]) + (s_w[2][tid] + s_w[3][tid]);
      }
      __syncthreads();
      // finalise in parallel: thread 0 the mean / variance, thread j < D the j-th Jacobian entries
      if (tid == 0) {
        s_mean[e] = s.mean0 + s_sum[0];
        s_var[e] = g.var_scale * (KFn<DT>::kdiag(s, x) - s_sum[1]);
      }
      if (JAC && tid < DT && tid < D) {
        const int j = tid;
        const double xj = s_feat[j];
        // d k(x,x) / dx_j of the Volterra terms: deg 1: 2 w1_j x_j ; deg 2: 2 x_j (w2a_j L2b(x,x) + w2b_j L2a(x,x))
        double dkd = 0.0;
        if (NP >= 1) dkd = 2.0 * s.poly_w2[0][0][j] * xj;
        if (NP >= 2) {
          double La = s.poly_w2[1][0][MCP_MAX_D], Lb = s.poly_w2[1][1][MCP_MAX_D];
#pragma unroll
          for (int jj = 0; jj < DT; jj++) {
            La = fma(s.poly_w2[1][0][jj] * x[jj], x[jj], La);
            Lb = fma(s.poly_w2[1][1][jj] * x[jj], x[jj], Lb);
          }
          dkd = fma(2.0 * xj, s.poly_w2[1][0][j] * Lb + s.poly_w2[1][1][j] * La, dkd);
        }
        const double il2 = -2.0 * s.inv_ls[j] * s.inv_ls[j];
        double ga = il2 * (xj * s_sum[2] - s_sum[4 + j]), gv = il2 * (xj * s_sum[3] - s_sum[4 + DT + j]);
        if (NP >= 1) {
          ga = fma(s.poly_w2[0][0][j], s_sum[4 + 2 * DT + j], ga);
          gv = fma(s.poly_w2[0][0][j], s_sum[4 + 3 * DT + j], gv);
        }
        if (NP >= 2) {
          ga = fma(s.poly_w2[1][0][j], s_sum[4 + 4 * DT + j], ga);
          gv = fma(s.poly_w2[1][0][j], s_sum[4 + 5 * DT + j], gv);
          ga = fma(s.poly_w2[1][1][j], s_sum[4 + 6 * DT + j], ga);
          gv = fma(s.poly_w2[1][1][j], s_sum[4 + 7 * DT + j], gv);
        }
        s_jm[e][j] = ga;
        s_jv[e][j] = g.var_scale * (dkd - 2.0 * gv);
      }
      __syncthreads();
    }
    // ---- reparameterised sample, integration, checkpoint, measurement model (one warp; cf. integrate_warp_kernel) ----
    if (warp == 0) {
      double* xn = r.states + ((size_t)t * M + m) * Ds;
      double delta = 0.0, coef = 0.0;
      if (lane < E) {
        const double mu = s_mean[lane], v = s_var[lane];
        delta = mu;
        if (mdl.particle_pred) {
          const double eps = nz.eps ? nz.eps[((size_t)tp * M + m) * E + lane] : rng_normal(noise_seed(nz), nz.particle_offset + (uint64_t)m, tp, RNG_EPS, lane);
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
          if (idx < n) jo[idx] = fma(ce, s_jv[e][d], s_jm[e][d]);
        }
      }
      double xnew = 0.0;  // lane j < Ds holds x_t[j]
      if (mdl.kind == 1) {
        // vel' = vel + delta, pos' = pos + T vel + T/2 delta; built per output lane, gathered per state lane
        for (int e = 0; e < E; e++) {
          const double de = __shfl_sync(0xffffffffu, delta, e);
          const int iv = mdl.vel_idx[e], ip = mdl.pos_idx[e];
          if (lane == iv) xnew = s_x[iv] + de;
          if (lane == ip) xnew = s_x[ip] + mdl.T * s_x[iv] + 0.5 * mdl.T * de;
        }
      } else {
        const double de = __shfl_sync(0xffffffffu, delta, min(lane, E - 1));
        if (lane < E) xnew = s_x[lane] + de;
      }
      if (lane < Ds) xn[lane] = xnew;
      if (ms.enabled) {
        const double* pp = r.pol_in + ((size_t)tp * M + m) * Ds;
        double* pn = r.pol_in + ((size_t)t * M + m) * Ds;
        double pin = xnew;
        for (int i = 0; i < ms.n_pos; i++) {
          const int ip = ms.pos_idx[i], iv = ms.vel_idx[i];
          const double e = nz.meas_eps ? nz.meas_eps[((size_t)tp * M + m) * ms.n_pos + i]
                                       : rng_normal(noise_seed(nz), nz.particle_offset + (uint64_t)m, tp, RNG_MEAS, i);
          const double xp = __shfl_sync(0xffffffffu, xnew, ip);
          const double np_new = fma(ms.std_pos[i], e, xp);
          const double nv_old = nv[(size_t)m * ms.n_pos + i];
          const double nv_new = (np_new - pp[ip]) / ms.T;
          const double mv_new = (ms.b0 * nv_new + ms.b1 * nv_old - ms.a1 * pp[iv]) / ms.a0;
          if (lane == 0) nv[(size_t)m * ms.n_pos + i] = nv_new;
          if (lane == ip) pin = np_new;
          if (lane == iv) pin = mv_new;
        }
        if (lane < Ds) { pn[lane] = pin; s_pin[lane] = pin; }
      }
      if (lane < Ds) { s_x[lane] = xnew; if (!ms.enabled) s_pin[lane] = xnew; }
    }
    __syncthreads();
  } else {
    if (tid < Ds) {
      s_x[tid] = r.states[(size_t)m * Ds + tid];
      s_pin[tid] = ms.enabled ? r.pol_in[(size_t)m * Ds + tid] : s_x[tid];
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------------- pre(t): policy, features, K* rows
  if (tid < pol.Dp) {
    s_il[tid] = exp(-pol.log_ls[tid]);
    s_z[tid] = policy_feature(pol, s_pin, t, tid);
  }
  __syncthreads();
  {
    const bool drop = dropout_active(pol, nz);
    const double keep_scale = drop ? 1.0 / (1.0 - nz.p_dropout) : 1.0;
    double a[MCP_MAX_DU];
#pragma unroll
    for (int k = 0; k < MCP_MAX_DU; k++) a[k] = 0.0;
    for (int b = tid; b < pol.nb; b += SS_THREADS) {
      const double* c = pol.centers + (size_t)b * pol.Dp;
      double d = 0.0;
#pragma unroll 8
      for (int j = 0; j < pol.Dp; j++) {
        double rr = (s_z[j] - c[j]) * s_il[j];
        d = fma(rr, rr, d);
      }
      double h = exp(-d);
      if (drop) h = keep_unit(nz, M, pol.nb, t, t, m, b) ? h * keep_scale : 0.0;
#pragma unroll
      for (int k = 0; k < MCP_MAX_DU; k++)
        if (k < pol.Du) a[k] = fma(pol.W[(size_t)k * pol.nb + b], h, a[k]);
    }
#pragma unroll
    for (int k = 0; k < MCP_MAX_DU; k++)
      if (k < pol.Du) {
        double vv = warp_sum(a[k]);
        if (lane == 0) s_part[warp][k] = vv;
      }
    __syncthreads();
    if (tid < pol.Du) {
      double vv = (s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]);
      if (pol.has_bias) vv += pol.bias[tid];
      if (pol.squash) vv = pol.u_max[tid] * tanh(vv / pol.u_max[tid]);
      r.inputs[((size_t)t * M + m) * Du + tid] = vv;
      s_u[tid] = vv;
    }
    __syncthreads();
  }
  if (t == H - 1) return;
  if (tid < DT) {
    const int j = tid;
    double f = 0.0;
    if (j < D) {
      if (mdl.use_trig) {
        if (j < mdl.n_na) f = s_x[mdl.na_idx[j]];
        else if (j < mdl.n_na + mdl.n_a) f = sin(s_x[mdl.a_idx[j - mdl.n_na]]);
        else if (j < mdl.n_na + 2 * mdl.n_a) f = cos(s_x[mdl.a_idx[j - mdl.n_na - mdl.n_a]]);
        else f = s_u[j - mdl.n_na - 2 * mdl.n_a];
      } else {
        f = (j < mdl.Ds) ? s_x[j] : s_u[j - mdl.Ds];
      }
      Xs[(size_t)m * D + j] = f;
    }
    s_feat[j] = f;
  }
  __syncthreads();
  {
    double x[DT];
#pragma unroll
    for (int j = 0; j < DT; j++) x[j] = s_feat[j];
    for (int e = 0; e < E; e++) {
      const McpGpDev& g = gps[e];
      const McpGpSpec& s = g.spec;
      double* krow = Ks + e * gp_stride + (size_t)m * ldk;
      for (int n = tid; n < ldk; n += SS_THREADS) {
        double kv = 0.0;
        if (n < g.N) {
          double y[DT];
          KFn<DT>::load(y, g.Xtr + (size_t)n * D, D);
          double d2 = 0.0;
#pragma unroll
          for (int j = 0; j < DT; j++) {
            const double tt = (x[j] * s.inv_ls[j]) - y[j] * s.inv_ls[j];
            d2 = fma(tt, tt, d2);
          }
          kv = s.has_se ? s.lambda * exp(-d2) : 0.0;
          if (NP >= 1) {
            double L1 = s.poly_w2[0][0][MCP_MAX_D];
#pragma unroll
            for (int j = 0; j < DT; j++) L1 = fma(s.poly_w2[0][0][j] * x[j], y[j], L1);
            kv += L1;
          }
          if (NP >= 2) {
            double La = s.poly_w2[1][0][MCP_MAX_D], Lb = s.poly_w2[1][1][MCP_MAX_D];
#pragma unroll
            for (int j = 0; j < DT; j++) {
              La = fma(s.poly_w2[1][0][j] * x[j], y[j], La);
              Lb = fma(s.poly_w2[1][1][j] * x[j], y[j], Lb);
            }
            kv = fma(La, Lb, kv);
          }
        }
        krow[n] = kv;
      }
    }
  }
}

// host: is this rollout one for the fused small path?
int launch_small_gemm(const McpGpDev* tab, int M, int nmax, int E, const double* Ks, double* V, int ldk, size_t gp_stride, bool pdl,
                      cudaStream_t st) {
  static bool cfg[MCP_MAX_DEVICES] = {};
  MCP_CUDA(ensure_dynamic_smem(cfg, small_gemm_kernel, (int)SG_SMEM_BYTES));
  MCP_CUDA(launch_chain(pdl, small_gemm_kernel, dim3(cdiv(nmax, 32), cdiv(M, 32), E), dim3(128), SG_SMEM_BYTES, st, tab, M, Ks, V, ldk, gp_stride));
  count_launch();
  return MCP_OK;
}

bool small_path_ok(const McpRollout* r) {
  const char* off = getenv("MCPILCO_NO_SMALL_PATH");  // tests use it to hold the per-step kernels to the same golden vectors
  if (off != nullptr && off[0] == '1') return false;
  const int Mg = r->M_global > 0 ? r->M_global : r->M;
  if (Mg > 2048 || r->model.D > 8 || r->model.E > MCP_MAX_E) return false;
  int np = -1;
  for (int e = 0; e < r->model.E; e++) {
    const McpGpSpec& s = r->gps[e].spec;
    if (!s.has_se || s.n_poly > 2 || r->gps[e].N > 2048) return false;
    for (int p = 0; p < s.n_poly; p++)
      if (s.poly_deg[p] != p + 1) return false;
    if (np >= 0 && s.n_poly != np) return false;  // one template instance for all outputs
    np = s.n_poly;
    if (s.D > 6 && s.n_poly == 2) return false;
    if (r->gps[e].ld_kinv % 2 != 0 || ((uintptr_t)r->gps[e].Kinv % 16) != 0) return false;
  }
  return true;
}

size_t small_path_doubles(int M, int E, int Nmax) {
  const size_t ldk = (size_t)(Nmax + 15) / 16 * 16;
  return 2 * (size_t)E * M * ldk + (sizeof(McpGpDev) * (size_t)E + 7) / 8 + 64;
}

// host: the whole forward loop of a small rollout.  `scratch` holds [table | K* (E x M x ldk) | V (E x M x ldk)].
int rollout_fwd_small(const McpRollout* r, double* Xs, double* nv, double* scratch, size_t scratch_doubles, cudaStream_t st) {
  const int M = r->M, H = r->H, E = r->model.E, D = r->model.D;
  int nmax = 1;
  for (int e = 0; e < E; e++) nmax = r->gps[e].N > nmax ? r->gps[e].N : nmax;
  MCP_CHECK_ARG(scratch_doubles >= small_path_doubles(M, E, nmax), "rollout (small path): workspace too small");
  const int ldk = (nmax + 15) / 16 * 16;
  const size_t gp_stride = (size_t)M * ldk;
  McpGpDev* tab = reinterpret_cast<McpGpDev*>(scratch);
  double* Ks = scratch + (sizeof(McpGpDev) * (size_t)E + 7) / 8 + 8;
  Ks = reinterpret_cast<double*>(align_up((size_t)Ks, 256));
  double* V = Ks + (size_t)E * gp_stride;
  MCP_CUDA(gpdev_upload(tab, r->gps, E, st));
  const int np = r->gps[0].spec.n_poly;
  const bool jac = r->need_grad != 0;
  const bool pdl = pdl_enabled();
  for (int t = 0; t < H; t++) {
#define MCP_SS(DT_, NP_)                                                                                                                  \
  do {                                                                                                                                    \
    if (jac) MCP_CUDA(launch_chain(pdl, small_step_kernel<DT_, NP_, true>, dim3(M), dim3(SS_THREADS), 0, st, *r, tab, t, Ks, V, ldk, gp_stride, Xs, nv));  \
    else MCP_CUDA(launch_chain(pdl, small_step_kernel<DT_, NP_, false>, dim3(M), dim3(SS_THREADS), 0, st, *r, tab, t, Ks, V, ldk, gp_stride, Xs, nv));     \
  } while (0)
    if (D <= 4) { if (np == 0) MCP_SS(4, 0); else if (np == 1) MCP_SS(4, 1); else MCP_SS(4, 2); }
    else if (D <= 6) { if (np == 0) MCP_SS(6, 0); else if (np == 1) MCP_SS(6, 1); else MCP_SS(6, 2); }
    else { if (np == 0) MCP_SS(8, 0); else MCP_SS(8, 1); }
#undef MCP_SS
    count_launch();
    if (t == H - 1) break;
    if (int err = launch_small_gemm(tab, M, nmax, E, Ks, V, ldk, gp_stride, pdl, st)) return err;
  }
  return MCP_OK;
}

}  // namespace mcp
