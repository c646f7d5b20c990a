// Warp-per-QP OSQP-equivalent ADMM on the stage-structured LPV-MPC QP (generic variant).
//
// One warp owns one QP.  Its workspace (scaled problem data, iterates, block factor) lives in shared
// memory when it fits and in a per-warp global (L2) slab otherwise.  The decision vector keeps the
// reference order z = [x_0..x_N, u_0..u_{N-1}] (PathFollowingLPVMPC.py:157-158); constraint rows are kept
// internally as [dynamics rows (n(N+1)); single-variable rows] and mapped back to the reference's OSQP
// row order on output.
//
// The linear system of each ADMM step is the OSQP KKT system with the constraint block eliminated:
//     (P + sigma I + A' diag(rho) A) x~ = sigma x - q + A'(rho z - y),     z~ = A x~
// which in stage order w_k = [x_k; u_k] is block tridiagonal with (n+d)x(n+d) blocks.  It is factorised once
// per rho update as a block LDL' with explicitly inverted pivots (T_k = S~_k^-1, K_k = S_{k,k-1} T_{k-1}), so
// that a solve is two sweeps of small dense mat-vecs:  v_k = b_k - K_k v_{k-1};  x_k = T_k v_k - K_{k+1}' x_{k+1}.
//
// Algorithm (scaling, rho vector, iteration, termination, infeasibility certificates, adaptive rho,
// polish) follows OSQP 0.6 as restated in oracle/osqp_ref.c; see SURVEY.md A.5.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "lpvmpc.h"
#include "lpv_model.cuh"

namespace lpv {

constexpr double kInfty = 1e30;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr int NU = 2;
constexpr unsigned kFull = 0xffffffffu;
// The polish system is solved in condensed form with 1/delta row weights, which loses about six digits per
// solve compared with upstream's LDL' of the full reduced KKT.  Iterative refinement converges to the same
// KKT point regardless, so we run a few more passes than polish_refine_iter (measured on the cfg-2 batch:
// +2 passes reproduce the oracle's polish decisions and solutions to 1e-10; see DESIGN.md "Polish";
// tools/debug_polish_passes.py on the full 4,096-QP batch: 5 and 6 passes give identical decisions and solutions, 4
// passes lose 13 accepted polishes).
constexpr int kPolishExtraRefine = 2;   // legacy kernels only (polish_refine_iter + 2 outer passes of the condensed refinement)
// Refinement passes on the REGULARISED polish system after every condensed K_reg solve (see polish() in lpv_h8t.cuh):
// the condensed solve is accurate to ~1e-6 relative, one pass brings a K_reg solve to ~1e-12, upstream's LDL' level.
constexpr int kPolishInner = 1;

// Offsets (in doubles) of the per-QP workspace.
struct Layout {
  int N, nx, nz, md, ms, m;
  int G, gI, sc, Pxx, Puu, Pud, q, D, Dinv, E, Einv, l, u, z, y, x, xp, xt, tn, pdx, dy, tm, K, T, So, type, total;
};

struct Params {
  Layout L;
  Model M;
  lpvmpc_settings S;
  lpvmpc_args a;
  int B;
  double *gws;  // global workspace slabs (one per resident warp) when not in shared memory
};

template <int KIND> struct Spec;

template <> struct Spec<LPVMPC_CONTROLLER> {
  static constexpr int NX = 6;
  // single-variable rows: 2N state rows, 4N input rows, then `delay` steering pins (PathFollowingLPVMPC.py:334-348,518-527)
  __device__ static __forceinline__ int srow_var(const Layout &L, int s) {
    const int N = L.N;
    if (s < 2 * N) return (s >> 1) * NX;
    if (s < 6 * N) { const int t = s - 2 * N; return L.nx + (t >> 2) * NU + ((t & 3) >> 1); }
    return L.nx + (s - 6 * N) * NU;
  }
  __device__ static __forceinline__ double srow_coef(const Layout &L, int s) {
    const int N = L.N;
    if (s < 2 * N) return (s & 1) ? 1.0 : -1.0;
    if (s < 6 * N) return ((s - 2 * N) & 1) ? -1.0 : 1.0;
    return 1.0;
  }
  __device__ static __forceinline__ int var_srows(const Layout &L, int delay, int j, int *out) {
    const int N = L.N;
    if (j < L.nx) {
      const int k = j / NX, r = j - k * NX;
      if (r == 0 && k < N) { out[0] = 2 * k; out[1] = 2 * k + 1; return 2; }
      return 0;
    }
    const int t = j - L.nx, k = t >> 1, c = t & 1;
    out[0] = 2 * N + 4 * k + 2 * c; out[1] = out[0] + 1;
    if (c == 0 && k < delay) { out[2] = 6 * N + k; return 3; }
    return 2;
  }
  __device__ static __forceinline__ int ref_row(const Layout &L, int i) {
    if (i < L.md) return 6 * L.N + i;
    const int s = i - L.md;
    return (s < 6 * L.N) ? s : (6 * L.N + L.nx + (s - 6 * L.N));
  }
};

template <> struct Spec<LPVMPC_PLANNER> {
  static constexpr int NX = 5;
  // box rows: identity on every variable (LPV_MPC_Planner.py:179-181)
  __device__ static __forceinline__ int srow_var(const Layout &, int s) { return s; }
  __device__ static __forceinline__ double srow_coef(const Layout &, int) { return 1.0; }
  __device__ static __forceinline__ int var_srows(const Layout &, int, int j, int *out) { out[0] = j; return 1; }
  __device__ static __forceinline__ int ref_row(const Layout &L, int i) { return (i < L.md) ? i : (L.nx + (i - L.md)); }
};

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(kFull, v, o); v = (w > v) ? w : v; }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_or(int v) { return __any_sync(kFull, v); }
__device__ __forceinline__ double absmax(double mx, double v) { const double a = fabs(v); return (a > mx) ? a : mx; }
// NaN or +-inf.  absmax() and the clamps drop NaN (every comparison with it is false), so problem data is screened once,
// when it is built: a non-finite entry makes the QP a LPVMPC_DATA_ERROR instead of a "solved" all-NaN answer.
__device__ __forceinline__ bool notfinite(double v) { return !(fabs(v) <= 1.7976931348623157e308); }
__device__ __forceinline__ double limit_scaling(double v) {
  v = v < kMinScaling ? 1.0 : v;
  v = v > kMaxScaling ? kMaxScaling : v;
  return v;
}
__device__ __forceinline__ double clampd(double v, double lo, double hi) {
  const double a = (v > lo) ? v : lo;  // c_max(v, l)
  return (a < hi) ? a : hi;            // c_min(., u)
}

// Row weights of the condensed system: ADMM -> rho_vec by constraint type; polish -> 1/delta on active rows.
struct AdmmW {
  const int8_t *type; double rho, rho_eq;
  __device__ __forceinline__ double operator()(int i) const {
    const int t = type[i] & 3;
    return t == 1 ? rho_eq : (t == 0 ? rho : kRhoMin);
  }
};
struct PolishW {
  const int8_t *type; double inv_delta;
  __device__ __forceinline__ double operator()(int i) const { return (type[i] & 12) ? inv_delta : 0.0; }
};

template <int KIND>
struct QP {
  using S = Spec<KIND>;
  static constexpr int NX = S::NX;
  static constexpr int NB = NX + NU;
  static_assert(NB <= 8, "stage block must fit the 8x4 lane tiling");

  const Layout &L;
  const Model &M;
  double *w;  // workspace base
  int lane;

  __device__ QP(const Layout &L_, const Model &M_, double *w_, int lane_) : L(L_), M(M_), w(w_), lane(lane_) {}

  __device__ __forceinline__ int vidx(int k, int c) const { return c < NX ? k * NX + c : L.nx + k * NU + (c - NX); }
  __device__ __forceinline__ int8_t *types() const { return reinterpret_cast<int8_t *>(w + L.type); }

  // ---------------------------------------------------------------- mat-vecs on the structured A, P
  // out_i = (A v)_i for internal row i
  template <class V> __device__ __forceinline__ double rowA(int i, V v) const {
    const double *G = w + L.G, *gI = w + L.gI, *sc = w + L.sc;
    if (i < L.md) {
      const int k = i / NX, r = i - k * NX;
      double acc = gI[i] * v(i);  // x_k[r] has reference index k*NX + r == i
      if (k > 0) {
        const double *g = G + (size_t)((k - 1) * NX + r) * NB;
#pragma unroll
        for (int c = 0; c < NB; ++c) acc = fma(g[c], v(vidx(k - 1, c)), acc);
      }
      return acc;
    }
    const int s = i - L.md;
    return sc[s] * v(S::srow_var(L, s));
  }
  // out_j = (A' t)_j for variable j
  template <class T> __device__ __forceinline__ double colA(int j, T t) const {
    const double *G = w + L.G, *gI = w + L.gI, *sc = w + L.sc;
    double acc;
    if (j < L.nx) {
      const int k = j / NX, r = j - k * NX;
      acc = gI[j] * t(j);
      if (k < L.N) {
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) acc = fma(G[(size_t)(k * NX + rr) * NB + r], t((k + 1) * NX + rr), acc);
      }
    } else {
      const int tt = j - L.nx, k = tt >> 1, c = tt & 1;
      acc = 0.0;
#pragma unroll
      for (int rr = 0; rr < NX; ++rr) acc = fma(G[(size_t)(k * NX + rr) * NB + NX + c], t((k + 1) * NX + rr), acc);
    }
    int rows[3];
    const int cnt = S::var_srows(L, M.delay, j, rows);
    for (int q = 0; q < cnt; ++q) acc = fma(sc[rows[q]], t(L.md + rows[q]), acc);
    return acc;
  }
  // out_j = (P v)_j
  template <class V> __device__ __forceinline__ double rowP(int j, V v) const {
    const double *Pxx = w + L.Pxx, *Puu = w + L.Puu, *Pud = w + L.Pud;
    double acc = 0.0;
    if (j < L.nx) {
      const int k = j / NX, r = j - k * NX;
#pragma unroll
      for (int c = 0; c < NX; ++c) acc = fma(Pxx[(size_t)(k * NX + r) * NX + c], v(k * NX + c), acc);
    } else {
      const int tt = j - L.nx, k = tt >> 1, c = tt & 1;
#pragma unroll
      for (int cc = 0; cc < NU; ++cc) acc = fma(Puu[(k * NU + c) * NU + cc], v(L.nx + k * NU + cc), acc);
      if (k > 0) acc = fma(Pud[(k - 1) * NU + c], v(j - NU), acc);
      if (k < L.N - 1) acc = fma(Pud[k * NU + c], v(j + NU), acc);
    }
    return acc;
  }

  // ---------------------------------------------------------------- unscaled problem data
  // Cost blocks (PathFollowingLPVMPC.py:397-464, LPV_MPC_Planner.py:145-169) and bounds
  // (PathFollowingLPVMPC.py:334-348, LPV_MPC_Planner.py:173-181); A_k, B_k already sit (negated) in G.
  __device__ void build(const Params &p, int b, const double *x0) {
    const int N = L.N;
    double *Pxx = w + L.Pxx, *Puu = w + L.Puu, *Pud = w + L.Pud, *q = w + L.q, *gI = w + L.gI, *sc = w + L.sc;
    double *lo = w + L.l, *up = w + L.u;
    for (int e = lane; e < (N + 1) * NX * NX; e += 32) Pxx[e] = 2 * M.Q[e % (NX * NX)];
    for (int e = lane; e < N * NU * NU; e += 32) {
      const int k = e / (NU * NU), r = (e >> 1) & 1, c = e & 1;
      double v = M.R[r * NU + c] + (r == c ? 2 * M.dR[r] : 0.0);
      if (k == N - 1 && r == c) v = v - M.dR[r];
      Puu[e] = 2 * v;
    }
    for (int e = lane; e < (N - 1) * NU; e += 32) Pud[e] = 2 * (-M.dR[e & 1]);
    const double *uold = p.a.u_old ? p.a.u_old + (size_t)b * NU : nullptr;
    for (int j = lane; j < L.nz; j += 32) {
      double v;
      if (j < L.nx) {
        const int k = j / NX, r = j - k * NX;
        if (KIND == LPVMPC_CONTROLLER) v = -2 * (p.a.vel_ref[(size_t)b * (N + 1) + k] * M.Q[r]);
        else v = M.L_cf[r];
      } else {
        const int tt = j - L.nx;
        if (tt < NU) v = -2 * ((uold ? uold[tt] : 0.0) * M.dR[tt]);
        else v = (KIND == LPVMPC_CONTROLLER) ? -2 * 0.0 : 0.0;
      }
      q[j] = v;
    }
    for (int i = lane; i < L.md; i += 32) {
      gI[i] = 1.0;
      double v;
      if (i < NX) v = x0[i] + 0.0;
      else v = 0.0 + (p.a.C ? p.a.C[(size_t)b * N * NX + (i - NX)] : 0.0);
      lo[i] = v; up[i] = v;
    }
    for (int s = lane; s < L.ms; s += 32) {
      sc[s] = S::srow_coef(L, s);
      double l_, u_;
      if (KIND == LPVMPC_CONTROLLER) {
        if (s < 2 * N) { l_ = -kInfty; u_ = (s & 1) ? M.max_vel : -0.01; }
        else if (s < 6 * N) { const int t = (s - 2 * N) & 3; l_ = -kInfty; u_ = (t < 2) ? 0.249 : (t == 2 ? 4.0 : 1.0); }
        else { l_ = u_ = p.a.old_steering[(size_t)b * M.delay + (s - 6 * N)]; }
      } else {
        if (s < L.nx) {
          const int k = s / NX, r = s - k * NX;
          const double mey = p.a.max_ey[b];
          const double xmin[5] = {M.min_vel, -1.0, -2.0, -mey, -0.8};
          const double xmax[5] = {M.max_vel, 1.0, 2.0, mey, 0.8};
          l_ = xmin[r]; u_ = xmax[r];
          if (r == 3 && p.a.ey_lo) l_ = p.a.ey_lo[(size_t)b * (N + 1) + k];
          if (r == 3 && p.a.ey_hi) u_ = p.a.ey_hi[(size_t)b * (N + 1) + k];
        } else {
          const int c = (s - L.nx) & 1;
          l_ = c ? -0.7 : -0.249; u_ = c ? 2.0 : 0.249;
        }
      }
      // python wrapper: l = max(l, -OSQP_INFTY), u = min(u, OSQP_INFTY)
      lo[L.md + s] = (l_ > -kInfty || l_ != l_) ? l_ : -kInfty;   // a NaN bound stays NaN: bounds_invalid() reports it
      up[L.md + s] = (u_ < kInfty || u_ != u_) ? u_ : kInfty;
    }
    __syncwarp();
  }

  // l > u anywhere => upstream osqp.setup() rejects the problem
  __device__ bool bounds_invalid() const {
    const double *lo = w + L.l, *up = w + L.u, *q = w + L.q, *G = w + L.G;
    int bad = 0;
    for (int i = lane; i < L.m; i += 32) bad |= !(lo[i] <= up[i]);   // l > u, or a NaN bound / x0 / C
    for (int j = lane; j < L.nz; j += 32) bad |= notfinite(q[j]);
    for (int e = lane; e < L.N * NX * NB; e += 32) bad |= notfinite(G[e]);
    return warp_or(bad);
  }

  // ---------------------------------------------------------------- Ruiz equilibration (scale_data)
  __device__ double colnormP(int j) const {
    const double *Pxx = w + L.Pxx, *Puu = w + L.Puu, *Pud = w + L.Pud;
    double mx = 0.0;
    if (j < L.nx) {
      const int k = j / NX, r = j - k * NX;
#pragma unroll
      for (int c = 0; c < NX; ++c) mx = absmax(mx, Pxx[(size_t)(k * NX + r) * NX + c]);
    } else {
      const int tt = j - L.nx, k = tt >> 1, c = tt & 1;
#pragma unroll
      for (int cc = 0; cc < NU; ++cc) mx = absmax(mx, Puu[(k * NU + c) * NU + cc]);
      if (k > 0) mx = absmax(mx, Pud[(k - 1) * NU + c]);
      if (k < L.N - 1) mx = absmax(mx, Pud[k * NU + c]);
    }
    return mx;
  }
  __device__ double colnormA(int j) const {
    const double *G = w + L.G, *gI = w + L.gI, *sc = w + L.sc;
    double mx = 0.0;
    if (j < L.nx) {
      const int k = j / NX, r = j - k * NX;
      mx = absmax(mx, gI[j]);
      if (k < L.N)
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) mx = absmax(mx, G[(size_t)(k * NX + rr) * NB + r]);
    } else {
      const int tt = j - L.nx, k = tt >> 1, c = tt & 1;
#pragma unroll
      for (int rr = 0; rr < NX; ++rr) mx = absmax(mx, G[(size_t)(k * NX + rr) * NB + NX + c]);
    }
    int rows[3];
    const int cnt = S::var_srows(L, M.delay, j, rows);
    for (int q = 0; q < cnt; ++q) mx = absmax(mx, sc[rows[q]]);
    return mx;
  }
  __device__ double rownormA(int i) const {
    const double *G = w + L.G, *gI = w + L.gI, *sc = w + L.sc;
    if (i < L.md) {
      const int k = i / NX, r = i - k * NX;
      double mx = fabs(gI[i]);
      if (k > 0)
#pragma unroll
        for (int c = 0; c < NB; ++c) mx = absmax(mx, G[(size_t)((k - 1) * NX + r) * NB + c]);
      return mx;
    }
    return fabs(sc[i - L.md]);
  }

  __device__ double scale(int passes) {
    const int N = L.N;
    double *G = w + L.G, *gI = w + L.gI, *sc = w + L.sc, *Pxx = w + L.Pxx, *Puu = w + L.Puu, *Pud = w + L.Pud;
    double *q = w + L.q, *D = w + L.D, *Dinv = w + L.Dinv, *E = w + L.E, *Einv = w + L.Einv;
    double *Dt = w + L.tn, *Et = w + L.tm, *lo = w + L.l, *up = w + L.u;
    for (int j = lane; j < L.nz; j += 32) D[j] = 1.0;
    for (int i = lane; i < L.m; i += 32) E[i] = 1.0;
    double c = 1.0;
    __syncwarp();
    for (int it = 0; it < passes; ++it) {
      for (int j = lane; j < L.nz; j += 32) {
        const double a = colnormP(j), bq = colnormA(j);
        Dt[j] = 1.0 / sqrt(limit_scaling(a > bq ? a : bq));
      }
      for (int i = lane; i < L.m; i += 32) Et[i] = 1.0 / sqrt(limit_scaling(rownormA(i)));
      __syncwarp();
      // P <- D P D (rows, then columns, on the upper-triangle entry; the mirror gets the same value)
      for (int e = lane; e < (N + 1) * NX * NX; e += 32) {
        const int k = e / (NX * NX), rc = e - k * NX * NX, r = rc / NX, cc = rc - r * NX;
        const int i0 = k * NX + (r < cc ? r : cc), j0 = k * NX + (r < cc ? cc : r);
        Pxx[e] = (Pxx[e] * Dt[i0]) * Dt[j0];
      }
      for (int e = lane; e < N * NU * NU; e += 32) {
        const int k = e >> 2, r = (e >> 1) & 1, cc = e & 1;
        const int i0 = L.nx + k * NU + (r < cc ? r : cc), j0 = L.nx + k * NU + (r < cc ? cc : r);
        Puu[e] = (Puu[e] * Dt[i0]) * Dt[j0];
      }
      for (int e = lane; e < (N - 1) * NU; e += 32) Pud[e] = (Pud[e] * Dt[L.nx + e]) * Dt[L.nx + e + NU];
      // A <- E A D
      for (int e = lane; e < N * NX * NB; e += 32) {
        const int k = e / (NX * NB), rc = e - k * NX * NB, r = rc / NB, cc = rc - r * NB;
        G[e] = (G[e] * Et[(k + 1) * NX + r]) * Dt[vidx(k, cc)];
      }
      for (int i = lane; i < L.md; i += 32) gI[i] = (gI[i] * Et[i]) * Dt[i];
      for (int s = lane; s < L.ms; s += 32) sc[s] = (sc[s] * Et[L.md + s]) * Dt[S::srow_var(L, s)];
      for (int j = lane; j < L.nz; j += 32) { q[j] = Dt[j] * q[j]; D[j] = D[j] * Dt[j]; }
      for (int i = lane; i < L.m; i += 32) E[i] = E[i] * Et[i];
      __syncwarp();
      // cost scaling
      double qn = 0.0;
      for (int j = lane; j < L.nz; j += 32) { Dt[j] = colnormP(j); qn = absmax(qn, q[j]); }
      qn = warp_max(qn);
      __syncwarp();
      double ct = 0.0;
      if (lane == 0) {  // vec_mean in index order, as upstream
        for (int j = 0; j < L.nz; ++j) ct += Dt[j];
        ct = ct / L.nz;
      }
      ct = __shfl_sync(kFull, ct, 0);
      qn = limit_scaling(qn);
      ct = ct > qn ? ct : qn;
      ct = limit_scaling(ct);
      ct = 1.0 / ct;
      for (int e = lane; e < (N + 1) * NX * NX; e += 32) Pxx[e] *= ct;
      for (int e = lane; e < N * NU * NU; e += 32) Puu[e] *= ct;
      for (int e = lane; e < (N - 1) * NU; e += 32) Pud[e] *= ct;
      for (int j = lane; j < L.nz; j += 32) q[j] *= ct;
      c *= ct;
      __syncwarp();
    }
    for (int j = lane; j < L.nz; j += 32) Dinv[j] = 1.0 / D[j];
    for (int i = lane; i < L.m; i += 32) { Einv[i] = 1.0 / E[i]; lo[i] = E[i] * lo[i]; up[i] = E[i] * up[i]; }
    __syncwarp();
    return c;
  }

  // constraint types (set_rho_vec): 2 = loose, 1 = equality, 0 = inequality
  __device__ void classify() {
    int8_t *ty = types();
    const double *lo = w + L.l, *up = w + L.u;
    for (int i = lane; i < L.m; i += 32) {
      int t;
      if ((lo[i] < -kInfty * kMinScaling) && (up[i] > kInfty * kMinScaling)) t = 2;
      else if (up[i] - lo[i] < kRhoTol) t = 1;
      else t = 0;
      ty[i] = (int8_t)t;
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- block factorisation
  template <class W> __device__ void factor(double sigma, W wt) {
    const int N = L.N;
    const double *G = w + L.G, *gI = w + L.gI, *sc = w + L.sc, *Pxx = w + L.Pxx, *Puu = w + L.Puu, *Pud = w + L.Pud;
    double *K = w + L.K, *T = w + L.T, *So = w + L.So;
    for (int k = 0; k <= N; ++k) {
      const int nbk = (k < N) ? NB : NX;
      double *Tk = T + (size_t)k * NB * NB, *Kk = K + (size_t)k * NB * NB;
      const double *Tp = Tk - NB * NB;
      for (int e = lane; e < NB * NB; e += 32) {
        const int r = e / NB, c = e - r * NB;
        double s = 0.0, so = 0.0;
        if (r < nbk && c < nbk) {
          if (r < NX && c < NX) s = Pxx[(size_t)(k * NX + r) * NX + c];
          else if (r >= NX && c >= NX) s = Puu[(k * NU + (r - NX)) * NU + (c - NX)];
          if (r == c) {
            s += sigma;
            int rows[3];
            const int cnt = S::var_srows(L, M.delay, vidx(k, r), rows);
            for (int q = 0; q < cnt; ++q) { const double a = sc[rows[q]]; s = fma(wt(L.md + rows[q]) * a, a, s); }
            if (r < NX) { const int i = k * NX + r; s = fma(wt(i) * gI[i], gI[i], s); }
          }
          if (k < N) {
#pragma unroll
            for (int rr = 0; rr < NX; ++rr) {
              const double *g = G + (size_t)(k * NX + rr) * NB;
              s = fma(wt((k + 1) * NX + rr) * g[r], g[c], s);
            }
          }
        }
        if (k > 0 && r < nbk) {
          if (r < NX) { const int i = k * NX + r; so = (wt(i) * gI[i]) * G[(size_t)((k - 1) * NX + r) * NB + c]; }
          else if (c == r) so = Pud[(k - 1) * NU + (r - NX)];
        }
        Tk[e] = s;
        So[e] = so;
      }
      __syncwarp();
      if (k > 0) {
        for (int e = lane; e < NB * NB; e += 32) {
          const int r = e / NB, c = e - r * NB;
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < NB; ++j) acc = fma(So[r * NB + j], Tp[j * NB + c], acc);
          Kk[e] = acc;
        }
        __syncwarp();
        for (int e = lane; e < NB * NB; e += 32) {
          const int r = e / NB, c = e - r * NB;
          if (r < nbk && c < nbk) {
            double acc = Tk[e];
#pragma unroll
            for (int j = 0; j < NB; ++j) acc = fma(-Kk[r * NB + j], So[c * NB + j], acc);
            Tk[e] = acc;
          }
        }
        __syncwarp();
      }
      // in-place Gauss-Jordan inverse of the SPD pivot block (no pivoting)
      for (int p = 0; p < nbk; ++p) {
        const double piv = 1.0 / Tk[p * NB + p];
        double nv[2];
        int q = 0;
        for (int e = lane; e < NB * NB; e += 32, ++q) {
          const int r = e / NB, c = e - r * NB;
          double v = Tk[e];
          if (r < nbk && c < nbk) {
            if (r == p && c == p) v = piv;
            else if (r == p) v = v * piv;
            else if (c == p) v = -v * piv;
            else v = fma(-(Tk[r * NB + p] * piv), Tk[p * NB + c], v);
          }
          nv[q] = v;
        }
        __syncwarp();
        q = 0;
        for (int e = lane; e < NB * NB; e += 32, ++q) Tk[e] = nv[q];
        __syncwarp();
      }
    }
  }

  // ---------------------------------------------------------------- block solve, in place on v (uses tn)
  __device__ void solve(double *v) {
    const int N = L.N;
    const double *K = w + L.K, *T = w + L.T;
    double *wv = w + L.tn;
    const int r = lane >> 2, g = lane & 3;
    // forward: v_k -= K_k v_{k-1}
    for (int k = 1; k <= N; ++k) {
      const int nbk = (k < N) ? NB : NX;
      const double *Kk = K + (size_t)k * NB * NB + r * NB + 2 * g;
      double acc = 0.0;
      if (r < nbk) {
        if (2 * g < NB) acc = Kk[0] * v[vidx(k - 1, 2 * g)];
        if (2 * g + 1 < NB) acc = fma(Kk[1], v[vidx(k - 1, 2 * g + 1)], acc);
      }
      acc += __shfl_xor_sync(kFull, acc, 1);
      acc += __shfl_xor_sync(kFull, acc, 2);
      if (g == 0 && r < nbk) v[vidx(k, r)] -= acc;
      __syncwarp();
    }
    // pivots: w_k = T_k v_k (all stages in parallel)
    for (int e = lane; e < (N + 1) * NB; e += 32) {
      const int k = e / NB, rr = e - k * NB;
      const int nbk = (k < N) ? NB : NX;
      if (rr < nbk) {
        const double *Tk = T + (size_t)k * NB * NB + rr * NB;
        double acc = 0.0;
        for (int c = 0; c < nbk; ++c) acc = fma(Tk[c], v[vidx(k, c)], acc);
        wv[vidx(k, rr)] = acc;
      }
    }
    __syncwarp();
    // backward: x_k = w_k - K_{k+1}' x_{k+1}
    for (int e = lane; e < NX; e += 32) v[vidx(N, e)] = wv[vidx(N, e)];
    __syncwarp();
    for (int k = N - 1; k >= 0; --k) {
      const int nbn = (k + 1 < N) ? NB : NX;
      const double *Kn = K + (size_t)(k + 1) * NB * NB;
      const int c = r;  // variable within stage k
      double acc = 0.0;
      if (c < NB) {
        if (2 * g < nbn) acc = Kn[(2 * g) * NB + c] * v[vidx(k + 1, 2 * g)];
        if (2 * g + 1 < nbn) acc = fma(Kn[(2 * g + 1) * NB + c], v[vidx(k + 1, 2 * g + 1)], acc);
      }
      acc += __shfl_xor_sync(kFull, acc, 1);
      acc += __shfl_xor_sync(kFull, acc, 2);
      if (g == 0 && c < NB) v[vidx(k, c)] = wv[vidx(k, c)] - acc;
      __syncwarp();
    }
  }
};

// Result of one QP
struct Outcome {
  int status, iter, rho_updates, polish_status;
  double obj, pri_res, dua_res;
};

// ---------------------------------------------------------------- the whole OSQP run for one QP
template <int KIND>
__device__ void admm_run(QP<KIND> &qp, const lpvmpc_settings &S, double c, Outcome &out, const Params &p, int b) {
  using Q = QP<KIND>;
  const Layout &L = qp.L;
  const int lane = qp.lane;
  double *w = qp.w;
  double *x = w + L.x, *xp = w + L.xp, *xt = w + L.xt, *z = w + L.z, *y = w + L.y, *dy = w + L.dy, *tm = w + L.tm;
  const double *q = w + L.q, *lo = w + L.l, *up = w + L.u, *D = w + L.D, *Dinv = w + L.Dinv, *E = w + L.E, *Einv = w + L.Einv;
  int8_t *ty = qp.types();
  const double cinv = 1.0 / c;
  const double sigma = S.sigma, alpha = S.alpha;
  const bool unscale = S.scaling && !S.scaled_termination;

  double rho = fmin(fmax(S.rho, kRhoMin), kRhoMax);
  double rho_eq = kRhoEqOverIneq * rho;
  double rinv = 1.0 / rho, rinv_eq = 1.0 / rho_eq;
  const double rinv_loose = 1.0 / kRhoMin;

  for (int j = lane; j < L.nz; j += 32) { x[j] = 0.0; xp[j] = 0.0; }
  for (int i = lane; i < L.m; i += 32) { z[i] = 0.0; y[i] = 0.0; dy[i] = 0.0; }
  __syncwarp();
  qp.factor(sigma, AdmmW{ty, rho, rho_eq});

  int status = LPVMPC_UNSOLVED, iter_done = 0, rho_updates = 0;
  double pri_res = 0.0, dua_res = 0.0;
  // norms kept from the last update_info (scaled and unscaled)
  double n_rp = 0, n_z = 0, n_Ax = 0, n_rd = 0, n_q = 0, n_Aty = 0, n_Px = 0;      // scaled-space inf norms
  double u_z = 0, u_Ax = 0, u_q = 0, u_Aty = 0, u_Px = 0;                           // unscaled (termination)
  int adapt_interval = S.adaptive_rho_interval;
  if (S.adaptive_rho && !adapt_interval)
    adapt_interval = S.check_termination ? 4 * S.check_termination : 100;

  auto update_info = [&]() {
    double a_rp = 0, a_z = 0, a_Ax = 0, b_rp = 0, b_z = 0, b_Ax = 0;
    for (int i = lane; i < L.m; i += 32) {
      const double Ax = qp.rowA(i, [&](int j) { return x[j]; });
      const double r = Ax - z[i];
      a_rp = absmax(a_rp, r); a_z = absmax(a_z, z[i]); a_Ax = absmax(a_Ax, Ax);
      b_rp = absmax(b_rp, Einv[i] * r); b_z = absmax(b_z, Einv[i] * z[i]); b_Ax = absmax(b_Ax, Einv[i] * Ax);
    }
    double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, b_rd = 0, b_q = 0, b_Aty = 0, b_Px = 0;
    for (int j = lane; j < L.nz; j += 32) {
      const double Px = qp.rowP(j, [&](int jj) { return x[jj]; });
      const double Aty = qp.colA(j, [&](int i) { return y[i]; });
      const double r = (q[j] + Px) + Aty;
      a_rd = absmax(a_rd, r); a_q = absmax(a_q, q[j]); a_Aty = absmax(a_Aty, Aty); a_Px = absmax(a_Px, Px);
      b_rd = absmax(b_rd, Dinv[j] * r); b_q = absmax(b_q, Dinv[j] * q[j]);
      b_Aty = absmax(b_Aty, Dinv[j] * Aty); b_Px = absmax(b_Px, Dinv[j] * Px);
    }
    n_rp = warp_max(a_rp); n_z = warp_max(a_z); n_Ax = warp_max(a_Ax);
    n_rd = warp_max(a_rd); n_q = warp_max(a_q); n_Aty = warp_max(a_Aty); n_Px = warp_max(a_Px);
    if (unscale) {
      pri_res = warp_max(b_rp); u_z = warp_max(b_z); u_Ax = warp_max(b_Ax);
      dua_res = cinv * warp_max(b_rd); u_q = warp_max(b_q); u_Aty = warp_max(b_Aty); u_Px = warp_max(b_Px);
    } else {
      pri_res = n_rp; u_z = n_z; u_Ax = n_Ax; dua_res = n_rd; u_q = n_q; u_Aty = n_Aty; u_Px = n_Px;
    }
    if (L.m == 0) pri_res = 0.0;
  };

  auto primal_infeasible = [&](double eps) -> bool {
    double nrm = 0.0;
    for (int i = lane; i < L.m; i += 32) {
      double d = dy[i];
      if (up[i] > kInfty * kMinScaling) {
        if (lo[i] < -kInfty * kMinScaling) d = 0.0;
        else d = (d < 0.0) ? d : 0.0;
      } else if (lo[i] < -kInfty * kMinScaling) d = (d > 0.0) ? d : 0.0;
      dy[i] = d;
      nrm = absmax(nrm, unscale ? E[i] * d : d);
    }
    nrm = warp_max(nrm);
    __syncwarp();
    if (nrm > eps) {
      double lhs = 0.0;
      for (int i = lane; i < L.m; i += 32) {
        const double d = dy[i];
        lhs += up[i] * ((d > 0) ? d : 0) + lo[i] * ((d < 0) ? d : 0);
      }
      lhs = warp_sum(lhs);
      if (lhs < -eps * nrm) {
        double mx = 0.0;
        for (int j = lane; j < L.nz; j += 32) {
          double a = qp.colA(j, [&](int i) { return dy[i]; });
          if (unscale) a = Dinv[j] * a;
          mx = absmax(mx, a);
        }
        mx = warp_max(mx);
        return mx < eps * nrm;
      }
    }
    return false;
  };

  auto dual_infeasible = [&](double eps) -> bool {
    double nrm = 0.0, qdx = 0.0;
    for (int j = lane; j < L.nz; j += 32) {
      const double dx = x[j] - xp[j];
      nrm = absmax(nrm, unscale ? D[j] * dx : dx);
      qdx += q[j] * dx;
    }
    nrm = warp_max(nrm);
    qdx = warp_sum(qdx);
    const double cs = unscale ? c : 1.0;
    if (nrm > eps) {
      if (qdx < -cs * eps * nrm) {
        double mx = 0.0;
        for (int j = lane; j < L.nz; j += 32) {
          double a = qp.rowP(j, [&](int jj) { return x[jj] - xp[jj]; });
          if (unscale) a = Dinv[j] * a;
          mx = absmax(mx, a);
        }
        mx = warp_max(mx);
        if (mx < cs * eps * nrm) {
          int viol = 0;
          for (int i = lane; i < L.m; i += 32) {
            double a = qp.rowA(i, [&](int j) { return x[j] - xp[j]; });
            if (unscale) a = Einv[i] * a;
            if (((up[i] < kInfty * kMinScaling) && (a > eps * nrm)) || ((lo[i] > -kInfty * kMinScaling) && (a < -eps * nrm)))
              viol = 1;
          }
          return !warp_or(viol);
        }
      }
    }
    return false;
  };

  // returns 1 when a termination status was set
  auto check_termination = [&](int approximate) -> int {
    double eps_abs = S.eps_abs, eps_rel = S.eps_rel, eps_pi = S.eps_prim_inf, eps_di = S.eps_dual_inf;
    if (!(pri_res <= kInfty) || !(dua_res <= kInfty)) { status = LPVMPC_NON_CVX; out.obj = nan(""); return 1; }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; eps_di *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (L.m == 0) prim_ok = true;
    else {
      const double eps_prim = eps_abs + eps_rel * (u_z > u_Ax ? u_z : u_Ax);
      if (pri_res < eps_prim) prim_ok = true;
      else prim_inf = primal_infeasible(eps_pi);
    }
    double mr = u_q; mr = (u_Aty > mr) ? u_Aty : mr; mr = (u_Px > mr) ? u_Px : mr;
    if (unscale) mr *= cinv;
    const double eps_dual = eps_abs + eps_rel * mr;
    if (dua_res < eps_dual) dual_ok = true;
    else dual_inf = dual_infeasible(eps_di);
    if (prim_ok && dual_ok) { status = approximate ? LPVMPC_SOLVED_INACCURATE : LPVMPC_SOLVED; return 1; }
    if (prim_inf) { status = approximate ? LPVMPC_PRIMAL_INFEASIBLE_INACCURATE : LPVMPC_PRIMAL_INFEASIBLE; out.obj = kInfty; return 1; }
    if (dual_inf) { status = approximate ? LPVMPC_DUAL_INFEASIBLE_INACCURATE : LPVMPC_DUAL_INFEASIBLE; out.obj = -kInfty; return 1; }
    return 0;
  };

  bool checked_last = false;
  int iter;
  for (iter = 1; iter <= S.max_iter; ++iter) {
    const bool can_check = S.check_termination && (iter % S.check_termination == 0);
    const bool can_adapt = S.adaptive_rho && adapt_interval && (iter % adapt_interval == 0);
    // right-hand side of the condensed system
    for (int j = lane; j < L.nz; j += 32) {
      const double at = qp.colA(j, [&](int i) {
        const int t = ty[i] & 3;
        const double r = t == 1 ? rho_eq : (t == 0 ? rho : kRhoMin);
        return r * z[i] - y[i];
      });
      xt[j] = (sigma * x[j] - q[j]) + at;
    }
    __syncwarp();
    qp.solve(xt);
    // z, y updates (z~ = A x~ row by row), then x
    for (int i = lane; i < L.m; i += 32) {
      const double zt = qp.rowA(i, [&](int j) { return xt[j]; });
      const int t = ty[i] & 3;
      const double r = t == 1 ? rho_eq : (t == 0 ? rho : kRhoMin);
      const double ri = t == 1 ? rinv_eq : (t == 0 ? rinv : rinv_loose);
      const double zr = alpha * zt + (1.0 - alpha) * z[i];
      const double zn = clampd(zr + ri * y[i], lo[i], up[i]);
      const double d = r * (zr - zn);
      y[i] += d;
      z[i] = zn;
      if (can_check || iter == S.max_iter) dy[i] = d;
    }
    for (int j = lane; j < L.nz; j += 32) {
      const double xo = x[j];
      xp[j] = xo;
      x[j] = alpha * xt[j] + (1.0 - alpha) * xo;
    }
    __syncwarp();
    checked_last = can_check;
    if (can_check) {
      update_info();
      iter_done = iter;
      if (check_termination(0)) break;
    }
    if (can_adapt) {
      if (!can_check) { update_info(); iter_done = iter; }
      const double pr = n_rp / ((n_z > n_Ax ? n_z : n_Ax) + 1e-10);
      double dn = n_q; dn = (n_Aty > dn) ? n_Aty : dn; dn = (n_Px > dn) ? n_Px : dn;
      const double dr = n_rd / (dn + 1e-10);
      double rho_new = rho * sqrt(pr / (dr + 1e-10));
      rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
      if ((rho_new > rho * S.adaptive_rho_tolerance) || (rho_new < rho / S.adaptive_rho_tolerance)) {
        rho = rho_new; rho_eq = kRhoEqOverIneq * rho; rinv = 1.0 / rho; rinv_eq = 1.0 / rho_eq;
        ++rho_updates;
        qp.factor(sigma, AdmmW{ty, rho, rho_eq});
      }
    }
  }
  if (!checked_last) {
    update_info();
    iter_done = iter - 1;
    check_termination(0);
  }
  if (status == LPVMPC_UNSOLVED) {
    if (!check_termination(1)) status = LPVMPC_MAX_ITER_REACHED;
  }
  const bool has_sol = !(status == LPVMPC_PRIMAL_INFEASIBLE || status == LPVMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                         status == LPVMPC_DUAL_INFEASIBLE || status == LPVMPC_DUAL_INFEASIBLE_INACCURATE ||
                         status == LPVMPC_NON_CVX);
  auto objective = [&](const double *xv) -> double {
    double acc = 0.0;
    for (int j = lane; j < L.nz; j += 32) {
      const double Px = qp.rowP(j, [&](int jj) { return xv[jj]; });
      acc += (0.5 * Px + q[j]) * xv[j];
    }
    return warp_sum(acc) * (S.scaling ? cinv : 1.0);
  };
  if (has_sol) out.obj = objective(x);

  // scaled iterates before polish (parity instrumentation)
  if (p.a.xs) for (int j = lane; j < L.nz; j += 32) p.a.xs[(size_t)b * L.nz + j] = x[j];
  if (p.a.zs) for (int i = lane; i < L.m; i += 32) p.a.zs[(size_t)b * L.m + Q::S::ref_row(L, i)] = z[i];
  if (p.a.ys) for (int i = lane; i < L.m; i += 32) p.a.ys[(size_t)b * L.m + Q::S::ref_row(L, i)] = y[i];

  // ---------------------------------------------------------------- polish
  int polish_status = 0;
  for (int i = lane; i < L.m; i += 32) ty[i] &= 3;
  if (S.polish && status == LPVMPC_SOLVED) {
    // active-set guess (form_Ared)
    for (int i = lane; i < L.m; i += 32) {
      int t = ty[i] & 3;
      if (z[i] - lo[i] < -y[i]) t |= 4;
      if (up[i] - z[i] < y[i]) t |= 8;
      ty[i] = (int8_t)t;
    }
    __syncwarp();
    const double delta = S.delta, idel = 1.0 / S.delta;
    qp.factor(delta, PolishW{ty, idel});
    double *px = xp, *py = dy, *pz = tm;  // polished x, y (reduced rows, 0 elsewhere), z
    // active-row targets: l for lower-active (priority), u for upper-active
    auto bred = [&](int i) { return (ty[i] & 4) ? lo[i] : up[i]; };
    // Upstream's refinement, pass for pass: s_0 = K_reg^-1 rhs, s_{j+1} = s_j + K_reg^-1 (rhs - K s_j); every K_reg solve
    // (condensed form, accurate to ~1e-6) is refined kPolishInner times on the regularised system.  See polish() in lpv_h8t.cuh.
    double *dxv = w + L.pdx;   // x part of the increment of the running K_reg solve
    auto inner = [&]() {
      for (int ii = 0; ii < kPolishInner; ++ii) {
        for (int j = lane; j < L.nz; j += 32) {
          const double Px = qp.rowP(j, [&](int jj) { return px[jj]; });
          const double Aty = qp.colA(j, [&](int i) { return py[i]; });
          xt[j] = ((-q[j] - Px) - Aty) - delta * dxv[j];
        }
        __syncwarp();
        qp.solve(xt);
        for (int i = lane; i < L.m; i += 32)
          if (ty[i] & 12) py[i] += idel * qp.rowA(i, [&](int j) { return xt[j]; });
        for (int j = lane; j < L.nz; j += 32) { px[j] += xt[j]; dxv[j] += xt[j]; }
        __syncwarp();
      }
    };
    // first solve: rhs = [-q ; b_red]
    for (int j = lane; j < L.nz; j += 32) {
      const double at = qp.colA(j, [&](int i) { return (ty[i] & 12) ? idel * bred(i) : 0.0; });
      xt[j] = -q[j] + at;
    }
    __syncwarp();
    qp.solve(xt);
    for (int j = lane; j < L.nz; j += 32) { px[j] = xt[j]; dxv[j] = xt[j]; }
    __syncwarp();
    for (int i = lane; i < L.m; i += 32) {
      double v = 0.0;
      if (ty[i] & 12) v = (qp.rowA(i, [&](int j) { return px[j]; }) - bred(i)) * idel;
      py[i] = v;
    }
    __syncwarp();
    inner();
    for (int it = 0; it < S.polish_refine_iter; ++it) {
      // condensed right-hand side of rhs - K s_j: -q - P tx - A'(ty - r2 / delta), r2 = b_red - A tx (row temporary in pz)
      for (int i = lane; i < L.m; i += 32)
        pz[i] = (ty[i] & 12) ? fma(-idel, bred(i) - qp.rowA(i, [&](int j) { return px[j]; }), py[i]) : 0.0;
      __syncwarp();
      for (int j = lane; j < L.nz; j += 32) {
        const double Px = qp.rowP(j, [&](int jj) { return px[jj]; });
        const double At = qp.colA(j, [&](int i) { return pz[i]; });
        xt[j] = (-q[j] - Px) - At;
      }
      __syncwarp();
      qp.solve(xt);
      for (int j = lane; j < L.nz; j += 32) { dxv[j] = xt[j]; px[j] += xt[j]; }
      __syncwarp();
      for (int i = lane; i < L.m; i += 32)
        if (ty[i] & 12) py[i] += idel * (qp.rowA(i, [&](int j) { return px[j]; }) - bred(i));
      __syncwarp();
      inner();
    }
    // pol->z = A x ; project (z, y) onto the normal cone
    for (int i = lane; i < L.m; i += 32) {
      const double zz = qp.rowA(i, [&](int j) { return px[j]; });
      const double t = zz + py[i];
      const double zc = clampd(t, lo[i], up[i]);
      pz[i] = zc;
      py[i] = t - zc;
    }
    __syncwarp();
    // residuals at the polished point
    double a_rp = 0, a_rd = 0;
    for (int i = lane; i < L.m; i += 32) {
      const double r = qp.rowA(i, [&](int j) { return px[j]; }) - pz[i];
      a_rp = absmax(a_rp, unscale ? Einv[i] * r : r);
    }
    for (int j = lane; j < L.nz; j += 32) {
      const double Px = qp.rowP(j, [&](int jj) { return px[jj]; });
      const double Aty = qp.colA(j, [&](int i) { return py[i]; });
      const double r = (q[j] + Px) + Aty;
      a_rd = absmax(a_rd, unscale ? Dinv[j] * r : r);
    }
    const double pol_pri = (L.m == 0) ? 0.0 : warp_max(a_rp);
    const double pol_dua = (unscale ? cinv : 1.0) * warp_max(a_rd);
    const double pol_obj = objective(px);
    const bool ok = (pol_pri < pri_res && pol_dua < dua_res) || (pol_pri < pri_res && dua_res < 1e-10) ||
                    (pol_dua < dua_res && pri_res < 1e-10);
    if (ok) {
      out.obj = pol_obj; pri_res = pol_pri; dua_res = pol_dua; polish_status = 1;
      for (int j = lane; j < L.nz; j += 32) x[j] = px[j];
      for (int i = lane; i < L.m; i += 32) { z[i] = pz[i]; y[i] = py[i]; }
    } else polish_status = -1;
    __syncwarp();
  }
  out.status = status; out.iter = iter_done; out.rho_updates = rho_updates; out.polish_status = polish_status;
  out.pri_res = pri_res; out.dua_res = dua_res;
  if (!has_sol) { /* obj already set to +-infty / nan */ }
}

}  // namespace lpv
