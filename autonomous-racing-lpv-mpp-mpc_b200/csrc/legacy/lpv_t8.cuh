// "T8" kernel: the controller QP (6 states, 2 inputs) with a compile-time horizon, 8 lanes per QP, 4 QPs per warp.
//
// Lane r of a group owns row r of every 8x8 stage block: component r of the stage variable w_k = [x_k; u_k]
// (r < 6: state r, r = 6: steering, r = 7: acceleration), the dynamics row (k, r) and the single-variable
// inequality rows on its variable (vx on lane 0, delta on lane 6, a on lane 7).  All iterates live in
// registers (loops over stages are fully unrolled), the block factor (K_k, T_k) and the scaled dynamics
// blocks live in shared memory (14 KB per QP -> 16 QPs per SM), rarely used data (scalings, P) in an
// L2-resident scratch slab.  Stage-to-stage exchange is an 8-lane all-gather by warp shuffles.
//
// Same algorithm and operation order as the generic kernel (lpv_qp.cuh) / oracle/osqp_ref.c; restricted to
// diagonal Q and R (the reference's tunings, controllerMain.py:139-148) and steering_delay = 0.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../lpv_qp.cuh"

namespace lpv {
namespace t8 {

constexpr int NX = 6, NB = 8, LDB = 9;  // LDB: padded row stride of the 8x8 blocks (bank-conflict free)

template <int N>
struct Reg {  // shared-memory region of one QP, offsets in doubles
  static constexpr int K = 0;                        // K_1..K_N          (block k at (k-1))
  static constexpr int T = K + N * NB * LDB;         // T_0..T_N
  static constexpr int G = T + (N + 1) * NB * LDB;   // Gt_0..Gt_{N-1}: Gt[c*6 + r'] = -(E [A_k B_k] D)[r'][c]
  static constexpr int ED = G + N * NB * NX;         // scaled identity entry of dynamics row (k, r)
  static constexpr int SI = ED + (N + 1) * NX;       // inequality coefficients (k, slot, t)
  static constexpr int UI = SI + N * 6;              // inequality upper bounds (scaled)
  static constexpr int EX = (UI + N * 6 + 1) & ~1;   // 2 x 8 doubles: ping-pong exchange buffer of the 8-lane all-gather
  static constexpr int TOTAL = EX + 16;
};
// cold per-lane data in the global scratch slab (doubles per lane)
template <int N>
struct Cold {
  static constexpr int D = 0, DINV = D + (N + 1), ED = DINV + (N + 1), EDINV = ED + (N + 1), EI = EDINV + (N + 1),
                       EIINV = EI + 2 * N, PD = EIINV + 2 * N, PO = PD + (N + 1),
                       PREV_X = PO + N, PREV_YD = PREV_X + (N + 1), PREV_YI = PREV_YD + (N + 1),  // iterate before the last step
                       SNAP = PREV_YI + 2 * N,                                                     // iterates frozen at termination
                       TOTAL = SNAP + 2 * (N + 1) + 4 * N;
};

__device__ __forceinline__ double gshfl(double v, int src) { return __shfl_sync(kFull, v, src, 8); }
__device__ __forceinline__ double gmax(double v) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) { const double w = __shfl_xor_sync(kFull, v, o, 8); v = (w > v) ? w : v; }
  return v;
}
__device__ __forceinline__ double gsum(double v) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o, 8);
  return v;
}
__device__ __forceinline__ int gany(int v) {  // any lane of my 8-lane group
  const unsigned m = __ballot_sync(kFull, v);
  return ((m >> ((threadIdx.x & 31) & ~7)) & 0xffu) != 0;
}

// 8-lane all-gather through shared memory: 1 STS.64 + 4 broadcast LDS.128 instead of 16 SHFL.32 (+ the register moves
// that re-pair the halves).  Same bytes through the shared-memory crossbar, a third of the instructions — the unrolled
// hot loop is bound by instruction fetch.  PAR alternates so a slow lane never sees the next round's values.
template <int PAR>
__device__ __forceinline__ void gather_smem(double *ex, int r, double v, double (&g)[8]) {
  double *b = ex + PAR * 8;
  b[r] = v;
  __syncwarp();
  const double2 a0 = *reinterpret_cast<const double2 *>(b), a1 = *reinterpret_cast<const double2 *>(b + 2),
                a2 = *reinterpret_cast<const double2 *>(b + 4), a3 = *reinterpret_cast<const double2 *>(b + 6);
  g[0] = a0.x; g[1] = a0.y; g[2] = a1.x; g[3] = a1.y; g[4] = a2.x; g[5] = a2.y; g[6] = a3.x; g[7] = a3.y;
}

template <int N>
struct Lane {
  double X[N + 1], Q[N + 1], YD[N + 1], BE[N + 1];
  double ZI[N][2], YI[N][2];
};

template <int N>
struct Ctx {
  double *S;      // my QP's shared region
  double *cold;   // my lane's cold slab
  int r;          // lane within group
  bool xl, il;    // state lane / inequality lane
  int slot;       // inequality slot 0..2 (lanes 0, 6, 7)
  __device__ __forceinline__ double *Kb(int k) const { return S + Reg<N>::K + (k - 1) * NB * LDB; }
  __device__ __forceinline__ double *Tb(int k) const { return S + Reg<N>::T + k * NB * LDB; }
  __device__ __forceinline__ double *Gt(int k) const { return S + Reg<N>::G + k * NB * NX; }
  __device__ __forceinline__ double &ed(int k) const { return S[Reg<N>::ED + k * NX + (xl ? r : 0)]; }
  __device__ __forceinline__ double &si(int k, int t) const { return S[Reg<N>::SI + (k * 3 + slot) * 2 + t]; }
  __device__ __forceinline__ double &ui(int k, int t) const { return S[Reg<N>::UI + (k * 3 + slot) * 2 + t]; }
};

// ---------------------------------------------------------------- block factorisation (lane r = row r)
// Row weights of the condensed system: ADMM -> rho (rho_eq on the dynamics rows); polish -> 1/delta on active rows.
struct FW {
  int polish;
  double rho, rho_eq, idel;
  unsigned act_d, act_i;  // polish: bit k = my dynamics row (k, r) active; bit 2k+t = my inequality row active
  __device__ __forceinline__ double wd(int k) const { return polish ? (((act_d >> k) & 1u) ? idel : 0.0) : rho_eq; }
  __device__ __forceinline__ double wi(int k, int t) const { return polish ? (((act_i >> (2 * k + t)) & 1u) ? idel : 0.0) : rho; }
};

template <int N>
__device__ __noinline__ void factor(const Ctx<N> c, const FW fw, const double sigma) {
  double PD[N + 1], PO[N];
#pragma unroll
  for (int k = 0; k <= N; ++k) PD[k] = c.cold[Cold<N>::PD + k];
#pragma unroll
  for (int k = 0; k < N; ++k) PO[k] = c.cold[Cold<N>::PO + k];
  __syncwarp();
  const int r = c.r;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    const int nbk = (k < N) ? NB : NX;
    const bool rowlive = r < nbk;
    double s[NB], so[NB];
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) { s[cc] = 0.0; so[cc] = 0.0; }
    // diagonal: P + sigma I + own rows
    {
      double d = PD[k] + sigma;
      if (c.il && k < N) {
        const double a0 = c.si(k, 0), a1 = c.si(k, 1);
        d = fma(fw.wi(k, 0) * a0, a0, d);
        d = fma(fw.wi(k, 1) * a1, a1, d);
      }
      if (c.xl) { const double e = c.ed(k); d = fma(fw.wd(k) * e, e, d); }
      if (!rowlive) d = 1.0;
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) if (cc == r) s[cc] = d;
    }
    // next-stage dynamics rows: sum_r' w(k+1,r') G[r'][r] G[r'][c]
    if (k < N) {
      const double *g = c.Gt(k);
      const double wme = c.xl ? fw.wd(k + 1) : 0.0;
      double col[NX];
#pragma unroll
      for (int rr = 0; rr < NX; ++rr) col[rr] = gshfl(wme, rr) * g[r * NX + rr];
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) {
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) s[cc] = fma(col[rr], g[cc * NX + rr], s[cc]);
      }
    }
    if (k > 0) {
      // my row of S_{k,k-1}
      if (c.xl) {
        const double f = fw.wd(k) * c.ed(k);
        const double *gp = c.Gt(k - 1);
#pragma unroll
        for (int cc = 0; cc < NB; ++cc) so[cc] = f * gp[cc * NX + r];
      } else if (k < N) {
#pragma unroll
        for (int cc = NX; cc < NB; ++cc) if (cc == r) so[cc] = PO[k - 1];
      }
      double *Kk = c.Kb(k);
      const double *Tp = c.Tb(k - 1);
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) Kk[r * LDB + cc] = so[cc];  // park S_{k,k-1} in K_k's slot
      __syncwarp();
      double kr[NB];
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NB; ++j) acc = fma(so[j], Tp[j * LDB + cc], acc);
        kr[cc] = acc;
      }
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) {
        double acc = s[cc];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc = fma(-kr[j], Kk[cc * LDB + j], acc);
        s[cc] = acc;
      }
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) Kk[r * LDB + cc] = kr[cc];
    }
    if (!rowlive) {
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) s[cc] = (cc == r) ? 1.0 : 0.0;
    }
    // Gauss-Jordan inverse of the pivot block, rows across lanes
#pragma unroll
    for (int p = 0; p < NB; ++p) {
      if (p < nbk) {
        double pr[NB];
#pragma unroll
        for (int cc = 0; cc < NB; ++cc) pr[cc] = gshfl(s[cc], p);
        const double piv = 1.0 / pr[p];
        if (r == p) {
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) s[cc] = (cc == p) ? piv : s[cc] * piv;
        } else {
          const double fp = s[p] * piv;
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) s[cc] = (cc == p) ? -fp : fma(-fp, pr[cc], s[cc]);
        }
      }
    }
    double *Tk = c.Tb(k);
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) Tk[r * LDB + cc] = s[cc];
    __syncwarp();
  }
}

// ---------------------------------------------------------------- block solve: rhs b[k] (my row) -> x[k]
// Optionally returns zt[k] = (A x)_dyn(k, r) computed with the gathered stage vectors of the backward sweep.
template <int N>
__device__ __forceinline__ void solve(const Ctx<N> &c, const double (&b)[N + 1], double (&x)[N + 1], double (&zt)[N + 1]) {
  const int r = c.r;
  double W[N + 1];
  double v = b[0];
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    const int nbk = (k < N) ? NB : NX;
    double gv[NB];
    if (k & 1) gather_smem<1>(c.S + Reg<N>::EX, r, v, gv); else gather_smem<0>(c.S + Reg<N>::EX, r, v, gv);
    const double *Tk = c.Tb(k) + r * LDB;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int cc = 0; cc < NB; cc += 2) {
      if (cc < nbk) a0 = fma(Tk[cc], gv[cc], a0);
      if (cc + 1 < nbk) a1 = fma(Tk[cc + 1], gv[cc + 1], a1);
    }
    W[k] = a0 + a1;
    if (k < N) {
      const double *Kn = c.Kb(k + 1) + r * LDB;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int cc = 0; cc < NB; cc += 2) { c0 = fma(Kn[cc], gv[cc], c0); c1 = fma(Kn[cc + 1], gv[cc + 1], c1); }
      v = b[k + 1] - (c0 + c1);
    }
  }
  x[N] = W[N];
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    const int nbn = (k + 1 < N) ? NB : NX;
    double gx[NB];
    if ((N + k) & 1) gather_smem<1>(c.S + Reg<N>::EX, r, x[k + 1], gx); else gather_smem<0>(c.S + Reg<N>::EX, r, x[k + 1], gx);
    const double *Kn = c.Kb(k + 1);
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int rr = 0; rr < NB; rr += 2) {
      if (rr < nbn) c0 = fma(Kn[rr * LDB + r], gx[rr], c0);
      if (rr + 1 < nbn) c1 = fma(Kn[(rr + 1) * LDB + r], gx[rr + 1], c1);
    }
    x[k] = W[k] - (c0 + c1);
    // dynamics rows of stage k+2 read w_{k+1}
    if (k + 2 <= N) {
      const double *g = c.Gt(k + 1);
      double a0 = c.ed(k + 2) * x[k + 2], a1 = 0.0;
#pragma unroll
      for (int cc = 0; cc < NB; cc += 2) { a0 = fma(g[cc * NX + (c.xl ? r : 0)], gx[cc], a0); a1 = fma(g[(cc + 1) * NX + (c.xl ? r : 0)], gx[cc + 1], a1); }
      zt[k + 2] = a0 + a1;
    }
  }
  {
    double gx[NB];
    if ((N + 1) & 1) gather_smem<0>(c.S + Reg<N>::EX, r, x[0], gx); else gather_smem<1>(c.S + Reg<N>::EX, r, x[0], gx);
    const double *g = c.Gt(0);
    double a0 = c.ed(1) * x[1], a1 = 0.0;
#pragma unroll
    for (int cc = 0; cc < NB; cc += 2) { a0 = fma(g[cc * NX + (c.xl ? r : 0)], gx[cc], a0); a1 = fma(g[(cc + 1) * NX + (c.xl ? r : 0)], gx[cc + 1], a1); }
    zt[1] = a0 + a1;
    zt[0] = c.ed(0) * x[0];
  }
}

// (A v)_dyn(k, r) for an arbitrary stage vector v (my rows); used by checks and polish
template <int N>
__device__ __forceinline__ void rowsA(const Ctx<N> &c, const double (&v)[N + 1], double (&out)[N + 1]) {
  const int r = c.xl ? c.r : 0;
  out[0] = c.ed(0) * v[0];
#pragma unroll
  for (int k = 1; k <= N; ++k) {
    const double *g = c.Gt(k - 1);
    double a0 = c.ed(k) * v[k], a1 = 0.0;
#pragma unroll
    for (int cc = 0; cc < NB; cc += 2) {
      a0 = fma(g[cc * NX + r], gshfl(v[k - 1], cc), a0);
      a1 = fma(g[(cc + 1) * NX + r], gshfl(v[k - 1], cc + 1), a1);
    }
    out[k] = a0 + a1;
  }
}
// (A' t)_(k, r): td = values on my dynamics rows (0 on u lanes), ti = values on my inequality rows
template <int N>
__device__ __forceinline__ void colsA(const Ctx<N> &c, const double (&td)[N + 1], const double (&ti)[N][2], double (&out)[N + 1]) {
  const int r = c.r;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    double acc = c.xl ? c.ed(k) * td[k] : 0.0;
    if (k < N) {
      const double *g = c.Gt(k) + r * NX;
      double a1 = 0.0;
      double gt[NB];
      if (k & 1) gather_smem<1>(c.S + Reg<N>::EX, r, td[k + 1], gt); else gather_smem<0>(c.S + Reg<N>::EX, r, td[k + 1], gt);
#pragma unroll
      for (int rr = 0; rr < NX; rr += 2) {
        acc = fma(g[rr], gt[rr], acc);
        a1 = fma(g[rr + 1], gt[rr + 1], a1);
      }
      acc += a1;
      if (c.il) { acc = fma(c.si(k, 0), ti[k][0], acc); acc = fma(c.si(k, 1), ti[k][1], acc); }
    }
    out[k] = acc;
  }
}
// (P v)_(k, r) with diagonal Q, R and the slew-rate coupling of the inputs
template <int N>
__device__ __forceinline__ void rowsP(const Ctx<N> &c, const double (&PD)[N + 1], const double (&PO)[N], const double (&v)[N + 1],
                                      double (&out)[N + 1]) {
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    double acc = PD[k] * v[k];
    if (!c.xl) {
      if (k > 0 && k < N) acc = fma(PO[k - 1], v[k - 1], acc);
      if (k < N - 1) acc = fma(PO[k], v[k + 1], acc);
      if (k == N) acc = 0.0;
    }
    out[k] = acc;
  }
}

// ---------------------------------------------------------------- per-QP scalars shared by the cold routines
struct Info {
  double pri_res, dua_res, obj;
  double n_rp, n_z, n_Ax, n_rd, n_q, n_Aty, n_Px;  // scaled-space inf-norms of the last update_info
  double u_z, u_Ax, u_q, u_Aty, u_Px;              // the same, unscaled (termination)
  double csc, cinv;
  int status, unscale;
};

// ---------------------------------------------------------------- setup: schedule + build + Ruiz (cold, once per QP)
template <int N>
__device__ __noinline__ void setup(const Ctx<N> c, const Params &p, const int b, const bool valid, Lane<N> *Lp, double *csc_out,
                                   int *sched_err_out) {
  Lane<N> &L = *Lp;
  const int r = c.r;
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &S = p.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N;
  const int ucomp = r - NX;
  double PD[N + 1], PO[N], D[N + 1], Ed[N + 1], Ei[N][2];
  int sched_err = 0;
  double x0r = 0.0;
  // ---- schedule: Gt_k = -[A_k B_k] (unscaled)
  if (a.sched_mode == LPVMPC_SCHED_GIVEN) {
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      double *gk = c.Gt(k);
      if (c.xl) {
        const double *Ar = a.A + ((size_t)b * N + k) * 36 + r * 6, *Br = a.Bm + ((size_t)b * N + k) * 12 + r * 2;
#pragma unroll
        for (int cc = 0; cc < NX; ++cc) gk[cc * NX + r] = -Ar[cc];
        gk[6 * NX + r] = -Br[0]; gk[7 * NX + r] = -Br[1];
      }
    }
    x0r = c.xl ? a.x0[(size_t)b * NX + r] : 0.0;
  } else {
    const bool predict = a.sched_mode == LPVMPC_SCHED_PREDICT;
    double st[NX];
    const double *xs = (predict && a.x_sched) ? a.x_sched + (size_t)b * NX : a.x0 + (size_t)b * NX;
#pragma unroll
    for (int q = 0; q < NX; ++q) st[q] = xs[q];
    const double *up = a.u_prev + (size_t)b * N * 2;
    const int lap = a.lap ? a.lap[b] : a.lap_all;
    x0r = c.xl ? a.x0[(size_t)b * NX + r] : 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      double vx, vy, epsi, ey, cur, Cf, Cr;
      const double delta = up[k * 2];
      if (predict) {
        vy = st[1]; epsi = st[3]; ey = st[5];
        cur = (lap == 0) ? curvature(M.track, M.nseg, st[4], sched_err) : a.curv_ref[(size_t)b * N + k];
        vx = a.vel_ref[(size_t)b * (N + 1) + k];
        Cf = a.Cf_new; Cr = a.Cf_new;
      } else {
        const double *t = a.traj + ((size_t)b * N + k) * 6;
        vx = t[0]; vy = t[1]; epsi = t[3]; ey = t[5];
        cur = curvature(M.track, M.nseg, t[4], sched_err);
        Cf = M.Cf; Cr = M.Cr;
      }
      double Ai[36], Bi[12];
      ctrl_stage(M, Cf, Cr, vx, vy, epsi, ey, cur, delta, Ai, Bi);
      double *gk = c.Gt(k);
      double row[NB];  // my row (x lanes) of [A B], selected without dynamic register indexing
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) {
        double v = 0.0;
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) {
          const double e = (cc < NX) ? Ai[rr * NX + cc] : Bi[rr * 2 + (cc - NX)];
          v = (r == rr) ? e : v;
        }
        row[cc] = v;
      }
      if (c.xl) {
#pragma unroll
        for (int cc = 0; cc < NB; ++cc) gk[cc * NX + r] = -row[cc];
        if (valid && a.A_out) {
#pragma unroll
          for (int cc = 0; cc < NX; ++cc) a.A_out[((size_t)b * N + k) * 36 + r * 6 + cc] = row[cc];
        }
        if (valid && a.B_out) { a.B_out[((size_t)b * N + k) * 12 + r * 2] = row[6]; a.B_out[((size_t)b * N + k) * 12 + r * 2 + 1] = row[7]; }
      }
      if (predict) {
        propagate<NX>(Ai, Bi, up + k * 2, st);
        if (c.xl) {
          double mine = 0.0;
#pragma unroll
          for (int rr = 0; rr < NX; ++rr) mine = (r == rr) ? st[rr] : mine;
          if (valid && a.states_out) a.states_out[((size_t)b * N + k) * NX + r] = mine;
          if (k == 0 && a.x0_from_prediction) x0r = mine;
        }
      }
    }
  }
  __syncwarp();
  *sched_err_out = gany(sched_err);

  // ---- build (PathFollowingLPVMPC.py:334-348, 397-464)
  {
    const double Qrr = c.xl ? M.Q[r * NX + r] : 0.0;
    const double Q0r = c.xl ? M.Q[r] : 0.0;
    const double Rcc = c.xl ? 0.0 : M.R[ucomp * 2 + ucomp];
    const double dRc = c.xl ? 0.0 : M.dR[ucomp];
    const double uold = (!c.xl && a.u_old) ? a.u_old[(size_t)b * 2 + ucomp] : 0.0;
#pragma unroll
    for (int k = 0; k <= N; ++k) {
      if (c.xl) {
        PD[k] = 2 * Qrr;
        L.Q[k] = -2 * (a.vel_ref[(size_t)b * (N + 1) + k] * Q0r);
        L.BE[k] = (k == 0) ? (x0r + 0.0) : (0.0 + (a.C ? a.C[((size_t)b * N + (k - 1)) * NX + r] : 0.0));
        c.ed(k) = 1.0;
      } else {
        double v = Rcc + 2 * dRc;
        if (k == N - 1) v = v - dRc;
        PD[k] = (k < N) ? 2 * v : 0.0;
        L.Q[k] = (k == 0) ? -2 * (uold * dRc) : -2 * 0.0;
        L.BE[k] = 0.0;
      }
      D[k] = 1.0; Ed[k] = 1.0;
      L.X[k] = 0.0; L.YD[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
      PO[k] = (!c.xl && k < N - 1) ? 2 * (-dRc) : 0.0;
      Ei[k][0] = 1.0; Ei[k][1] = 1.0;
      L.ZI[k][0] = L.ZI[k][1] = 0.0; L.YI[k][0] = L.YI[k][1] = 0.0;
      if (c.il) {
        c.si(k, 0) = (r == 0) ? -1.0 : 1.0;
        c.si(k, 1) = (r == 0) ? 1.0 : -1.0;
        c.ui(k, 0) = (r == 0) ? -0.01 : (r == 6 ? 0.249 : 4.0);
        c.ui(k, 1) = (r == 0) ? ((M.max_vel < kInfty) ? M.max_vel : kInfty) : (r == 6 ? 0.249 : 1.0);
      }
    }
    __syncwarp();
  }

  // ---- Ruiz equilibration (OSQP scale_data)
  double csc = 1.0;
  if (S.scaling) {
    double *scr = c.S + Reg<N>::K;  // the factor area is free during setup: column norms in reference order
#pragma unroll 1
    for (int it = 0; it < S.scaling; ++it) {
      double Dt[N + 1], Etd[N + 1], Eti[N][2];
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        double pa = fabs(PD[k]);
        if (!c.xl) {
          if (k < N - 1) pa = absmax(pa, PO[k]);
          if (k > 0 && k < N) pa = absmax(pa, PO[k - 1]);
        }
        double qa = c.xl ? fabs(c.ed(k)) : 0.0;
        if (k < N) {
          const double *gk = c.Gt(k) + r * NX;
#pragma unroll
          for (int rr = 0; rr < NX; ++rr) qa = absmax(qa, gk[rr]);
          if (c.il) { qa = absmax(qa, c.si(k, 0)); qa = absmax(qa, c.si(k, 1)); }
        }
        Dt[k] = 1.0 / sqrt(limit_scaling(pa > qa ? pa : qa));
        double ea = c.xl ? fabs(c.ed(k)) : 0.0;
        if (k > 0 && c.xl) {
          const double *gp = c.Gt(k - 1);
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) ea = absmax(ea, gp[cc * NX + r]);
        }
        Etd[k] = 1.0 / sqrt(limit_scaling(ea));
      }
#pragma unroll
      for (int k = 0; k < N; ++k) {
        Eti[k][0] = 1.0 / sqrt(limit_scaling(c.il ? fabs(c.si(k, 0)) : 1.0));
        Eti[k][1] = 1.0 / sqrt(limit_scaling(c.il ? fabs(c.si(k, 1)) : 1.0));
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        if (k < N) {
          double *gk = c.Gt(k) + r * NX;
#pragma unroll
          for (int rr = 0; rr < NX; ++rr) gk[rr] = (gk[rr] * gshfl(Etd[k + 1], rr)) * Dt[k];
          if (c.il) { c.si(k, 0) = (c.si(k, 0) * Eti[k][0]) * Dt[k]; c.si(k, 1) = (c.si(k, 1) * Eti[k][1]) * Dt[k]; }
          Ei[k][0] = Ei[k][0] * Eti[k][0]; Ei[k][1] = Ei[k][1] * Eti[k][1];
          if (k < N - 1) PO[k] = (PO[k] * Dt[k]) * Dt[k + 1];
        }
        if (c.xl) c.ed(k) = (c.ed(k) * Etd[k]) * Dt[k];
        PD[k] = (PD[k] * Dt[k]) * Dt[k];
        L.Q[k] = Dt[k] * L.Q[k];
        D[k] = D[k] * Dt[k];
        Ed[k] = Ed[k] * Etd[k];
      }
      // cost scaling: mean of the column norms of P in the reference variable order
      double qn = 0.0;
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        double pa = fabs(PD[k]);
        if (!c.xl) {
          if (k < N - 1) pa = absmax(pa, PO[k]);
          if (k > 0 && k < N) pa = absmax(pa, PO[k - 1]);
        }
        if (c.xl) scr[k * NX + r] = pa;
        else if (k < N) scr[nx + k * 2 + ucomp] = pa;
        if (c.xl || k < N) qn = absmax(qn, L.Q[k]);
      }
      qn = gmax(qn);
      __syncwarp();
      double ct = 0.0;
#pragma unroll 2
      for (int j = 0; j < nz; ++j) ct += scr[j];
      ct = ct / nz;
      qn = limit_scaling(qn);
      ct = ct > qn ? ct : qn;
      ct = limit_scaling(ct);
      ct = 1.0 / ct;
#pragma unroll
      for (int k = 0; k <= N; ++k) { PD[k] *= ct; L.Q[k] *= ct; }
#pragma unroll
      for (int k = 0; k < N; ++k) PO[k] *= ct;
      csc *= ct;
      __syncwarp();
    }
  }
  *csc_out = csc;
  // ---- scaled bounds and cold data
  {
    double *cd = c.cold;
#pragma unroll
    for (int k = 0; k <= N; ++k) {
      L.BE[k] = Ed[k] * L.BE[k];
      cd[Cold<N>::D + k] = D[k]; cd[Cold<N>::DINV + k] = 1.0 / D[k];
      cd[Cold<N>::ED + k] = Ed[k]; cd[Cold<N>::EDINV + k] = 1.0 / Ed[k];
      cd[Cold<N>::PD + k] = PD[k];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
      cd[Cold<N>::PO + k] = PO[k];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        cd[Cold<N>::EI + 2 * k + t] = Ei[k][t]; cd[Cold<N>::EIINV + 2 * k + t] = 1.0 / Ei[k][t];
        if (c.il) c.ui(k, t) = Ei[k][t] * c.ui(k, t);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- residual norms at the current iterate (update_info)
template <int N>
__device__ __noinline__ void update_info(const Ctx<N> c, const Lane<N> *Lp, Info *ip) {
  const Lane<N> &L = *Lp;
  Info &I = *ip;
  const double *cd = c.cold;
  double Ax[N + 1], Px[N + 1], Aty[N + 1], tdv[N + 1], tiv[N][2];
  rowsA<N>(c, L.X, Ax);
  double a_rp = 0, a_z = 0, a_Ax = 0, b_rp = 0, b_z = 0, b_Ax = 0;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    if (c.xl) {
      const double z = L.BE[k], rr = Ax[k] - z, ei = cd[Cold<N>::EDINV + k];
      a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, Ax[k]);
      b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * Ax[k]);
    }
    tdv[k] = c.xl ? L.YD[k] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      tiv[k][t] = L.YI[k][t];
      if (c.il) {
        const double ax = c.si(k, t) * L.X[k], z = L.ZI[k][t], rr = ax - z, ei = cd[Cold<N>::EIINV + 2 * k + t];
        a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, ax);
        b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * ax);
      }
    }
  }
  double PDc[N + 1], POc[N];
#pragma unroll
  for (int k = 0; k <= N; ++k) PDc[k] = cd[Cold<N>::PD + k];
#pragma unroll
  for (int k = 0; k < N; ++k) POc[k] = cd[Cold<N>::PO + k];
  rowsP<N>(c, PDc, POc, L.X, Px);
  colsA<N>(c, tdv, tiv, Aty);
  double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, b_rd = 0, b_q = 0, b_Aty = 0, b_Px = 0;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    if (c.xl || k < N) {
      const double rr = (L.Q[k] + Px[k]) + Aty[k], di = cd[Cold<N>::DINV + k];
      a_rd = absmax(a_rd, rr); a_q = absmax(a_q, L.Q[k]); a_Aty = absmax(a_Aty, Aty[k]); a_Px = absmax(a_Px, Px[k]);
      b_rd = absmax(b_rd, di * rr); b_q = absmax(b_q, di * L.Q[k]); b_Aty = absmax(b_Aty, di * Aty[k]); b_Px = absmax(b_Px, di * Px[k]);
    }
  }
  I.n_rp = gmax(a_rp); I.n_z = gmax(a_z); I.n_Ax = gmax(a_Ax); I.n_rd = gmax(a_rd); I.n_q = gmax(a_q); I.n_Aty = gmax(a_Aty); I.n_Px = gmax(a_Px);
  if (I.unscale) {
    I.pri_res = gmax(b_rp); I.u_z = gmax(b_z); I.u_Ax = gmax(b_Ax);
    I.dua_res = I.cinv * gmax(b_rd); I.u_q = gmax(b_q); I.u_Aty = gmax(b_Aty); I.u_Px = gmax(b_Px);
  } else {
    I.pri_res = I.n_rp; I.u_z = I.n_z; I.u_Ax = I.n_Ax; I.dua_res = I.n_rd; I.u_q = I.n_q; I.u_Aty = I.n_Aty; I.u_Px = I.n_Px;
  }
}

// ---------------------------------------------------------------- infeasibility certificates (rare)
// delta_y / delta_x are rebuilt from the iterate saved before the last ADMM step (cold PREV_*).
template <int N>
__device__ __noinline__ bool primal_infeasible(const Ctx<N> c, const Lane<N> *Lp, const Info *ip, const double eps) {
  const Lane<N> &L = *Lp;
  const bool unscale = ip->unscale;
  const double *cd = c.cold;
  double DYD[N + 1], DYI[N][2];
  double nrm = 0.0;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    DYD[k] = c.xl ? L.YD[k] - cd[Cold<N>::PREV_YD + k] : 0.0;  // equality rows: both bounds finite, no projection
    nrm = absmax(nrm, unscale ? cd[Cold<N>::ED + k] * DYD[k] : DYD[k]);
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      double d = c.il ? L.YI[k][t] - cd[Cold<N>::PREV_YI + 2 * k + t] : 0.0;
      d = (d > 0.0) ? d : 0.0;  // l = -inf: project onto the polar of the recession cone
      DYI[k][t] = d;
      nrm = absmax(nrm, unscale ? cd[Cold<N>::EI + 2 * k + t] * d : d);
    }
  nrm = gmax(nrm);
  double lhs = 0.0;
#pragma unroll
  for (int k = 0; k <= N; ++k) if (c.xl) { const double d = DYD[k]; lhs += L.BE[k] * ((d > 0) ? d : 0) + L.BE[k] * ((d < 0) ? d : 0); }
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int t = 0; t < 2; ++t) if (c.il) lhs += c.ui(k, t) * DYI[k][t];
  lhs = gsum(lhs);
  double at[N + 1];
  colsA<N>(c, DYD, DYI, at);
  double mx = 0.0;
#pragma unroll
  for (int k = 0; k <= N; ++k) if (c.xl || k < N) mx = absmax(mx, unscale ? cd[Cold<N>::DINV + k] * at[k] : at[k]);
  mx = gmax(mx);
  return (nrm > eps) && (lhs < -eps * nrm) && (mx < eps * nrm);
}

template <int N>
__device__ __noinline__ bool dual_infeasible(const Ctx<N> c, const Lane<N> *Lp, const Info *ip, const double eps) {
  const Lane<N> &L = *Lp;
  const bool unscale = ip->unscale;
  const double *cd = c.cold;
  double dx[N + 1], nrm = 0.0, qdx = 0.0;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    dx[k] = (c.xl || k < N) ? L.X[k] - cd[Cold<N>::PREV_X + k] : 0.0;
    nrm = absmax(nrm, unscale ? cd[Cold<N>::D + k] * dx[k] : dx[k]);
    qdx += L.Q[k] * dx[k];
  }
  nrm = gmax(nrm); qdx = gsum(qdx);
  const double cs = unscale ? ip->csc : 1.0;
  double PDc[N + 1], POc[N], Pdx[N + 1], Adx[N + 1];
#pragma unroll
  for (int k = 0; k <= N; ++k) PDc[k] = cd[Cold<N>::PD + k];
#pragma unroll
  for (int k = 0; k < N; ++k) POc[k] = cd[Cold<N>::PO + k];
  rowsP<N>(c, PDc, POc, dx, Pdx);
  double mx = 0.0;
#pragma unroll
  for (int k = 0; k <= N; ++k) if (c.xl || k < N) mx = absmax(mx, unscale ? cd[Cold<N>::DINV + k] * Pdx[k] : Pdx[k]);
  mx = gmax(mx);
  rowsA<N>(c, dx, Adx);
  int viol = 0;
#pragma unroll
  for (int k = 0; k <= N; ++k) if (c.xl) {
    const double v = unscale ? cd[Cold<N>::EDINV + k] * Adx[k] : Adx[k];
    if (v > eps * nrm || v < -eps * nrm) viol = 1;
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int t = 0; t < 2; ++t) if (c.il) {
      double v = c.si(k, t) * dx[k];
      if (unscale) v = cd[Cold<N>::EIINV + 2 * k + t] * v;
      if (v > eps * nrm) viol = 1;  // u finite, l = -inf
    }
  viol = gany(viol);
  return (nrm > eps) && (qdx < -cs * eps * nrm) && (mx < cs * eps * nrm) && !viol;
}

// returns 1 when a termination status was set for my group (check_termination)
template <int N>
__device__ __noinline__ int check_termination(const Ctx<N> c, const lpvmpc_settings &S, const Lane<N> *Lp, Info *ip, const bool live,
                                              const int approximate) {
  Info &I = *ip;
  double eps_abs = S.eps_abs, eps_rel = S.eps_rel, eps_pi = S.eps_prim_inf, eps_di = S.eps_dual_inf;
  const bool ncvx = (I.pri_res > kInfty) || (I.dua_res > kInfty);
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; eps_di *= 10; }
  const double eps_prim = eps_abs + eps_rel * (I.u_z > I.u_Ax ? I.u_z : I.u_Ax);
  const bool prim_ok = I.pri_res < eps_prim;
  double mr = I.u_q; mr = (I.u_Aty > mr) ? I.u_Aty : mr; mr = (I.u_Px > mr) ? I.u_Px : mr;
  if (I.unscale) mr *= I.cinv;
  const double eps_dual = eps_abs + eps_rel * mr;
  const bool dual_ok = I.dua_res < eps_dual;
  bool prim_inf = false, dual_inf = false;
  if (__any_sync(kFull, live && !ncvx && !prim_ok)) prim_inf = primal_infeasible<N>(c, Lp, ip, eps_pi) && !prim_ok;
  if (__any_sync(kFull, live && !ncvx && !dual_ok)) dual_inf = dual_infeasible<N>(c, Lp, ip, eps_di) && !dual_ok;
  if (!live) return 0;
  if (ncvx) { I.status = LPVMPC_NON_CVX; I.obj = nan(""); return 1; }
  if (prim_ok && dual_ok) { I.status = approximate ? LPVMPC_SOLVED_INACCURATE : LPVMPC_SOLVED; return 1; }
  if (prim_inf) { I.status = approximate ? LPVMPC_PRIMAL_INFEASIBLE_INACCURATE : LPVMPC_PRIMAL_INFEASIBLE; I.obj = kInfty; return 1; }
  if (dual_inf) { I.status = approximate ? LPVMPC_DUAL_INFEASIBLE_INACCURATE : LPVMPC_DUAL_INFEASIBLE; I.obj = -kInfty; return 1; }
  return 0;
}

template <int N>
__device__ __forceinline__ double objective(const Ctx<N> &c, const double (&Q)[N + 1], const double (&xv)[N + 1], double scale) {
  double PDc[N + 1], POc[N], Px[N + 1], acc = 0.0;
#pragma unroll
  for (int k = 0; k <= N; ++k) PDc[k] = c.cold[Cold<N>::PD + k];
#pragma unroll
  for (int k = 0; k < N; ++k) POc[k] = c.cold[Cold<N>::PO + k];
  rowsP<N>(c, PDc, POc, xv, Px);
#pragma unroll
  for (int k = 0; k <= N; ++k) if (c.xl || k < N) acc += (0.5 * Px[k] + Q[k]) * xv[k];
  return gsum(acc) * scale;
}

// ---------------------------------------------------------------- polish (cold, once per QP)
// act_*: my rows' active-set guess; on success the polished (x, z, y) replace the iterate.
template <int N>
__device__ __noinline__ int polish(const Ctx<N> c, const lpvmpc_settings &S, Lane<N> *Lp, Info *ip, const bool do_pol, unsigned *acts) {
  Lane<N> &L = *Lp;
  Info &I = *ip;
  const bool unscale = I.unscale;
  const double *cd = c.cold;
  unsigned act_lo_d = 0, act_up_d = 0, act_lo_i = 0, act_up_i = 0;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    if (c.xl) {  // equality row: z == l == u
      if (0.0 < -L.YD[k]) act_lo_d |= 1u << k;
      if (0.0 < L.YD[k]) act_up_d |= 1u << k;
    }
    if (k < N && c.il) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const double lo = -kInfty * cd[Cold<N>::EI + 2 * k + t];
        if (L.ZI[k][t] - lo < -L.YI[k][t]) act_lo_i |= 1u << (2 * k + t);
        if (c.ui(k, t) - L.ZI[k][t] < L.YI[k][t]) act_up_i |= 1u << (2 * k + t);
      }
    }
  }
  acts[0] = act_lo_d; acts[1] = act_up_d; acts[2] = act_lo_i; acts[3] = act_up_i;
  const unsigned act_d = act_lo_d | act_up_d, act_i = act_lo_i | act_up_i;
  const double delta = S.delta, idel = 1.0 / S.delta;
  FW fw; fw.polish = 1; fw.rho = 0.0; fw.rho_eq = 0.0; fw.idel = idel; fw.act_d = act_d; fw.act_i = act_i;
  __syncwarp();
  factor<N>(c, fw, delta);
  auto bred_i = [&](int k, int t) { return ((act_lo_i >> (2 * k + t)) & 1u) ? (-kInfty * cd[Cold<N>::EI + 2 * k + t]) : c.ui(k, t); };
  double px[N + 1], pyd[N + 1], pyi[N][2], zt[N + 1];
  {
    double tdv[N + 1], tiv[N][2], at[N + 1], bv[N + 1];
#pragma unroll
    for (int k = 0; k <= N; ++k) tdv[k] = (c.xl && ((act_d >> k) & 1u)) ? idel * L.BE[k] : 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
      for (int t = 0; t < 2; ++t) tiv[k][t] = (c.il && ((act_i >> (2 * k + t)) & 1u)) ? idel * bred_i(k, t) : 0.0;
    colsA<N>(c, tdv, tiv, at);
#pragma unroll
    for (int k = 0; k <= N; ++k) bv[k] = -L.Q[k] + at[k];
    solve<N>(c, bv, px, zt);
#pragma unroll
    for (int k = 0; k <= N; ++k) pyd[k] = (c.xl && ((act_d >> k) & 1u)) ? (zt[k] - L.BE[k]) * idel : 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
      for (int t = 0; t < 2; ++t) pyi[k][t] = (c.il && ((act_i >> (2 * k + t)) & 1u)) ? (c.si(k, t) * px[k] - bred_i(k, t)) * idel : 0.0;
  }
  double PDc[N + 1], POc[N];
#pragma unroll
  for (int k = 0; k <= N; ++k) PDc[k] = cd[Cold<N>::PD + k];
#pragma unroll
  for (int k = 0; k < N; ++k) POc[k] = cd[Cold<N>::PO + k];
#pragma unroll 1
  for (int it = 0; it < S.polish_refine_iter + kPolishExtraRefine; ++it) {
    double Ax[N + 1], r2d[N + 1], r2i[N][2], Px[N + 1], Aty[N + 1], at[N + 1], bv[N + 1], dx[N + 1], tdv[N + 1], tiv[N][2];
    rowsA<N>(c, px, Ax);
#pragma unroll
    for (int k = 0; k <= N; ++k) { r2d[k] = (c.xl && ((act_d >> k) & 1u)) ? (L.BE[k] - Ax[k]) : 0.0; tdv[k] = idel * r2d[k]; }
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        r2i[k][t] = (c.il && ((act_i >> (2 * k + t)) & 1u)) ? (bred_i(k, t) - c.si(k, t) * px[k]) : 0.0;
        tiv[k][t] = idel * r2i[k][t];
      }
    rowsP<N>(c, PDc, POc, px, Px);
    colsA<N>(c, pyd, pyi, Aty);
    colsA<N>(c, tdv, tiv, at);
#pragma unroll
    for (int k = 0; k <= N; ++k) bv[k] = ((-L.Q[k] - Px[k]) - Aty[k]) + at[k];
    solve<N>(c, bv, dx, zt);
#pragma unroll
    for (int k = 0; k <= N; ++k) {
      if (c.xl && ((act_d >> k) & 1u)) pyd[k] += (zt[k] - r2d[k]) * idel;
      px[k] += dx[k];
    }
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
      for (int t = 0; t < 2; ++t) if (c.il && ((act_i >> (2 * k + t)) & 1u)) pyi[k][t] += (c.si(k, t) * dx[k] - r2i[k][t]) * idel;
  }
  // pol z = A x, normal-cone projection, residuals, acceptance
  double Ax[N + 1], pzi[N][2], Px[N + 1], Aty[N + 1];
  rowsA<N>(c, px, Ax);
  double a_rp = 0, a_rd = 0;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    if (c.xl) {
      const double t = Ax[k] + pyd[k];
      pyd[k] = t - L.BE[k];
      const double rr = Ax[k] - L.BE[k];
      a_rp = absmax(a_rp, unscale ? cd[Cold<N>::EDINV + k] * rr : rr);
    } else pyd[k] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (c.il) {
        const double ax = c.si(k, t) * px[k], tt = ax + pyi[k][t];
        const double lo = -kInfty * cd[Cold<N>::EI + 2 * k + t], up = c.ui(k, t);
        const double z0 = (tt > lo) ? tt : lo, zc = (z0 < up) ? z0 : up;
        pzi[k][t] = zc; pyi[k][t] = tt - zc;
        const double rr = ax - zc;
        a_rp = absmax(a_rp, unscale ? cd[Cold<N>::EIINV + 2 * k + t] * rr : rr);
      } else { pzi[k][t] = 0.0; pyi[k][t] = 0.0; }
    }
  rowsP<N>(c, PDc, POc, px, Px);
  colsA<N>(c, pyd, pyi, Aty);
#pragma unroll
  for (int k = 0; k <= N; ++k) if (c.xl || k < N) {
    const double rr = (L.Q[k] + Px[k]) + Aty[k];
    a_rd = absmax(a_rd, unscale ? cd[Cold<N>::DINV + k] * rr : rr);
  }
  const double pol_pri = gmax(a_rp), pol_dua = (unscale ? I.cinv : 1.0) * gmax(a_rd);
  const double pol_obj = objective<N>(c, L.Q, px, S.scaling ? I.cinv : 1.0);
  const bool ok = (pol_pri < I.pri_res && pol_dua < I.dua_res) || (pol_pri < I.pri_res && I.dua_res < 1e-10) ||
                  (pol_dua < I.dua_res && I.pri_res < 1e-10);
  if (!do_pol) return 0;
  if (!ok) return -1;
  I.obj = pol_obj; I.pri_res = pol_pri; I.dua_res = pol_dua;
#pragma unroll
  for (int k = 0; k <= N; ++k) { L.X[k] = px[k]; L.YD[k] = pyd[k]; }
#pragma unroll
  for (int k = 0; k < N; ++k) { L.ZI[k][0] = pzi[k][0]; L.ZI[k][1] = pzi[k][1]; L.YI[k][0] = pyi[k][0]; L.YI[k][1] = pyi[k][1]; }
  return 1;
}

// ---------------------------------------------------------------- one ADMM step (hot; everything in registers)
// zsel = 0 on the very first step (z_dyn is the cold-start zero), 1 afterwards (z_dyn == equality bound).
template <int N>
__device__ __forceinline__ void admm_step(const Ctx<N> &c, Lane<N> &L, const double rho, const double rho_eq, const double rinv,
                                          const double sigma, const double alpha, const double zsel) {
  double bv[N + 1], xt[N + 1], zt[N + 1];
  {
    double tdv[N + 1], tiv[N][2], at[N + 1];
#pragma unroll
    for (int k = 0; k <= N; ++k) tdv[k] = rho_eq * (zsel * L.BE[k]) - L.YD[k];   // u lanes: BE = YD = 0
#pragma unroll
    for (int k = 0; k < N; ++k) { tiv[k][0] = rho * L.ZI[k][0] - L.YI[k][0]; tiv[k][1] = rho * L.ZI[k][1] - L.YI[k][1]; }
    colsA<N>(c, tdv, tiv, at);
#pragma unroll
    for (int k = 0; k <= N; ++k) bv[k] = (sigma * L.X[k] - L.Q[k]) + at[k];
  }
  solve<N>(c, bv, xt, zt);
  const double oma = 1.0 - alpha;
#pragma unroll
  for (int k = 0; k <= N; ++k) {
    if (c.xl) {
      const double zr = alpha * zt[k] + oma * (zsel * L.BE[k]);
      L.YD[k] += rho_eq * (zr - L.BE[k]);
    }
    if (k < N && c.il) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const double zr = alpha * (c.si(k, t) * xt[k]) + oma * L.ZI[k][t];
        const double zc = zr + rinv * L.YI[k][t];
        const double up = c.ui(k, t);
        const double zn = (zc < up) ? zc : up;  // l = -1e30 E_i cannot bind on a finite iterate
        L.YI[k][t] += rho * (zr - zn);
        L.ZI[k][t] = zn;
      }
    }
    L.X[k] = alpha * xt[k] + oma * L.X[k];
  }
}

// ---------------------------------------------------------------- one persistent warp = 4 QPs at a time
// QPW = QPs per warp (4 or 2).  With QPW = 2 the upper half-warp mirrors the lower one (same problems, same shared
// and scratch regions, same values written by the same instruction) and 8 warps instead of 4 fit an SM: the unrolled
// code is bound by instruction fetch per warp (profiles/r1_icache_probe.jsonl), so two half-empty warps per SM
// sub-partition issue about twice as fast as one full warp.
template <int N, int QPW>
__global__ void __launch_bounds__(32, 16 / QPW) lpv_solve_t8_kernel(const __grid_constant__ Params p, unsigned *queue, double *cold_slab) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int g = lane >> 3, r = lane & 7;
  const int gg = g % QPW;
  Ctx<N> c;
  c.S = smem + gg * Reg<N>::TOTAL;
  c.cold = cold_slab + ((size_t)(blockIdx.x * QPW + gg) * 8 + r) * Cold<N>::TOTAL;
  c.r = r; c.xl = r < NX; c.il = (r == 0) || (r >= NX); c.slot = (r == 0) ? 0 : ((r >= NX) ? r - 5 : 0);
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &S = p.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N, m = 6 * N + nx;
  const int ucomp = r - NX;

  for (;;) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(queue, (unsigned)QPW);
    base = __shfl_sync(kFull, base, 0);
    if ((int)base >= p.B) break;
    const bool valid = (g < QPW) && ((int)(base + gg) < p.B);
    const int b = ((int)(base + gg) < p.B) ? (int)(base + gg) : (int)base;  // idle groups shadow another group and never write results

    Lane<N> L;
    Info I;
    int sched_err = 0;
    {
      Lane<N> tmp;
      double csc;
      setup<N>(c, p, b, valid, &tmp, &csc, &sched_err);
      L = tmp;
      I.csc = csc; I.cinv = 1.0 / csc;
    }
    I.unscale = (S.scaling && !S.scaled_termination) ? 1 : 0;
    I.pri_res = 0.0; I.dua_res = 0.0; I.obj = nan("");
    I.n_rp = I.n_z = I.n_Ax = I.n_rd = I.n_q = I.n_Aty = I.n_Px = 0.0;
    I.u_z = I.u_Ax = I.u_q = I.u_Aty = I.u_Px = 0.0;
    I.status = sched_err ? LPVMPC_SCHEDULE_ERROR : LPVMPC_UNSOLVED;

    const double sigma = S.sigma, alpha = S.alpha;
    double rho = fmin(fmax(S.rho, kRhoMin), kRhoMax);
    double rho_eq = kRhoEqOverIneq * rho, rinv = 1.0 / rho;
    {
      FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0; fw.act_d = 0; fw.act_i = 0;
      factor<N>(c, fw, sigma);
    }
    bool live = !sched_err;
    int iter_done = 0, rho_updates = 0;
    int adapt_interval = S.adaptive_rho_interval;
    if (S.adaptive_rho && !adapt_interval) adapt_interval = S.check_termination ? 4 * S.check_termination : 100;
    const int ct = S.check_termination, ai = S.adaptive_rho ? adapt_interval : 0;

    int iter = 0;
    bool checked_last = false;
    double zsel = 0.0;
    while (iter < S.max_iter && __any_sync(kFull, live)) {
      int stop = S.max_iter;
      if (ct) { const int nxt = (iter / ct + 1) * ct; stop = nxt < stop ? nxt : stop; }
      if (ai) { const int nxt = (iter / ai + 1) * ai; stop = nxt < stop ? nxt : stop; }
#pragma unroll 1
      for (; iter < stop - 1; ++iter) { admm_step<N>(c, L, rho, rho_eq, rinv, sigma, alpha, zsel); zsel = 1.0; }
      {  // keep the iterate before the last step of the chunk: delta_x, delta_y for the infeasibility tests
        double *cd = c.cold;
#pragma unroll
        for (int k = 0; k <= N; ++k) { cd[Cold<N>::PREV_X + k] = L.X[k]; cd[Cold<N>::PREV_YD + k] = L.YD[k]; }
#pragma unroll
        for (int k = 0; k < N; ++k) { cd[Cold<N>::PREV_YI + 2 * k] = L.YI[k][0]; cd[Cold<N>::PREV_YI + 2 * k + 1] = L.YI[k][1]; }
      }
      admm_step<N>(c, L, rho, rho_eq, rinv, sigma, alpha, zsel); zsel = 1.0;
      ++iter;
      const bool can_check = ct && (iter % ct == 0);
      const bool can_adapt = ai && (iter % ai == 0);
      checked_last = can_check;
      if (can_check || can_adapt) {
        Lane<N> cp = L;
        update_info<N>(c, &cp, &I);
        if (live) iter_done = iter;
        if (can_check) {
          if (check_termination<N>(c, S, &cp, &I, live, 0)) {
            live = false;  // freeze: the group keeps stepping with its warp, its iterate is parked
            double *cd = c.cold + Cold<N>::SNAP;
#pragma unroll
            for (int k = 0; k <= N; ++k) { cd[k] = L.X[k]; cd[(N + 1) + k] = L.YD[k]; }
#pragma unroll
            for (int k = 0; k < N; ++k) {
              cd[2 * (N + 1) + 2 * k] = L.ZI[k][0]; cd[2 * (N + 1) + 2 * k + 1] = L.ZI[k][1];
              cd[2 * (N + 1) + 2 * N + 2 * k] = L.YI[k][0]; cd[2 * (N + 1) + 2 * N + 2 * k + 1] = L.YI[k][1];
            }
          }
        }
        if (can_adapt) {
          const double pr = I.n_rp / ((I.n_z > I.n_Ax ? I.n_z : I.n_Ax) + 1e-10);
          double dn = I.n_q; dn = (I.n_Aty > dn) ? I.n_Aty : dn; dn = (I.n_Px > dn) ? I.n_Px : dn;
          const double dr = I.n_rd / (dn + 1e-10);
          double rho_new = rho * sqrt(pr / (dr + 1e-10));
          rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
          const bool upd = live && ((rho_new > rho * S.adaptive_rho_tolerance) || (rho_new < rho / S.adaptive_rho_tolerance));
          if (__any_sync(kFull, upd)) {
            if (upd) { rho = rho_new; rho_eq = kRhoEqOverIneq * rho; rinv = 1.0 / rho; ++rho_updates; }
            FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0; fw.act_d = 0; fw.act_i = 0;
            __syncwarp();
            factor<N>(c, fw, sigma);
          }
        }
      }
    }
    // parked iterates of the groups that stopped before their warp did
    const bool parked = !live && !sched_err && I.status != LPVMPC_UNSOLVED;
    if (parked) {
      const double *cd = c.cold + Cold<N>::SNAP;
#pragma unroll
      for (int k = 0; k <= N; ++k) { L.X[k] = cd[k]; L.YD[k] = cd[(N + 1) + k]; }
#pragma unroll
      for (int k = 0; k < N; ++k) {
        L.ZI[k][0] = cd[2 * (N + 1) + 2 * k]; L.ZI[k][1] = cd[2 * (N + 1) + 2 * k + 1];
        L.YI[k][0] = cd[2 * (N + 1) + 2 * N + 2 * k]; L.YI[k][1] = cd[2 * (N + 1) + 2 * N + 2 * k + 1];
      }
    }
    Lane<N> cp = L;
    if (!checked_last && __any_sync(kFull, live)) {
      Info J = I;
      update_info<N>(c, &cp, &J);
      if (live) { I = J; iter_done = iter; }
      if (check_termination<N>(c, S, &cp, &I, live, 0)) live = false;
    }
    {
      const bool unsolved = (I.status == LPVMPC_UNSOLVED);
      if (__any_sync(kFull, unsolved)) {
        if (!check_termination<N>(c, S, &cp, &I, unsolved, 1) && unsolved) I.status = LPVMPC_MAX_ITER_REACHED;
      }
    }
    const int status = I.status;
    const bool has_sol = !(status == LPVMPC_PRIMAL_INFEASIBLE || status == LPVMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_DUAL_INFEASIBLE || status == LPVMPC_DUAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_NON_CVX || status == LPVMPC_SCHEDULE_ERROR);
    {
      const double o = objective<N>(c, cp.Q, cp.X, S.scaling ? I.cinv : 1.0);
      if (has_sol) I.obj = o;
    }
    // row / variable indices in the reference order
    auto ref_dyn = [&](int k) { return 6 * N + k * NX + r; };
    auto ref_in = [&](int k, int t) { return (r == 0) ? (2 * k + t) : (2 * N + 4 * k + 2 * ucomp + t); };
    auto ref_var = [&](int k) { return c.xl ? (k * NX + r) : (nx + k * 2 + ucomp); };
    if (valid && (a.xs || a.zs || a.ys)) {
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        if ((c.xl || k < N) && a.xs) a.xs[(size_t)b * nz + ref_var(k)] = cp.X[k];
        if (c.xl) {
          if (a.zs) a.zs[(size_t)b * m + ref_dyn(k)] = (iter > 0) ? cp.BE[k] : 0.0;
          if (a.ys) a.ys[(size_t)b * m + ref_dyn(k)] = cp.YD[k];
        }
        if (k < N && c.il) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (a.zs) a.zs[(size_t)b * m + ref_in(k, t)] = cp.ZI[k][t];
            if (a.ys) a.ys[(size_t)b * m + ref_in(k, t)] = cp.YI[k][t];
          }
        }
      }
    }
    int polish_status = 0;
    unsigned acts[4] = {0, 0, 0, 0};
    const bool do_pol = S.polish && status == LPVMPC_SOLVED;
    if (__any_sync(kFull, do_pol)) {
      polish_status = polish<N>(c, S, &cp, &I, do_pol, acts);
      if (!do_pol) { acts[0] = acts[1] = acts[2] = acts[3] = 0; }
    }
    // ---- outputs
    if (valid) {
      const double *cd = c.cold;
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        const double v = has_sol ? cd[Cold<N>::D + k] * cp.X[k] : nan("");
        if (c.xl) a.x_pred[(size_t)b * nx + k * NX + r] = v;
        else if (k < N) a.u_pred[(size_t)b * 2 * N + k * 2 + ucomp] = v;
        if (c.xl) {
          const size_t o = (size_t)b * m + ref_dyn(k);
          if (a.y) a.y[o] = has_sol ? I.cinv * (cd[Cold<N>::ED + k] * cp.YD[k]) : nan("");
          if (a.active_lo) a.active_lo[o] = (acts[0] >> k) & 1u;
          if (a.active_up) a.active_up[o] = (acts[1] >> k) & 1u;
        }
        if (k < N && c.il) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const size_t o = (size_t)b * m + ref_in(k, t);
            if (a.y) a.y[o] = has_sol ? I.cinv * (cd[Cold<N>::EI + 2 * k + t] * cp.YI[k][t]) : nan("");
            if (a.active_lo) a.active_lo[o] = (acts[2] >> (2 * k + t)) & 1u;
            if (a.active_up) a.active_up[o] = (acts[3] >> (2 * k + t)) & 1u;
          }
        }
      }
      if (r == 0) {
        a.status[b] = status;
        if (a.iters) a.iters[b] = iter_done;
        if (a.rho_updates) a.rho_updates[b] = rho_updates;
        if (a.polish_status) a.polish_status[b] = polish_status;
        if (a.obj) a.obj[b] = I.obj;
        if (a.pri_res) a.pri_res[b] = sched_err ? nan("") : I.pri_res;
        if (a.dua_res) a.dua_res[b] = sched_err ? nan("") : I.dua_res;
      }
    }
    __syncwarp();
  }
}

}  // namespace t8
}  // namespace lpv
