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

#include "lpv_qp.cuh"

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
  static constexpr int TOTAL = (UI + N * 6 + 1) & ~1;
};
// cold per-lane data in the global scratch slab (doubles per lane)
template <int N>
struct Cold {
  static constexpr int D = 0, DINV = D + (N + 1), ED = DINV + (N + 1), EDINV = ED + (N + 1), EI = EDINV + (N + 1),
                       EIINV = EI + 2 * N, PD = EIINV + 2 * N, PO = PD + (N + 1), TOTAL = PO + N;
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
  unsigned loose; // bit (2k+t): inequality row is "loose" (both bounds infinite) -> rho_min
  __device__ __forceinline__ double *Kb(int k) const { return S + Reg<N>::K + (k - 1) * NB * LDB; }
  __device__ __forceinline__ double *Tb(int k) const { return S + Reg<N>::T + k * NB * LDB; }
  __device__ __forceinline__ double *Gt(int k) const { return S + Reg<N>::G + k * NB * NX; }
  __device__ __forceinline__ double &ed(int k) const { return S[Reg<N>::ED + k * NX + (xl ? r : 0)]; }
  __device__ __forceinline__ double &si(int k, int t) const { return S[Reg<N>::SI + (k * 3 + slot) * 2 + t]; }
  __device__ __forceinline__ double &ui(int k, int t) const { return S[Reg<N>::UI + (k * 3 + slot) * 2 + t]; }
};

// ---------------------------------------------------------------- block factorisation (lane r = row r)
// wd(k): weight of my dynamics row (k, r) (x lanes); wi(k, t): weight of my inequality row.
template <int N, class WD, class WI>
__device__ __forceinline__ void factor(const Ctx<N> &c, const double (&PD)[N + 1], const double (&PO)[N], double sigma,
                                       WD wd, WI wi) {
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
        d = fma(wi(k, 0) * a0, a0, d);
        d = fma(wi(k, 1) * a1, a1, d);
      }
      if (c.xl) { const double e = c.ed(k); d = fma(wd(k) * e, e, d); }
      if (!rowlive) d = 1.0;
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) if (cc == r) s[cc] = d;
    }
    // next-stage dynamics rows: sum_r' w(k+1,r') G[r'][r] G[r'][c]
    if (k < N) {
      const double *g = c.Gt(k);
      const double wme = c.xl ? wd(k + 1) : 0.0;
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
        const double f = wd(k) * c.ed(k);
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
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) gv[cc] = (cc < nbk) ? gshfl(v, cc) : 0.0;
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
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) gx[rr] = (rr < nbn) ? gshfl(x[k + 1], rr) : 0.0;
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
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) gx[rr] = gshfl(x[0], rr);
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
#pragma unroll
      for (int rr = 0; rr < NX; rr += 2) {
        acc = fma(g[rr], gshfl(td[k + 1], rr), acc);
        a1 = fma(g[rr + 1], gshfl(td[k + 1], rr + 1), a1);
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

// ---------------------------------------------------------------- one persistent warp = 4 QPs at a time
template <int N>
__global__ void __launch_bounds__(32, 4) lpv_solve_t8_kernel(const __grid_constant__ Params p, unsigned *queue, double *cold_slab) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int g = lane >> 3, r = lane & 7;
  Ctx<N> c;
  c.S = smem + g * Reg<N>::TOTAL;
  c.cold = cold_slab + ((size_t)(blockIdx.x * 4 + g) * 8 + r) * Cold<N>::TOTAL;
  c.r = r; c.xl = r < NX; c.il = (r == 0) || (r >= NX); c.slot = (r == 0) ? 0 : ((r >= NX) ? r - 5 : 0);
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &S = p.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N, m = 6 * N + nx;
  const int ucomp = r - NX;  // input component on u lanes

  for (;;) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(queue, 4u);
    base = __shfl_sync(kFull, base, 0);
    if ((int)base >= p.B) break;
    const bool valid = (int)(base + g) < p.B;
    const int b = valid ? (int)(base + g) : (int)base;  // idle groups shadow group 0 and never write

    Lane<N> L;
    double PD[N + 1], PO[N], D[N + 1], Ed[N + 1], Ei[N][2];
    int sched_err = 0;
    // ================================================================ schedule: Gt_k = -[A_k B_k] (unscaled)
    double x0r = 0.0;
    {
      if (a.sched_mode == LPVMPC_SCHED_GIVEN) {
#pragma unroll
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
          // my row (x lanes) of -[A B]; selected without dynamic register indexing
          double row[NB];
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
    }
    sched_err = gany(sched_err);

    // ================================================================ build (unscaled data)
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

    // ================================================================ Ruiz equilibration
    double csc = 1.0;
    if (S.scaling) {
      double *scr = c.S + Reg<N>::K;  // factor area is free during setup: column norms in reference order
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
    const double cinv = 1.0 / csc;
    // scaled bounds, loose flags, cold data
    c.loose = 0;
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
          if (c.il) {
            const double u = Ei[k][t] * c.ui(k, t);
            c.ui(k, t) = u;
            if (u > kInfty * kMinScaling) c.loose |= 1u << (2 * k + t);  // l = -inf always for these rows
          }
        }
      }
      __syncwarp();
    }
    const bool unscale = S.scaling && !S.scaled_termination;
    const double sigma = S.sigma, alpha = S.alpha;
    double rho = fmin(fmax(S.rho, kRhoMin), kRhoMax);
    double rho_eq = kRhoEqOverIneq * rho, rinv = 1.0 / rho, rinv_eq = 1.0 / rho_eq;
    const double rinv_loose = 1.0 / kRhoMin;
    const unsigned loose = c.loose;
    auto rho_i = [&](int k, int t) { return ((loose >> (2 * k + t)) & 1u) ? kRhoMin : rho; };
    auto rinv_i = [&](int k, int t) { return ((loose >> (2 * k + t)) & 1u) ? rinv_loose : rinv; };

    factor<N>(c, PD, PO, sigma, [&](int) { return rho_eq; }, [&](int k, int t) { return rho_i(k, t); });

    // ================================================================ ADMM
    int status = sched_err ? LPVMPC_SCHEDULE_ERROR : LPVMPC_UNSOLVED;
    bool live = !sched_err;
    int iter_done = 0, rho_updates = 0;
    double pri_res = 0.0, dua_res = 0.0, obj = nan("");
    double n_rp = 0, n_z = 0, n_Ax = 0, n_rd = 0, n_q = 0, n_Aty = 0, n_Px = 0, u_z = 0, u_Ax = 0, u_q = 0, u_Aty = 0, u_Px = 0;
    int adapt_interval = S.adaptive_rho_interval;
    if (S.adaptive_rho && !adapt_interval) adapt_interval = S.check_termination ? 4 * S.check_termination : 100;
    double Xp[N + 1], DYD[N + 1], DYI[N][2];
#pragma unroll
    for (int k = 0; k <= N; ++k) { Xp[k] = 0.0; DYD[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < N; ++k) { DYI[k][0] = DYI[k][1] = 0.0; }
    bool zinit = true;  // z_dyn is still the cold-start zero (afterwards it equals the equality bound)

    // residual norms at the current iterate (update_info)
    auto update_info = [&]() {
      const double *cd = c.cold;
      double Ax[N + 1], Px[N + 1], Aty[N + 1], tdv[N + 1], tiv[N][2];
      rowsA<N>(c, L.X, Ax);
      double a_rp = 0, a_z = 0, a_Ax = 0, b_rp = 0, b_z = 0, b_Ax = 0;
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        if (c.xl) {
          const double z = zinit ? 0.0 : L.BE[k], rr = Ax[k] - z, ei = cd[Cold<N>::EDINV + k];
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
      n_rp = gmax(a_rp); n_z = gmax(a_z); n_Ax = gmax(a_Ax); n_rd = gmax(a_rd); n_q = gmax(a_q); n_Aty = gmax(a_Aty); n_Px = gmax(a_Px);
      if (unscale) {
        pri_res = gmax(b_rp); u_z = gmax(b_z); u_Ax = gmax(b_Ax);
        dua_res = cinv * gmax(b_rd); u_q = gmax(b_q); u_Aty = gmax(b_Aty); u_Px = gmax(b_Px);
      } else {
        pri_res = n_rp; u_z = n_z; u_Ax = n_Ax; dua_res = n_rd; u_q = n_q; u_Aty = n_Aty; u_Px = n_Px;
      }
    };

    auto primal_infeasible = [&](double eps) -> bool {
      const double *cd = c.cold;
      double nrm = 0.0;
#pragma unroll
      for (int k = 0; k <= N; ++k) if (c.xl) nrm = absmax(nrm, unscale ? cd[Cold<N>::ED + k] * DYD[k] : DYD[k]);  // equality rows: both bounds finite
#pragma unroll
      for (int k = 0; k < N; ++k)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (c.il) {
            double d = DYI[k][t];
            if ((loose >> (2 * k + t)) & 1u) d = 0.0; else d = (d > 0.0) ? d : 0.0;  // l = -inf
            DYI[k][t] = d;
            nrm = absmax(nrm, unscale ? cd[Cold<N>::EI + 2 * k + t] * d : d);
          } else DYI[k][t] = 0.0;
        }
      nrm = gmax(nrm);
      bool res = false;
      if (__any_sync(kFull, nrm > eps)) {
        double lhs = 0.0;
#pragma unroll
        for (int k = 0; k <= N; ++k) if (c.xl) { const double d = DYD[k]; lhs += L.BE[k] * ((d > 0) ? d : 0) + L.BE[k] * ((d < 0) ? d : 0); }
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int t = 0; t < 2; ++t) if (c.il) { const double d = DYI[k][t]; lhs += c.ui(k, t) * ((d > 0) ? d : 0) + (-kInfty * cd[Cold<N>::EI + 2 * k + t]) * ((d < 0) ? d : 0); }
        lhs = gsum(lhs);
        double tdv[N + 1], at[N + 1];
#pragma unroll
        for (int k = 0; k <= N; ++k) tdv[k] = c.xl ? DYD[k] : 0.0;
        colsA<N>(c, tdv, DYI, at);
        double mx = 0.0;
#pragma unroll
        for (int k = 0; k <= N; ++k) if (c.xl || k < N) mx = absmax(mx, unscale ? cd[Cold<N>::DINV + k] * at[k] : at[k]);
        mx = gmax(mx);
        res = (nrm > eps) && (lhs < -eps * nrm) && (mx < eps * nrm);
      }
      return res;
    };

    auto dual_infeasible = [&](double eps) -> bool {
      const double *cd = c.cold;
      double dx[N + 1], nrm = 0.0, qdx = 0.0;
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        dx[k] = (c.xl || k < N) ? L.X[k] - Xp[k] : 0.0;
        nrm = absmax(nrm, unscale ? cd[Cold<N>::D + k] * dx[k] : dx[k]);
        qdx += L.Q[k] * dx[k];
      }
      nrm = gmax(nrm); qdx = gsum(qdx);
      const double cs = unscale ? csc : 1.0;
      bool res = false;
      if (__any_sync(kFull, (nrm > eps) && (qdx < -cs * eps * nrm))) {
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
            if ((c.ui(k, t) < kInfty * kMinScaling) && (v > eps * nrm)) viol = 1;
          }
        res = (nrm > eps) && (qdx < -cs * eps * nrm) && (mx < cs * eps * nrm) && !gany(viol);
      }
      return res;
    };

    auto check_termination = [&](int approximate) -> int {
      double eps_abs = S.eps_abs, eps_rel = S.eps_rel, eps_pi = S.eps_prim_inf, eps_di = S.eps_dual_inf;
      const bool ncvx = (pri_res > kInfty) || (dua_res > kInfty);
      if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; eps_di *= 10; }
      const double eps_prim = eps_abs + eps_rel * (u_z > u_Ax ? u_z : u_Ax);
      const bool prim_ok = pri_res < eps_prim;
      double mr = u_q; mr = (u_Aty > mr) ? u_Aty : mr; mr = (u_Px > mr) ? u_Px : mr;
      if (unscale) mr *= cinv;
      const double eps_dual = eps_abs + eps_rel * mr;
      const bool dual_ok = dua_res < eps_dual;
      bool prim_inf = false, dual_inf = false;
      if (__any_sync(kFull, live && !ncvx && !prim_ok)) prim_inf = primal_infeasible(eps_pi) && !prim_ok;
      if (__any_sync(kFull, live && !ncvx && !dual_ok)) dual_inf = dual_infeasible(eps_di) && !dual_ok;
      if (!live) return 0;
      if (ncvx) { status = LPVMPC_NON_CVX; obj = nan(""); return 1; }
      if (prim_ok && dual_ok) { status = approximate ? LPVMPC_SOLVED_INACCURATE : LPVMPC_SOLVED; return 1; }
      if (prim_inf) { status = approximate ? LPVMPC_PRIMAL_INFEASIBLE_INACCURATE : LPVMPC_PRIMAL_INFEASIBLE; obj = kInfty; return 1; }
      if (dual_inf) { status = approximate ? LPVMPC_DUAL_INFEASIBLE_INACCURATE : LPVMPC_DUAL_INFEASIBLE; obj = -kInfty; return 1; }
      return 0;
    };

    bool checked_last = false;
    int iter;
#pragma unroll 1
    for (iter = 1; iter <= S.max_iter; ++iter) {
      if (!__any_sync(kFull, live)) break;
      const bool can_check = S.check_termination && (iter % S.check_termination == 0);
      const bool can_adapt = S.adaptive_rho && adapt_interval && (iter % adapt_interval == 0);
      const bool keep = can_check || can_adapt || iter == S.max_iter;
      // right-hand side
      double bv[N + 1], xt[N + 1], zt[N + 1];
      {
        double tdv[N + 1], tiv[N][2], at[N + 1];
#pragma unroll
        for (int k = 0; k <= N; ++k) tdv[k] = c.xl ? (rho_eq * (zinit ? 0.0 : L.BE[k]) - L.YD[k]) : 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) { tiv[k][0] = rho_i(k, 0) * L.ZI[k][0] - L.YI[k][0]; tiv[k][1] = rho_i(k, 1) * L.ZI[k][1] - L.YI[k][1]; }
        colsA<N>(c, tdv, tiv, at);
#pragma unroll
        for (int k = 0; k <= N; ++k) bv[k] = (sigma * L.X[k] - L.Q[k]) + at[k];
      }
      solve<N>(c, bv, xt, zt);
      // z, y, x updates
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        if (c.xl) {
          const double zr = alpha * zt[k] + (1.0 - alpha) * (zinit ? 0.0 : L.BE[k]);
          const double d = rho_eq * (zr - L.BE[k]);
          if (live) { L.YD[k] += d; if (keep) DYD[k] = d; }
        }
        if (k < N && c.il) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const double zti = c.si(k, t) * xt[k];
            const double zr = alpha * zti + (1.0 - alpha) * L.ZI[k][t];
            const double zc = zr + rinv_i(k, t) * L.YI[k][t];
            // l = -1e30 * E_i (<= -1e10) cannot bind on a finite iterate: the projection is min(., u)
            const double up = c.ui(k, t);
            const double zn = (zc < up) ? zc : up;
            const double d = rho_i(k, t) * (zr - zn);
            if (live) { L.YI[k][t] += d; L.ZI[k][t] = zn; if (keep) DYI[k][t] = d; }
          }
        }
        if (live && (c.xl || k < N)) {
          const double xo = L.X[k];
          if (keep) Xp[k] = xo;
          L.X[k] = alpha * xt[k] + (1.0 - alpha) * xo;
        }
      }
      zinit = false;
      checked_last = can_check;
      if (can_check) {
        update_info();
        if (live) iter_done = iter;
        if (check_termination(0)) live = false;
      }
      if (can_adapt) {
        if (!can_check) { update_info(); if (live) iter_done = iter; }
        const double pr = n_rp / ((n_z > n_Ax ? n_z : n_Ax) + 1e-10);
        double dn = n_q; dn = (n_Aty > dn) ? n_Aty : dn; dn = (n_Px > dn) ? n_Px : dn;
        const double dr = n_rd / (dn + 1e-10);
        double rho_new = rho * sqrt(pr / (dr + 1e-10));
        rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
        const bool upd = live && ((rho_new > rho * S.adaptive_rho_tolerance) || (rho_new < rho / S.adaptive_rho_tolerance));
        if (__any_sync(kFull, upd)) {
          if (upd) { rho = rho_new; rho_eq = kRhoEqOverIneq * rho; rinv = 1.0 / rho; rinv_eq = 1.0 / rho_eq; ++rho_updates; }
          double PDc[N + 1], POc[N];
#pragma unroll
          for (int k = 0; k <= N; ++k) PDc[k] = c.cold[Cold<N>::PD + k];
#pragma unroll
          for (int k = 0; k < N; ++k) POc[k] = c.cold[Cold<N>::PO + k];
          __syncwarp();
          factor<N>(c, PDc, POc, sigma, [&](int) { return rho_eq; }, [&](int k, int t) { return rho_i(k, t); });
        }
      }
    }
    (void)rinv_eq;
    if (!checked_last) {
      const bool was_live = live;
      update_info();
      if (was_live) iter_done = iter - 1;
      if (check_termination(0)) live = false;
    }
    {
      const bool unsolved = (status == LPVMPC_UNSOLVED);
      const bool was_live = live;
      live = unsolved;
      if (__any_sync(kFull, unsolved)) { if (!check_termination(1) && unsolved) status = LPVMPC_MAX_ITER_REACHED; }
      live = was_live && false;
    }
    const bool has_sol = !(status == LPVMPC_PRIMAL_INFEASIBLE || status == LPVMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_DUAL_INFEASIBLE || status == LPVMPC_DUAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_NON_CVX || status == LPVMPC_SCHEDULE_ERROR);
    double PDc[N + 1], POc[N];
#pragma unroll
    for (int k = 0; k <= N; ++k) PDc[k] = c.cold[Cold<N>::PD + k];
#pragma unroll
    for (int k = 0; k < N; ++k) POc[k] = c.cold[Cold<N>::PO + k];
    auto objective = [&](const double (&xv)[N + 1]) -> double {
      double Px[N + 1], acc = 0.0;
      rowsP<N>(c, PDc, POc, xv, Px);
#pragma unroll
      for (int k = 0; k <= N; ++k) if (c.xl || k < N) acc += (0.5 * Px[k] + L.Q[k]) * xv[k];
      return gsum(acc) * (S.scaling ? cinv : 1.0);
    };
    if (has_sol) obj = objective(L.X);

    // row / variable indices in the reference order
    auto ref_dyn = [&](int k) { return 6 * N + k * NX + r; };
    auto ref_in = [&](int k, int t) { return (r == 0) ? (2 * k + t) : (2 * N + 4 * k + 2 * ucomp + t); };
    auto ref_var = [&](int k) { return c.xl ? (k * NX + r) : (nx + k * 2 + ucomp); };
    if (valid && (a.xs || a.zs || a.ys)) {
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        if ((c.xl || k < N) && a.xs) a.xs[(size_t)b * nz + ref_var(k)] = L.X[k];
        if (c.xl) {
          if (a.zs) a.zs[(size_t)b * m + ref_dyn(k)] = zinit ? 0.0 : L.BE[k];
          if (a.ys) a.ys[(size_t)b * m + ref_dyn(k)] = L.YD[k];
        }
        if (k < N && c.il) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (a.zs) a.zs[(size_t)b * m + ref_in(k, t)] = L.ZI[k][t];
            if (a.ys) a.ys[(size_t)b * m + ref_in(k, t)] = L.YI[k][t];
          }
        }
      }
    }

    // ================================================================ polish
    int polish_status = 0;
    unsigned act_lo_d = 0, act_up_d = 0, act_lo_i = 0, act_up_i = 0;  // bit k (dyn) / bit 2k+t (ineq)
    const bool do_pol = S.polish && status == LPVMPC_SOLVED;
    if (__any_sync(kFull, do_pol)) {
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        if (c.xl) {  // equality row: z == l == u
          const double z = zinit ? 0.0 : L.BE[k];
          if (z - L.BE[k] < -L.YD[k]) act_lo_d |= 1u << k;
          if (L.BE[k] - z < L.YD[k]) act_up_d |= 1u << k;
        }
        if (k < N && c.il) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const double lo = -kInfty * c.cold[Cold<N>::EI + 2 * k + t];
            if (L.ZI[k][t] - lo < -L.YI[k][t]) act_lo_i |= 1u << (2 * k + t);
            if (c.ui(k, t) - L.ZI[k][t] < L.YI[k][t]) act_up_i |= 1u << (2 * k + t);
          }
        }
      }
      const unsigned act_d = act_lo_d | act_up_d, act_i = act_lo_i | act_up_i;
      const double delta = S.delta, idel = 1.0 / S.delta;
      __syncwarp();
      factor<N>(c, PDc, POc, delta, [&](int k) { return ((act_d >> k) & 1u) ? idel : 0.0; },
                [&](int k, int t) { return ((act_i >> (2 * k + t)) & 1u) ? idel : 0.0; });
      auto bred_i = [&](int k, int t) { return ((act_lo_i >> (2 * k + t)) & 1u) ? (-kInfty * c.cold[Cold<N>::EI + 2 * k + t]) : c.ui(k, t); };
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
      const double *cd = c.cold;
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
      const double pol_pri = gmax(a_rp), pol_dua = (unscale ? cinv : 1.0) * gmax(a_rd);
      const double pol_obj = objective(px);
      const bool ok = (pol_pri < pri_res && pol_dua < dua_res) || (pol_pri < pri_res && dua_res < 1e-10) ||
                      (pol_dua < dua_res && pri_res < 1e-10);
      if (do_pol) {
        if (ok) {
          obj = pol_obj; pri_res = pol_pri; dua_res = pol_dua; polish_status = 1;
#pragma unroll
          for (int k = 0; k <= N; ++k) { L.X[k] = px[k]; L.YD[k] = pyd[k]; }
#pragma unroll
          for (int k = 0; k < N; ++k) { L.ZI[k][0] = pzi[k][0]; L.ZI[k][1] = pzi[k][1]; L.YI[k][0] = pyi[k][0]; L.YI[k][1] = pyi[k][1]; }
        } else polish_status = -1;
      }
    }

    // ================================================================ outputs
    if (valid) {
      const double *cd = c.cold;
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        const double v = has_sol ? cd[Cold<N>::D + k] * L.X[k] : nan("");
        if (c.xl) a.x_pred[(size_t)b * nx + k * NX + r] = v;
        else if (k < N) a.u_pred[(size_t)b * 2 * N + k * 2 + ucomp] = v;
        if (c.xl) {
          const size_t o = (size_t)b * m + ref_dyn(k);
          if (a.y) a.y[o] = has_sol ? cinv * (cd[Cold<N>::ED + k] * L.YD[k]) : nan("");
          if (a.active_lo) a.active_lo[o] = do_pol ? ((act_lo_d >> k) & 1u) : 0;
          if (a.active_up) a.active_up[o] = do_pol ? ((act_up_d >> k) & 1u) : 0;
        }
        if (k < N && c.il) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const size_t o = (size_t)b * m + ref_in(k, t);
            if (a.y) a.y[o] = has_sol ? cinv * (cd[Cold<N>::EI + 2 * k + t] * L.YI[k][t]) : nan("");
            if (a.active_lo) a.active_lo[o] = do_pol ? ((act_lo_i >> (2 * k + t)) & 1u) : 0;
            if (a.active_up) a.active_up[o] = do_pol ? ((act_up_i >> (2 * k + t)) & 1u) : 0;
          }
        }
      }
      if (r == 0) {
        a.status[b] = status;
        if (a.iters) a.iters[b] = iter_done;
        if (a.rho_updates) a.rho_updates[b] = rho_updates;
        if (a.polish_status) a.polish_status[b] = polish_status;
        if (a.obj) a.obj[b] = obj;
        if (a.pri_res) a.pri_res[b] = sched_err ? nan("") : pri_res;
        if (a.dua_res) a.dua_res[b] = sched_err ? nan("") : dua_res;
      }
    }
    __syncwarp();
  }
}

}  // namespace t8
}  // namespace lpv
