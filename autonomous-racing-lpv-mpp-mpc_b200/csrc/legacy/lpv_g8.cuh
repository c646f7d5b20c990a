// "G8" kernel: 8 lanes per QP, QPW QPs per warp, compact code (run-time loops over the stages), all per-QP state
// in shared memory.
//
// Why (profiles/r1a_ncu_t8_ctrl4096.txt, profiles/r1_icache_probe.jsonl): the fully unrolled T8 kernel is bound by
// instruction fetch — one warp per SM sub-partition issues 0.8 IPC from a loop body of <= 8 KB but only 0.31 / 0.17 IPC
// from 32 KB / > 64 KB bodies, and T8's ADMM step is 32 KB, its cold code 1.2 MB.  Here every loop body is a few
// hundred instructions: the ADMM iteration is two sweeps over the stages,
//     forward :  gather v_k ;  W_k = T_k v_k ;  v_{k+1} = b_{k+1} - K_{k+1} v_k
//     backward:  x~_k = W_k - K_{k+1}' x~_{k+1} ; gather x~_k ; then everything that only needs stage k+1 and x~_k:
//                z~ of the dynamics rows, the z / y / x updates (relaxation, projection) and the NEXT right-hand side
//                b = sigma x - q + A'(rho z - y), so no other pass over the data is needed.
// Lane r of a group owns component r of every stage variable w_k = [x_k; u_k], the dynamics row (k, r) and the
// single-variable rows on its variable.  8x8 blocks are stored row-major with the 16-byte chunks of row rr XOR-swizzled
// by (rr >> 1): row reads (LDS.128, one row per lane) and column reads (LDS.64) are both bank-conflict free without
// padding.  Rarely used data (scalings, P, previous iterate, polish vectors) sits in an L2-resident scratch slab.
//
// Same algorithm as lpv_qp.cuh / lpv_t8.cuh / oracle/osqp_ref.c (OSQP 0.6: Ruiz scaling, rho vector, relaxed ADMM,
// termination + infeasibility certificates every check_termination iterations, adaptive rho, polish); restricted to
// diagonal Q and R (the reference's tunings, controllerMain.py:139-148, plannerMain.py:96-99) and steering_delay = 0.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../lpv_qp.cuh"

namespace lpv {
namespace g8 {

struct Lay {  // per-QP shared-memory offsets (doubles); computed on the host (lpvmpc.cu: make_g8_layout)
  int N, gs;  // horizon; doubles per G block (NX * 8)
  int T, K, G;                       // (N+1) x 64, N x 64 (K_k at k-1), N x gs
  int X, Q, B, YD, BE, ED;           // (N+1) x 8: iterate x, linear cost, rhs / sweep vector, dynamics-row dual, rhs, identity coef
  int ZI, YI, SI, UI, LI;            // (N+1) x 8: single-variable rows (z, y, coefficient, bounds)
  int total;
  int cold_total;                    // doubles per QP slot in the global scratch slab
};
enum { C_D = 0, C_DINV, C_E, C_EINV, C_PD, C_PO, C_EI, C_EIINV, C_PVX, C_PVYD, C_PVYI, C_PX, C_PYD, C_PYI, C_R2D, C_R2I,
       C_ACTD, C_ACTI, C_COUNT };

struct G8Params {
  Lay L;
  Model M;
  lpvmpc_settings S;
  lpvmpc_args a;
  int B;
  unsigned *queue;
  double *cold;
};

template <int KIND> struct Dims;
template <> struct Dims<LPVMPC_CONTROLLER> { static constexpr int NX = 6, NT = 2; };
template <> struct Dims<LPVMPC_PLANNER> { static constexpr int NX = 5, NT = 1; };

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
__device__ __forceinline__ int gany(int v) {
  const unsigned m = __ballot_sync(kFull, v);
  return ((m >> ((threadIdx.x & 31) & ~7)) & 0xffu) != 0;
}
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
__device__ __forceinline__ int swz(int rr) { return (rr >> 1) & 3; }
// offset of logical 16-byte chunk j of row rr inside a swizzled 8-wide block
__device__ __forceinline__ int chunk(int rr, int j) { return rr * 8 + ((j ^ swz(rr)) << 1); }

template <int KIND>
struct Ctx {
  static constexpr int NX = Dims<KIND>::NX, NB = NX + 2, NT = Dims<KIND>::NT;
  double *S;     // my QP's shared region
  double *cold;  // my QP slot in the global scratch slab
  const Lay *L;
  int N, r;
  int ro[4];     // my row: offset of logical chunk j
  int co[4];     // my column: offset inside row rr is co[rr >> 1]
  bool xl, ul;   // state lane / input lane (neither: idle lane)
  int islot;     // first single-variable-row slot of my variable inside a stage (t adds 1)
  uint64_t eqm, loosem;  // planner: bit k = my box row at stage k is an equality / both bounds infinite

  __device__ __forceinline__ bool var_live(int k) const { return xl || (ul && k < N); }
  __device__ __forceinline__ bool has_in(int k) const {
    if (KIND == LPVMPC_CONTROLLER) return (r == 0 || ul) && k < N;
    return xl || (ul && k < N);
  }
  __device__ __forceinline__ double *cd(int arr) const { return cold + arr * (N + 1) * 8; }
  __device__ __forceinline__ double *Tb(int k) const { return S + L->T + k * 64; }
  __device__ __forceinline__ double *Kb(int k) const { return S + L->K + (k - 1) * 64; }
  __device__ __forceinline__ double *Gb(int k) const { return S + L->G + k * L->gs; }
  // row-weight classes of my single-variable row (planner); controller rows are always plain inequalities
  __device__ __forceinline__ double rho_of(int k, double rho, double rho_eq) const {
    if (KIND == LPVMPC_CONTROLLER) return rho;
    return ((eqm >> k) & 1ull) ? rho_eq : (((loosem >> k) & 1ull) ? kRhoMin : rho);
  }
};

// dot product of my row (logical chunks at ro[]) of block `blk` with the gathered vector g
__device__ __forceinline__ double rowdot(const double *blk, const int (&ro)[4], const double (&g)[8]) {
  const double2 t0 = ld2(blk + ro[0]), t1 = ld2(blk + ro[1]), t2 = ld2(blk + ro[2]), t3 = ld2(blk + ro[3]);
  double a0 = t0.x * g[0], a1 = t1.x * g[2], a2 = t2.x * g[4], a3 = t3.x * g[6];
  a0 = fma(t0.y, g[1], a0); a1 = fma(t1.y, g[3], a1); a2 = fma(t2.y, g[5], a2); a3 = fma(t3.y, g[7], a3);
  return (a0 + a1) + (a2 + a3);
}
// dot product of my column of block `blk` (rows 0..NR-1) with the gathered vector g
template <int NR>
__device__ __forceinline__ double coldot(const double *blk, const int (&co)[4], const double (&g)[8]) {
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int rr = 0; rr < NR; rr += 2) {
    a0 = fma(blk[rr * 8 + co[rr >> 1]], g[rr], a0);
    if (rr + 1 < NR) a1 = fma(blk[(rr + 1) * 8 + co[rr >> 1]], g[rr + 1], a1);
  }
  return a0 + a1;
}
__device__ __forceinline__ void gather8(double v, double (&g)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) g[c] = gshfl(v, c);
}

// ---------------------------------------------------------------- block factorisation (cold)
// Row weights: ADMM -> rho_eq on the dynamics rows, rho / rho_eq / rho_min on the single-variable rows;
// polish -> 1/delta on the active rows (C_ACTD / C_ACTI), 0 elsewhere.
struct FW { int polish; double rho, rho_eq, idel; };

template <int KIND>
__device__ __noinline__ void factor(const Ctx<KIND> c, const FW fw, const double sigma) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
  const double *ED = c.S + L.ED, *SI = c.S + L.SI;
  auto wd = [&](int k) -> double {  // weight of my dynamics row (k, r)
    if (!c.xl) return 0.0;
    return fw.polish ? ((ACTD[k * 8 + r] != 0.0) ? fw.idel : 0.0) : fw.rho_eq;
  };
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int nbk = (k < N) ? NB : NX;
    const bool rowlive = r < nbk;
    double s[8], so[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) { s[cc] = 0.0; so[cc] = 0.0; }
    const double wdk = wd(k);
    {
      double d = PD[k * 8 + r] + sigma;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const double a = SI[k * 8 + c.islot + t];
          const double w = fw.polish ? ((ACTI[k * 8 + c.islot + t] != 0.0) ? fw.idel : 0.0) : c.rho_of(k, fw.rho, fw.rho_eq);
          d = fma(w * a, a, d);
        }
      }
      if (c.xl) { const double e = ED[k * 8 + r]; d = fma(wdk * e, e, d); }
      if (!rowlive) d = 1.0;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) if (cc == r) s[cc] = d;
    }
    if (k < N) {  // next-stage dynamics rows: sum_rr w(k+1, rr) G[rr][r] G[rr][cc]
      const double *g = c.Gb(k);
      const double wn = wd(k + 1);
#pragma unroll
      for (int rr = 0; rr < NX; ++rr) {
        const double col = gshfl(wn, rr) * g[rr * 8 + c.co[rr >> 1]];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 e = ld2(g + chunk(rr, j));
          s[2 * j] = fma(col, e.x, s[2 * j]);
          s[2 * j + 1] = fma(col, e.y, s[2 * j + 1]);
        }
      }
    }
    if (k > 0) {
      if (c.xl) {  // my row of S_{k,k-1}
        const double f = wdk * ED[k * 8 + r];
        const double *gp = c.Gb(k - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(gp + c.ro[j]); so[2 * j] = f * e.x; so[2 * j + 1] = f * e.y; }
      } else if (c.ul && k < N) {
#pragma unroll
        for (int cc = NX; cc < NB; ++cc) if (cc == r) so[cc] = PO[(k - 1) * 8 + r];
      }
      double *Kk = c.Kb(k);
      const double *Tp = c.Tb(k - 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) st2(Kk + c.ro[j], so[2 * j], so[2 * j + 1]);  // park S_{k,k-1} in K_k's slot
      __syncwarp();
      double kr[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) kr[cc] = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 e = ld2(Tp + chunk(j, q));
          kr[2 * q] = fma(so[j], e.x, kr[2 * q]);
          kr[2 * q + 1] = fma(so[j], e.y, kr[2 * q + 1]);
        }
      }
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        double acc = s[cc];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 e = ld2(Kk + chunk(cc, q));  // S_{k,k-1}[cc][2q..2q+1]
          acc = fma(-kr[2 * q], e.x, acc);
          acc = fma(-kr[2 * q + 1], e.y, acc);
        }
        s[cc] = acc;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) st2(Kk + c.ro[j], kr[2 * j], kr[2 * j + 1]);
    }
    if (!rowlive) {
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) s[cc] = (cc == r) ? 1.0 : 0.0;
    }
    // Gauss-Jordan inverse of the pivot block, rows across lanes (no pivoting: the block is SPD)
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (p < nbk) {
        double pr[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) pr[cc] = gshfl(s[cc], p);
        const double piv = 1.0 / pr[p];
        if (r == p) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) s[cc] = (cc == p) ? piv : s[cc] * piv;
        } else {
          const double fp = s[p] * piv;
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) s[cc] = (cc == p) ? -fp : fma(-fp, pr[cc], s[cc]);
        }
      }
    }
    double *Tk = c.Tb(k);
#pragma unroll
    for (int j = 0; j < 4; ++j) st2(Tk + c.ro[j], s[2 * j], s[2 * j + 1]);
    __syncwarp();
  }
}

// ---------------------------------------------------------------- sweeps
// forward: B holds the right-hand side on entry, W_k = T_k v_k on exit
template <int KIND>
__device__ __forceinline__ void sweep_fwd(const Ctx<KIND> &c, const bool live) {
  const int N = c.N, r = c.r;
  double *BV = c.S + c.L->B;
  const double *Tm = c.S + c.L->T, *Km = c.S + c.L->K;
  double v = BV[r];
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    double g[8];
    gather8(v, g);
    const double w = rowdot(Tm + k * 64, c.ro, g);
    if (k < N) v = BV[(k + 1) * 8 + r] - rowdot(Km + k * 64, c.ro, g);
    if (live) BV[k * 8 + r] = w;
  }
}

// backward sweep only (polish): x_k = W_k - K_{k+1}' x_{k+1}, written back into B; also returns z~ = A x on my dynamics
// rows in ZT (cold array pointer) when zt != nullptr
template <int KIND>
__device__ __forceinline__ void sweep_bwd_plain(const Ctx<KIND> &c, double *zt) {
  constexpr int NX = Ctx<KIND>::NX;
  const int N = c.N, r = c.r;
  double *BV = c.S + c.L->B;
  const double *Km = c.S + c.L->K, *ED = c.S + c.L->ED;
  double gn[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) gn[q] = 0.0;
  double xn = 0.0;
#pragma unroll 1
  for (int k = N; k >= 0; --k) {
    double x = BV[k * 8 + r];
    if (k < N) x -= coldot<8>(Km + k * 64, c.co, gn);
    double g[8];
    gather8(x, g);
    if (zt && k < N && c.xl) zt[(k + 1) * 8 + r] = ED[(k + 1) * 8 + r] * xn + rowdot(c.Gb(k), c.ro, g);
    BV[k * 8 + r] = x;
    xn = x;
#pragma unroll
    for (int q = 0; q < 8; ++q) gn[q] = g[q];
  }
  if (zt && c.xl) zt[r] = ED[r] * xn;
}

// backward sweep fused with the ADMM updates of stage k+1 and the next right-hand side (hot)
template <int KIND>
__device__ __forceinline__ void sweep_bwd_admm(const Ctx<KIND> &c, const bool live, const double rho, const double rho_eq,
                                               const double rinv, const double rinv_eq, const double sigma,
                                               const double alpha, const double zsel) {
  constexpr int NX = Ctx<KIND>::NX, NT = Ctx<KIND>::NT;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  double *S = c.S;
  double *BV = S + L.B, *X = S + L.X, *YD = S + L.YD, *ZI = S + L.ZI, *YI = S + L.YI;
  const double *Km = S + L.K, *ED = S + L.ED, *BE = S + L.BE, *QV = S + L.Q, *SI = S + L.SI, *UI = S + L.UI, *LI = S + L.LI;
  const double oma = 1.0 - alpha, omz = oma * zsel;
  double gn[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) gn[q] = 0.0;
  double xn = 0.0, carry = 0.0;  // x~_{k+1}[r];  (G_{k+1}' td_{k+2})[r]
#pragma unroll 1
  for (int k = N; k >= -1; --k) {
    double g[8];
    double xt = 0.0;
    if (k >= 0) {
      xt = BV[k * 8 + r];
      if (k < N) xt -= coldot<8>(Km + k * 64, c.co, gn);
      gather8(xt, g);
    }
    if (k < N) {  // finish stage j = k + 1 (x~_j = xn is mine, x~_{j-1} gathered in g)
      const int j = k + 1, o = j * 8 + r;
      double td = 0.0, edtd = 0.0;
      if (c.xl) {
        const double ed = ED[o], be = BE[o];
        double zt = ed * xn;
        if (k >= 0) zt += rowdot(c.Gb(k), c.ro, g);
        const double zr = alpha * zt + omz * be;
        const double yd = YD[o] + rho_eq * (zr - be);
        if (live) YD[o] = yd;
        td = rho_eq * be - yd;
        edtd = ed * td;
      }
      double carry_new = 0.0;
      if (k >= 0) {
        const double *gk = c.Gb(k);
        double a1 = 0.0;
#pragma unroll
        for (int rr = 0; rr < NX; rr += 2) {
          carry_new = fma(gk[rr * 8 + c.co[rr >> 1]], gshfl(td, rr), carry_new);
          if (rr + 1 < NX) a1 = fma(gk[(rr + 1) * 8 + c.co[rr >> 1]], gshfl(td, rr + 1), a1);
        }
        carry_new += a1;
      }
      double acc = edtd + carry;
      if (c.has_in(j)) {
        const double rt = c.rho_of(j, rho, rho_eq);
        const double ri = (KIND == LPVMPC_CONTROLLER) ? rinv : (rt == rho ? rinv : (rt == rho_eq ? rinv_eq : 1.0 / kRhoMin));
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oi = j * 8 + c.islot + t;
          const double si = SI[oi], zi = ZI[oi], yi = YI[oi];
          const double zr = alpha * (si * xn) + oma * zi;
          const double zn = clampd(zr + ri * yi, LI[oi], UI[oi]);
          const double yn = yi + rt * (zr - zn);
          if (live) { ZI[oi] = zn; YI[oi] = yn; }
          acc = fma(si, rt * zn - yn, acc);
        }
      }
      if (c.var_live(j)) {
        const double x = alpha * xn + oma * X[o];
        if (live) { X[o] = x; BV[o] = (sigma * x - QV[o]) + acc; }
      }
      carry = carry_new;
    }
    xn = xt;
#pragma unroll
    for (int q = 0; q < 8; ++q) gn[q] = g[q];
  }
}

// right-hand side of the ADMM step from the current iterate: B = sigma x - q + A'(rho z - y)   (start, after rho updates)
template <int KIND>
__device__ __noinline__ void rhs_admm(const Ctx<KIND> c, const bool doit, const double rho, const double rho_eq, const double sigma,
                                     const double zsel) {
  constexpr int NX = Ctx<KIND>::NX, NT = Ctx<KIND>::NT;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  double *S = c.S;
  double *BV = S + L.B;
  const double *X = S + L.X, *YD = S + L.YD, *ZI = S + L.ZI, *YI = S + L.YI, *ED = S + L.ED, *BE = S + L.BE, *QV = S + L.Q, *SI = S + L.SI;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    double acc = 0.0;
    if (c.xl) acc = ED[o] * (rho_eq * (zsel * BE[o]) - YD[o]);
    if (k < N) {
      double g[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? (rho_eq * (zsel * BE[(k + 1) * 8 + rr]) - YD[(k + 1) * 8 + rr]) : 0.0;
      acc += coldot<NX>(c.Gb(k), c.co, g);
    }
    if (c.has_in(k)) {
      const double rt = c.rho_of(k, rho, rho_eq);
#pragma unroll
      for (int t = 0; t < NT; ++t) { const int oi = k * 8 + c.islot + t; acc = fma(SI[oi], rt * ZI[oi] - YI[oi], acc); }
    }
    if (doit) BV[o] = c.var_live(k) ? ((sigma * X[o] - QV[o]) + acc) : 0.0;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- per-QP scalars shared by the cold routines
struct Info {
  double pri_res, dua_res, obj;
  double n_rp, n_z, n_Ax, n_rd, n_q, n_Aty, n_Px;  // scaled-space inf-norms of the last update_info
  double u_z, u_Ax, u_q, u_Aty, u_Px;              // the same, unscaled (termination)
  double csc, cinv;
  int status, unscale;
};

// (P v)_(k, r) for a vector stored [k*8 + r] (diagonal Q, R and the slew-rate coupling of the inputs)
template <int KIND>
__device__ __forceinline__ double rowP(const Ctx<KIND> &c, const double *PD, const double *PO, const double *v, int k) {
  const int N = c.N, o = k * 8 + c.r;
  double acc = PD[o] * v[o];
  if (c.ul) {
    if (k > 0 && k < N) acc = fma(PO[o - 8], v[o - 8], acc);
    if (k < N - 1) acc = fma(PO[o], v[o + 8], acc);
    if (k == N) acc = 0.0;
  }
  return acc;
}
// (A v) on my dynamics row (k, r): v stored [k*8 + r]
template <int KIND>
__device__ __forceinline__ double rowA_dyn(const Ctx<KIND> &c, const double *v, int k) {
  const double *ED = c.S + c.L->ED;
  double acc = ED[k * 8 + c.r] * v[k * 8 + c.r];
  if (k > 0) {
    double g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = v[(k - 1) * 8 + q];
    acc += rowdot(c.Gb(k - 1), c.ro, g);
  }
  return acc;
}
// (A' t)_(k, r): td on the dynamics rows [k*8 + rr], ti on the single-variable rows [k*8 + slot]
template <int KIND>
__device__ __forceinline__ double colA(const Ctx<KIND> &c, const double *td, const double *ti, int k) {
  constexpr int NX = Ctx<KIND>::NX, NT = Ctx<KIND>::NT;
  const double *ED = c.S + c.L->ED, *SI = c.S + c.L->SI;
  const int o = k * 8 + c.r;
  double acc = c.xl ? ED[o] * td[o] : 0.0;
  if (k < c.N) {
    double g[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? td[(k + 1) * 8 + rr] : 0.0;
    acc += coldot<NX>(c.Gb(k), c.co, g);
  }
  if (c.has_in(k)) {
#pragma unroll
    for (int t = 0; t < NT; ++t) acc = fma(SI[k * 8 + c.islot + t], ti[k * 8 + c.islot + t], acc);
  }
  return acc;
}

// residual norms at the current iterate (update_info)
template <int KIND>
__device__ __noinline__ void update_info(const Ctx<KIND> c, Info *ip, const double zsel) {
  constexpr int NT = Ctx<KIND>::NT;
  Info &I = *ip;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  const double *S = c.S;
  const double *X = S + L.X, *YD = S + L.YD, *ZI = S + L.ZI, *YI = S + L.YI, *BE = S + L.BE, *QV = S + L.Q, *SI = S + L.SI;
  const double *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV), *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double a_rp = 0, a_z = 0, a_Ax = 0, b_rp = 0, b_z = 0, b_Ax = 0;
  double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, b_rd = 0, b_q = 0, b_Aty = 0, b_Px = 0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    if (c.xl) {
      const double Ax = rowA_dyn<KIND>(c, X, k), z = zsel * BE[o], rr = Ax - z, ei = EINV[o];
      a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, Ax);
      b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * Ax);
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oi = k * 8 + c.islot + t;
        const double ax = SI[oi] * X[o], z = ZI[oi], rr = ax - z, ei = EIINV[oi];
        a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, ax);
        b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * ax);
      }
    }
    if (c.var_live(k)) {
      const double Px = rowP<KIND>(c, PD, PO, X, k), Aty = colA<KIND>(c, YD, YI, k);
      const double rr = (QV[o] + Px) + Aty, di = DINV[o];
      a_rd = absmax(a_rd, rr); a_q = absmax(a_q, QV[o]); a_Aty = absmax(a_Aty, Aty); a_Px = absmax(a_Px, Px);
      b_rd = absmax(b_rd, di * rr); b_q = absmax(b_q, di * QV[o]); b_Aty = absmax(b_Aty, di * Aty); b_Px = absmax(b_Px, di * Px);
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
// delta_y / delta_x against the iterate saved before the last ADMM step (C_PV*).  The projected delta_y is written to
// the polish scratch (C_PYD / C_PYI), which is free while ADMM runs.
template <int KIND>
__device__ __noinline__ bool primal_infeasible(const Ctx<KIND> c, const Info *ip, const double eps) {
  constexpr int NT = Ctx<KIND>::NT;
  const bool unscale = ip->unscale;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  const double *S = c.S;
  const double *YD = S + L.YD, *YI = S + L.YI, *BE = S + L.BE, *UI = S + L.UI, *LI = S + L.LI;
  const double *PVYD = c.cd(C_PVYD), *PVYI = c.cd(C_PVYI), *E = c.cd(C_E), *EI = c.cd(C_EI), *DINV = c.cd(C_DINV);
  double *DYD = c.cd(C_PYD), *DYI = c.cd(C_PYI);
  double nrm = 0.0, lhs = 0.0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    double d = c.xl ? YD[o] - PVYD[o] : 0.0;  // equality rows: both bounds finite, no projection
    DYD[o] = d;
    if (c.xl) {
      nrm = absmax(nrm, unscale ? E[o] * d : d);
      lhs += BE[o] * ((d > 0) ? d : 0) + BE[o] * ((d < 0) ? d : 0);
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oi = k * 8 + c.islot + t;
        double di = YI[oi] - PVYI[oi];
        const double lo = LI[oi], up = UI[oi];
        if (up > kInfty * kMinScaling) {
          if (lo < -kInfty * kMinScaling) di = 0.0;
          else di = (di < 0.0) ? di : 0.0;
        } else if (lo < -kInfty * kMinScaling) di = (di > 0.0) ? di : 0.0;
        DYI[oi] = di;
        nrm = absmax(nrm, unscale ? EI[oi] * di : di);
        lhs += up * ((di > 0) ? di : 0) + lo * ((di < 0) ? di : 0);
      }
    }
  }
  nrm = gmax(nrm);
  lhs = gsum(lhs);
  __syncwarp();
  double mx = 0.0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    if (c.var_live(k)) {
      const double at = colA<KIND>(c, DYD, DYI, k);
      mx = absmax(mx, unscale ? DINV[k * 8 + r] * at : at);
    }
  }
  mx = gmax(mx);
  return (nrm > eps) && (lhs < -eps * nrm) && (mx < eps * nrm);
}

template <int KIND>
__device__ __noinline__ bool dual_infeasible(const Ctx<KIND> c, const Info *ip, const double eps) {
  constexpr int NT = Ctx<KIND>::NT;
  const bool unscale = ip->unscale;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  const double *S = c.S;
  const double *X = S + L.X, *QV = S + L.Q, *SI = S + L.SI, *UI = S + L.UI, *LI = S + L.LI;
  const double *PVX = c.cd(C_PVX), *D = c.cd(C_D), *DINV = c.cd(C_DINV), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV);
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double *DX = c.cd(C_PX);
  double nrm = 0.0, qdx = 0.0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    const double dx = c.var_live(k) ? X[o] - PVX[o] : 0.0;
    DX[o] = dx;
    nrm = absmax(nrm, unscale ? D[o] * dx : dx);
    qdx += QV[o] * dx;
  }
  nrm = gmax(nrm); qdx = gsum(qdx);
  __syncwarp();
  const double cs = unscale ? ip->csc : 1.0;
  double mx = 0.0;
  int viol = 0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    if (c.var_live(k)) {
      const double Pdx = rowP<KIND>(c, PD, PO, DX, k);
      mx = absmax(mx, unscale ? DINV[o] * Pdx : Pdx);
    }
    if (c.xl) {
      double v = rowA_dyn<KIND>(c, DX, k);
      if (unscale) v = EINV[o] * v;
      if (v > eps * nrm || v < -eps * nrm) viol = 1;
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oi = k * 8 + c.islot + t;
        double v = SI[oi] * DX[o];
        if (unscale) v = EIINV[oi] * v;
        if (((UI[oi] < kInfty * kMinScaling) && (v > eps * nrm)) || ((LI[oi] > -kInfty * kMinScaling) && (v < -eps * nrm))) viol = 1;
      }
    }
  }
  mx = gmax(mx);
  viol = gany(viol);
  return (nrm > eps) && (qdx < -cs * eps * nrm) && (mx < cs * eps * nrm) && !viol;
}

// returns 1 when a termination status was set for my group (check_termination)
template <int KIND>
__device__ __noinline__ int check_termination(const Ctx<KIND> c, const lpvmpc_settings &S, Info *ip, const bool live, const int approximate) {
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
  if (__any_sync(kFull, live && !ncvx && !prim_ok)) prim_inf = primal_infeasible<KIND>(c, ip, eps_pi) && !prim_ok;
  if (__any_sync(kFull, live && !ncvx && !dual_ok)) dual_inf = dual_infeasible<KIND>(c, ip, eps_di) && !dual_ok;
  if (!live) return 0;
  if (ncvx) { I.status = LPVMPC_NON_CVX; I.obj = nan(""); return 1; }
  if (prim_ok && dual_ok) { I.status = approximate ? LPVMPC_SOLVED_INACCURATE : LPVMPC_SOLVED; return 1; }
  if (prim_inf) { I.status = approximate ? LPVMPC_PRIMAL_INFEASIBLE_INACCURATE : LPVMPC_PRIMAL_INFEASIBLE; I.obj = kInfty; return 1; }
  if (dual_inf) { I.status = approximate ? LPVMPC_DUAL_INFEASIBLE_INACCURATE : LPVMPC_DUAL_INFEASIBLE; I.obj = -kInfty; return 1; }
  return 0;
}

template <int KIND>
__device__ __noinline__ double objective(const Ctx<KIND> c, const double *xv, const double scale) {
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *QV = c.S + c.L->Q;
  double acc = 0.0;
#pragma unroll 1
  for (int k = 0; k <= c.N; ++k)
    if (c.var_live(k)) acc += (0.5 * rowP<KIND>(c, PD, PO, xv, k) + QV[k * 8 + c.r]) * xv[k * 8 + c.r];
  return gsum(acc) * scale;
}

// ---------------------------------------------------------------- setup: schedule + build + Ruiz (cold, once per QP)
// Returns flags: bit 0 = Curvature() failed, bit 1 = l > u somewhere.
template <int KIND>
__device__ __noinline__ int setup(const Ctx<KIND> c, const G8Params &p, const int b, const bool valid, double *csc_out,
                                 uint64_t *eqm_out, uint64_t *loosem_out) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &St = p.S;
  double *S = c.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N, NS8 = (N + 1) * 8;
  const int ucomp = r - NX;
  int sched_err = 0, data_err = 0;
  double x0r = 0.0;
  // ---- schedule: G_k = -[A_k B_k] (unscaled), my row
  if (a.sched_mode == LPVMPC_SCHED_GIVEN) {
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      if (c.xl) {
        const double *Ar = a.A + ((size_t)b * N + k) * NX * NX + r * NX, *Br = a.Bm + ((size_t)b * N + k) * NX * 2 + r * 2;
        double row[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) row[cc] = (cc < NX) ? -Ar[cc < NX ? cc : 0] : ((cc < NB) ? -Br[cc - NX < 2 ? cc - NX : 0] : 0.0);
        double *gk = c.Gb(k);
#pragma unroll
        for (int j = 0; j < 4; ++j) st2(gk + c.ro[j], row[2 * j], row[2 * j + 1]);
      }
    }
    x0r = c.xl ? a.x0[(size_t)b * NX + r] : 0.0;
  } else {
    const bool predict = a.sched_mode == LPVMPC_SCHED_PREDICT;
    double st[NX];
    const double *xs = (predict && a.x_sched) ? a.x_sched + (size_t)b * NX : a.x0 + (size_t)b * NX;
#pragma unroll
    for (int q = 0; q < NX; ++q) st[q] = predict ? xs[q] : 0.0;
    const double *up = a.u_prev + (size_t)b * N * 2;
    const int lap = a.lap ? a.lap[b] : a.lap_all;
    x0r = c.xl ? a.x0[(size_t)b * NX + r] : 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      const double delta = up[k * 2];
      double Ai[NX * NX], Bi[NX * 2];
      if (KIND == LPVMPC_CONTROLLER) {
        double vx, vy, epsi, ey, cur, Cf, Cr;
        if (predict) {
          vy = st[1]; epsi = st[3]; ey = st[NX - 1];
          cur = (lap == 0) ? curvature(M.track, M.nseg, st[4], sched_err) : a.curv_ref[(size_t)b * N + k];
          vx = a.vel_ref[(size_t)b * (N + 1) + k];
          Cf = a.Cf_new; Cr = a.Cf_new;
        } else {
          const double *t = a.traj + ((size_t)b * N + k) * 6;
          vx = t[0]; vy = t[1]; epsi = t[3]; ey = t[5];
          cur = curvature(M.track, M.nseg, t[4], sched_err);
          Cf = M.Cf; Cr = M.Cr;
        }
        ctrl_stage(M, Cf, Cr, vx, vy, epsi, ey, cur, delta, Ai, Bi);
      } else {
        if (predict) {
          const double cur = curvature(M.track, M.nseg, a.SS[(size_t)b * (N + 1) + k], sched_err);
          plan_stage(M, st[0], st[1], st[3], st[4], cur, delta, Ai, Bi);
        } else {
          const double *t = a.traj + ((size_t)b * N + k) * 6;
          const double cur = curvature(M.track, M.nseg, t[5], sched_err);
          plan_stage(M, t[0], t[1], t[3], t[4], cur, delta, Ai, Bi);
        }
      }
      double row[8];  // my row of [A B], selected without dynamic register indexing
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        double v = 0.0;
        if (cc < NB) {
#pragma unroll
          for (int rr = 0; rr < NX; ++rr) {
            const double e = (cc < NX) ? Ai[rr * NX + (cc < NX ? cc : 0)] : Bi[rr * 2 + (cc - NX < 2 ? cc - NX : 0)];
            v = (r == rr) ? e : v;
          }
        }
        row[cc] = v;
      }
      if (c.xl) {
        double *gk = c.Gb(k);
#pragma unroll
        for (int j = 0; j < 4; ++j) st2(gk + c.ro[j], -row[2 * j], -row[2 * j + 1]);
        if (valid && a.A_out) {
#pragma unroll
          for (int cc = 0; cc < NX; ++cc) a.A_out[((size_t)b * N + k) * NX * NX + r * NX + cc] = row[cc];
        }
        if (valid && a.B_out) { a.B_out[((size_t)b * N + k) * NX * 2 + r * 2] = row[NX]; a.B_out[((size_t)b * N + k) * NX * 2 + r * 2 + 1] = row[NX + 1]; }
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
  sched_err = gany(sched_err);

  // ---- build (PathFollowingLPVMPC.py:334-348, 397-464; LPV_MPC_Planner.py:145-181).  Scratch in the factor area.
  double *X = S + L.X, *QV = S + L.Q, *BV = S + L.B, *YD = S + L.YD, *BE = S + L.BE, *ED = S + L.ED;
  double *ZI = S + L.ZI, *YI = S + L.YI, *SI = S + L.SI, *UI = S + L.UI, *LI = S + L.LI;
  double *sD = S + L.T, *sE = sD + NS8, *sEI = sE + NS8, *sDt = sEI + NS8, *sEt = sDt + NS8, *sEti = sEt + NS8;
  double *sPD = sEti + NS8, *sPO = sPD + NS8, *scr = sPO + NS8;  // 8 arrays of NS8 + nz doubles <= (N+1)*64 + N*64
  {
    const double Qrr = c.xl ? M.Q[r * NX + r] : 0.0;
    const double Q0r = c.xl ? M.Q[r] : 0.0;
    const double Rcc = c.ul ? M.R[ucomp * 2 + ucomp] : 0.0;
    const double dRc = c.ul ? M.dR[ucomp] : 0.0;
    const double uold = (c.ul && a.u_old) ? a.u_old[(size_t)b * 2 + ucomp] : 0.0;
    const double mey = (KIND == LPVMPC_PLANNER) ? a.max_ey[b] : 0.0;
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      double pd = 0.0, po = 0.0, q = 0.0, be = 0.0, ed = 0.0;
      if (c.xl) {
        pd = 2 * Qrr;
        if (KIND == LPVMPC_CONTROLLER) q = -2 * (a.vel_ref[(size_t)b * (N + 1) + k] * Q0r);
        else q = M.L_cf[r];
        be = (k == 0) ? (x0r + 0.0) : (0.0 + (a.C ? a.C[((size_t)b * N + (k - 1)) * NX + r] : 0.0));
        ed = 1.0;
      } else if (c.ul) {
        double v = Rcc + 2 * dRc;
        if (k == N - 1) v = v - dRc;
        pd = (k < N) ? 2 * v : 0.0;
        if (KIND == LPVMPC_CONTROLLER) q = (k == 0) ? -2 * (uold * dRc) : -2 * 0.0;
        else q = (k == 0) ? -2 * (uold * dRc) : 0.0;
        po = (k < N - 1) ? 2 * (-dRc) : 0.0;
      }
      sPD[o] = pd; sPO[o] = po; QV[o] = q; BE[o] = be; ED[o] = ed;
      sD[o] = 1.0; sE[o] = 1.0; sEI[o] = 1.0;
      X[o] = 0.0; YD[o] = 0.0; ZI[o] = 0.0; YI[o] = 0.0;
      SI[o] = 0.0; UI[o] = 0.0; LI[o] = 0.0;
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oi = k * 8 + c.islot + t;
          double si, lo, up;
          if (KIND == LPVMPC_CONTROLLER) {
            lo = -kInfty;
            if (r == 0) { si = t ? 1.0 : -1.0; up = t ? M.max_vel : -0.01; }
            else if (r == NX) { si = t ? -1.0 : 1.0; up = 0.249; }
            else { si = t ? -1.0 : 1.0; up = t ? 1.0 : 4.0; }
          } else {
            si = 1.0;
            if (c.xl) {
              lo = (r == 0) ? M.min_vel : (r == 1 ? -1.0 : (r == 2 ? -2.0 : (r == 3 ? -mey : -0.8)));
              up = (r == 0) ? M.max_vel : (r == 1 ? 1.0 : (r == 2 ? 2.0 : (r == 3 ? mey : 0.8)));
              if (r == 3 && a.ey_lo) lo = a.ey_lo[(size_t)b * (N + 1) + k];
              if (r == 3 && a.ey_hi) up = a.ey_hi[(size_t)b * (N + 1) + k];
            } else { lo = ucomp ? -0.7 : -0.249; up = ucomp ? 2.0 : 0.249; }
          }
          lo = (lo > -kInfty) ? lo : -kInfty;  // python wrapper: l = max(l, -OSQP_INFTY), u = min(u, OSQP_INFTY)
          up = (up < kInfty) ? up : kInfty;
          if (lo > up) data_err = 1;
          SI[oi] = si; LI[oi] = lo; UI[oi] = up;
        }
      }
    }
    __syncwarp();
  }
  data_err = gany(data_err);

  // ---- Ruiz equilibration (OSQP scale_data)
  double csc = 1.0;
#pragma unroll 1
  for (int it = 0; it < St.scaling; ++it) {
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      double pa = fabs(sPD[o]);
      if (c.ul) {
        if (k < N - 1) pa = absmax(pa, sPO[o]);
        if (k > 0 && k < N) pa = absmax(pa, sPO[o - 8]);
      }
      double qa = c.xl ? fabs(ED[o]) : 0.0;
      if (k < N) {
        const double *gk = c.Gb(k);
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) qa = absmax(qa, gk[rr * 8 + c.co[rr >> 1]]);
      }
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oi = k * 8 + c.islot + t;
          qa = absmax(qa, SI[oi]);
          sEti[oi] = 1.0 / sqrt(limit_scaling(fabs(SI[oi])));
        }
      }
      sDt[o] = 1.0 / sqrt(limit_scaling(pa > qa ? pa : qa));
      double ea = c.xl ? fabs(ED[o]) : 0.0;
      if (k > 0 && c.xl) {
        const double *gp = c.Gb(k - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(gp + c.ro[j]); ea = absmax(ea, e.x); ea = absmax(ea, e.y); }
      }
      sEt[o] = 1.0 / sqrt(limit_scaling(ea));
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      const double dt = sDt[o];
      if (k < N) {
        double *gk = c.Gb(k);
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) { double *e = gk + rr * 8 + c.co[rr >> 1]; *e = (*e * sEt[(k + 1) * 8 + rr]) * dt; }
        if (k < N - 1 && c.ul) sPO[o] = (sPO[o] * dt) * sDt[o + 8];
      }
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oi = k * 8 + c.islot + t;
          SI[oi] = (SI[oi] * sEti[oi]) * dt;
          sEI[oi] = sEI[oi] * sEti[oi];
        }
      }
      if (c.xl) ED[o] = (ED[o] * sEt[o]) * dt;
      sPD[o] = (sPD[o] * dt) * dt;
      QV[o] = dt * QV[o];
      sD[o] = sD[o] * dt;
      sE[o] = sE[o] * sEt[o];
    }
    __syncwarp();
    // cost scaling: mean of the column norms of P in the reference variable order
    double qn = 0.0;
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      double pa = fabs(sPD[o]);
      if (c.ul) {
        if (k < N - 1) pa = absmax(pa, sPO[o]);
        if (k > 0 && k < N) pa = absmax(pa, sPO[o - 8]);
      }
      if (c.xl) scr[k * NX + r] = pa;
      else if (c.ul && k < N) scr[nx + k * 2 + ucomp] = pa;
      if (c.var_live(k)) qn = absmax(qn, QV[o]);
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
#pragma unroll 1
    for (int k = 0; k <= N; ++k) { const int o = k * 8 + r; sPD[o] *= ct; QV[o] *= ct; sPO[o] *= ct; }
    csc *= ct;
    __syncwarp();
  }
  *csc_out = csc;
  // ---- scaled bounds, constraint classes, first right-hand side, cold data
  {
    double *cD = c.cd(C_D), *cDI = c.cd(C_DINV), *cE = c.cd(C_E), *cEI = c.cd(C_EINV), *cPD = c.cd(C_PD), *cPO = c.cd(C_PO);
    double *cEi = c.cd(C_EI), *cEiI = c.cd(C_EIINV);
    uint64_t eqm = 0, loosem = 0;
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      BE[o] = sE[o] * BE[o];
      cD[o] = sD[o]; cDI[o] = 1.0 / sD[o]; cE[o] = sE[o]; cEI[o] = 1.0 / sE[o]; cPD[o] = sPD[o]; cPO[o] = sPO[o];
      cEi[o] = sEI[o]; cEiI[o] = 1.0 / sEI[o];
      BV[o] = c.var_live(k) ? -QV[o] : 0.0;
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oi = k * 8 + c.islot + t;
          const double lo = sEI[oi] * LI[oi], up = sEI[oi] * UI[oi];
          LI[oi] = lo; UI[oi] = up;
          if (KIND == LPVMPC_PLANNER) {
            if ((lo < -kInfty * kMinScaling) && (up > kInfty * kMinScaling)) loosem |= 1ull << k;
            else if (up - lo < kRhoTol) eqm |= 1ull << k;
          }
        }
      }
    }
    *eqm_out = eqm; *loosem_out = loosem;
    __syncwarp();
  }
  return (sched_err ? 1 : 0) | (data_err ? 2 : 0);
}

// ---------------------------------------------------------------- polish (cold, once per QP)
// Works on the scratch slab (C_PX, C_PYD, C_PYI); on success the polished (x, z, y) replace the iterate.
template <int KIND>
__device__ __noinline__ int polish(const Ctx<KIND> c, const lpvmpc_settings &St, Info *ip, const bool do_pol) {
  constexpr int NT = Ctx<KIND>::NT;
  Info &I = *ip;
  const bool unscale = I.unscale;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  double *S = c.S;
  double *X = S + L.X, *YD = S + L.YD, *ZI = S + L.ZI, *YI = S + L.YI, *BV = S + L.B;
  const double *BE = S + L.BE, *QV = S + L.Q, *SI = S + L.SI, *UI = S + L.UI, *LI = S + L.LI;
  double *PX = c.cd(C_PX), *PYD = c.cd(C_PYD), *PYI = c.cd(C_PYI), *R2D = c.cd(C_R2D), *R2I = c.cd(C_R2I);
  double *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV);
  const double delta = St.delta, idel = 1.0 / St.delta;
  // active-set guess (form_Ared): 1 = lower, 2 = upper, 3 = both
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    double ad = 0.0;
    if (c.xl && do_pol) { if (0.0 < -YD[o]) ad += 1.0; if (0.0 < YD[o]) ad += 2.0; }  // equality row: z == l == u
    ACTD[o] = ad;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oi = k * 8 + c.islot + t;
        double ai = 0.0;
        if (do_pol) { if (ZI[oi] - LI[oi] < -YI[oi]) ai += 1.0; if (UI[oi] - ZI[oi] < YI[oi]) ai += 2.0; }
        ACTI[oi] = ai;
      }
    }
  }
  __syncwarp();
  FW fw; fw.polish = 1; fw.rho = 0.0; fw.rho_eq = 0.0; fw.idel = idel;
  factor<KIND>(c, fw, delta);
  auto bred_i = [&](int oi) { const double a = ACTI[oi]; return (a == 1.0 || a == 3.0) ? LI[oi] : UI[oi]; };
  // first solve: rhs = -q + A_red'(b_red / delta); targets go through R2D / R2I
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    R2D[o] = (c.xl && ACTD[o] != 0.0) ? idel * BE[o] : 0.0;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) { const int oi = k * 8 + c.islot + t; R2I[oi] = (ACTI[oi] != 0.0) ? idel * bred_i(oi) : 0.0; }
    }
  }
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k <= N; ++k) { const int o = k * 8 + r; BV[o] = c.var_live(k) ? (-QV[o] + colA<KIND>(c, R2D, R2I, k)) : 0.0; }
  __syncwarp();
  sweep_fwd<KIND>(c, true);
  sweep_bwd_plain<KIND>(c, R2D);  // R2D <- A_dyn x (all dynamics rows)
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    PX[o] = BV[o];
    PYD[o] = (c.xl && ACTD[o] != 0.0) ? (R2D[o] - BE[o]) * idel : 0.0;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) { const int oi = k * 8 + c.islot + t; PYI[oi] = (ACTI[oi] != 0.0) ? (SI[oi] * BV[o] - bred_i(oi)) * idel : 0.0; }
    }
  }
  __syncwarp();
#pragma unroll 1
  for (int it = 0; it < St.polish_refine_iter + kPolishExtraRefine; ++it) {
    // residual of the un-regularised reduced KKT: r2 on the active rows
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      R2D[o] = (c.xl && ACTD[o] != 0.0) ? (BE[o] - rowA_dyn<KIND>(c, PX, k)) : 0.0;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) { const int oi = k * 8 + c.islot + t; R2I[oi] = (ACTI[oi] != 0.0) ? (bred_i(oi) - SI[oi] * PX[o]) : 0.0; }
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      double b = 0.0;
      if (c.var_live(k)) {
        const double Px = rowP<KIND>(c, PD, PO, PX, k), Aty = colA<KIND>(c, PYD, PYI, k);
        // A'(r2 / delta): same column product on scaled entries
        constexpr int NX = Ctx<KIND>::NX;
        double at = c.xl ? (S + L.ED)[o] * (idel * R2D[o]) : 0.0;
        if (k < N) {
          double g[8];
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? idel * R2D[(k + 1) * 8 + rr] : 0.0;
          at += coldot<NX>(c.Gb(k), c.co, g);
        }
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) at = fma(SI[k * 8 + c.islot + t], idel * R2I[k * 8 + c.islot + t], at);
        }
        b = ((-QV[o] - Px) - Aty) + at;
      }
      BV[o] = b;
    }
    __syncwarp();
    sweep_fwd<KIND>(c, true);
    double *ZT = c.cd(C_PVYD);  // free after termination: z~ = A_dyn dx
    sweep_bwd_plain<KIND>(c, ZT);
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r;
      if (c.xl && ACTD[o] != 0.0) PYD[o] += (ZT[o] - R2D[o]) * idel;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) { const int oi = k * 8 + c.islot + t; if (ACTI[oi] != 0.0) PYI[oi] += (SI[oi] * BV[o] - R2I[oi]) * idel; }
      }
      PX[o] += BV[o];
    }
    __syncwarp();
  }
  // pol z = A x, normal-cone projection, residuals, acceptance.  Polished z of the single-variable rows -> R2I.
  double a_rp = 0, a_rd = 0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    if (c.xl) {
      const double Ax = rowA_dyn<KIND>(c, PX, k), t = Ax + PYD[o];
      PYD[o] = t - BE[o];
      const double rr = Ax - BE[o];
      a_rp = absmax(a_rp, unscale ? EINV[o] * rr : rr);
    } else PYD[o] = 0.0;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oi = k * 8 + c.islot + t;
        const double ax = SI[oi] * PX[o], tt = ax + PYI[oi];
        const double zc = clampd(tt, LI[oi], UI[oi]);
        R2I[oi] = zc; PYI[oi] = tt - zc;
        const double rr = ax - zc;
        a_rp = absmax(a_rp, unscale ? EIINV[oi] * rr : rr);
      }
    }
  }
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    if (c.var_live(k)) {
      const double rr = (QV[o] + rowP<KIND>(c, PD, PO, PX, k)) + colA<KIND>(c, PYD, PYI, k);
      a_rd = absmax(a_rd, unscale ? DINV[o] * rr : rr);
    }
  }
  const double pol_pri = gmax(a_rp), pol_dua = (unscale ? I.cinv : 1.0) * gmax(a_rd);
  const double pol_obj = objective<KIND>(c, PX, St.scaling ? I.cinv : 1.0);
  const bool ok = (pol_pri < I.pri_res && pol_dua < I.dua_res) || (pol_pri < I.pri_res && I.dua_res < 1e-10) ||
                  (pol_dua < I.dua_res && I.pri_res < 1e-10);
  if (!do_pol) return 0;
  if (!ok) return -1;
  I.obj = pol_obj; I.pri_res = pol_pri; I.dua_res = pol_dua;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    X[o] = PX[o]; YD[o] = PYD[o];
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) { const int oi = k * 8 + c.islot + t; ZI[oi] = R2I[oi]; YI[oi] = PYI[oi]; }
    }
  }
  return 1;
}

// ---------------------------------------------------------------- one persistent warp = QPW QPs at a time
template <int KIND, int QPW>
__global__ void __launch_bounds__(32) lpv_solve_g8_kernel(const __grid_constant__ G8Params p) {
  extern __shared__ double smem[];
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT;
  const int lane = threadIdx.x;
  const int g = lane >> 3, r = lane & 7;
  const Lay &L = p.L;
  const int N = L.N;
  Ctx<KIND> c;
  c.S = smem; c.cold = p.cold;
  c.L = &L; c.N = N; c.r = r;
  c.xl = r < NX; c.ul = (r >= NX) && (r < NB);
  if (KIND == LPVMPC_CONTROLLER) c.islot = (r == 0) ? 0 : ((r >= NX) ? (r - NX + 1) * 2 : 0);
  else c.islot = r;
#pragma unroll
  for (int j = 0; j < 4; ++j) { c.ro[j] = chunk(r, j); c.co[j] = (((r >> 1) ^ j) << 1) | (r & 1); }
  c.eqm = 0; c.loosem = 0;
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &S = p.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N;
  const int m = (KIND == LPVMPC_CONTROLLER) ? (6 * N + nx) : (nx + nz);
  const int ucomp = r - NX;

  for (;;) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(p.queue, (unsigned)QPW);
    base = __shfl_sync(kFull, base, 0);
    if ((int)base >= p.B) break;
    // Groups without a problem of their own (batch tail, or g >= QPW) mirror group 0 exactly: same problem, same
    // shared / scratch region, same values written by the same instruction; only user-visible outputs are guarded.
    const bool valid = (g < QPW) && ((int)(base + g) < p.B);
    const int gq = valid ? g : 0;
    const int b = (int)base + gq;
    c.S = smem + gq * L.total;
    c.cold = p.cold + (size_t)(blockIdx.x * QPW + gq) * L.cold_total;

    Info I;
    double csc = 1.0;
    int flags = 0;
    {
      uint64_t eqm = 0, loosem = 0;
      flags = setup<KIND>(c, p, b, valid, &csc, &eqm, &loosem);
      c.eqm = eqm; c.loosem = loosem;
    }
    I.csc = csc; I.cinv = 1.0 / csc;
    I.unscale = (S.scaling && !S.scaled_termination) ? 1 : 0;
    I.pri_res = 0.0; I.dua_res = 0.0; I.obj = nan("");
    I.n_rp = I.n_z = I.n_Ax = I.n_rd = I.n_q = I.n_Aty = I.n_Px = 0.0;
    I.u_z = I.u_Ax = I.u_q = I.u_Aty = I.u_Px = 0.0;
    I.status = (flags & 1) ? LPVMPC_SCHEDULE_ERROR : ((flags & 2) ? LPVMPC_DATA_ERROR : LPVMPC_UNSOLVED);

    const double sigma = S.sigma, alpha = S.alpha;
    double rho = fmin(fmax(S.rho, kRhoMin), kRhoMax);
    double rho_eq = kRhoEqOverIneq * rho, rinv = 1.0 / rho, rinv_eq = 1.0 / rho_eq;
    {
      FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0;
      factor<KIND>(c, fw, sigma);
    }
    bool live = (flags == 0);
    const bool failed = flags != 0;
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
      for (; iter < stop; ++iter) {
        if (iter == stop - 1 && live) {  // keep the iterate before the last step of the chunk: delta_x, delta_y
          double *PVX = c.cd(C_PVX), *PVYD = c.cd(C_PVYD), *PVYI = c.cd(C_PVYI);
          const double *X = c.S + L.X, *YD = c.S + L.YD, *YI = c.S + L.YI;
#pragma unroll 1
          for (int k = 0; k <= N; ++k) { const int o = k * 8 + r; PVX[o] = X[o]; PVYD[o] = YD[o]; PVYI[o] = YI[o]; }
        }
        sweep_fwd<KIND>(c, live);
        sweep_bwd_admm<KIND>(c, live, rho, rho_eq, rinv, rinv_eq, sigma, alpha, zsel);
        zsel = 1.0;
      }
      __syncwarp();
      const bool can_check = ct && (iter % ct == 0);
      const bool can_adapt = ai && (iter % ai == 0);
      checked_last = can_check;
      if (can_check || can_adapt) {
        Info J = I;
        update_info<KIND>(c, &J, zsel);
        if (live) { I = J; iter_done = iter; }
        if (can_check) {
          if (check_termination<KIND>(c, S, &I, live, 0)) live = false;  // frozen: stores are predicated on `live`
        }
        if (can_adapt) {
          const double pr = I.n_rp / ((I.n_z > I.n_Ax ? I.n_z : I.n_Ax) + 1e-10);
          double dn = I.n_q; dn = (I.n_Aty > dn) ? I.n_Aty : dn; dn = (I.n_Px > dn) ? I.n_Px : dn;
          const double dr = I.n_rd / (dn + 1e-10);
          double rho_new = rho * sqrt(pr / (dr + 1e-10));
          rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
          const bool upd = live && ((rho_new > rho * S.adaptive_rho_tolerance) || (rho_new < rho / S.adaptive_rho_tolerance));
          if (__any_sync(kFull, upd)) {
            // groups that do not update must keep their factor: re-factorising with unchanged rho reproduces it
            if (upd) { rho = rho_new; rho_eq = kRhoEqOverIneq * rho; rinv = 1.0 / rho; rinv_eq = 1.0 / rho_eq; ++rho_updates; }
            FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0;
            __syncwarp();
            factor<KIND>(c, fw, sigma);
            rhs_admm<KIND>(c, upd, rho, rho_eq, sigma, zsel);  // the pending right-hand side was built with the old rho
          }
        }
      }
    }
    if (!checked_last && __any_sync(kFull, live)) {
      Info J = I;
      update_info<KIND>(c, &J, zsel);
      if (live) { I = J; iter_done = iter; }
      if (check_termination<KIND>(c, S, &I, live, 0)) live = false;
    }
    {
      const bool unsolved = (I.status == LPVMPC_UNSOLVED);
      if (__any_sync(kFull, unsolved)) {
        if (!check_termination<KIND>(c, S, &I, unsolved, 1) && unsolved) I.status = LPVMPC_MAX_ITER_REACHED;
      }
    }
    const int status = I.status;
    const bool has_sol = !(status == LPVMPC_PRIMAL_INFEASIBLE || status == LPVMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_DUAL_INFEASIBLE || status == LPVMPC_DUAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_NON_CVX || status == LPVMPC_SCHEDULE_ERROR || status == LPVMPC_DATA_ERROR);
    {
      const double o = objective<KIND>(c, c.S + L.X, S.scaling ? I.cinv : 1.0);
      if (has_sol) I.obj = o;
    }
    // row / variable indices in the reference order
    auto ref_dyn = [&](int k) { return (KIND == LPVMPC_CONTROLLER) ? (6 * N + k * NX + r) : (k * NX + r); };
    auto ref_in = [&](int k, int t) {
      if (KIND == LPVMPC_CONTROLLER) return (r == 0) ? (2 * k + t) : (2 * N + 4 * k + 2 * ucomp + t);
      return nx + (c.xl ? (k * NX + r) : (nx + k * 2 + ucomp));
    };
    auto ref_var = [&](int k) { return c.xl ? (k * NX + r) : (nx + k * 2 + ucomp); };
    if (valid && (a.xs || a.zs || a.ys)) {
      const double *X = c.S + L.X, *YD = c.S + L.YD, *BE = c.S + L.BE, *ZI = c.S + L.ZI, *YI = c.S + L.YI;
#pragma unroll 1
      for (int k = 0; k <= N; ++k) {
        const int o = k * 8 + r;
        if (c.var_live(k) && a.xs) a.xs[(size_t)b * nz + ref_var(k)] = X[o];
        if (c.xl) {
          if (a.zs) a.zs[(size_t)b * m + ref_dyn(k)] = (iter > 0 && !failed) ? BE[o] : 0.0;
          if (a.ys) a.ys[(size_t)b * m + ref_dyn(k)] = YD[o];
        }
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (a.zs) a.zs[(size_t)b * m + ref_in(k, t)] = ZI[k * 8 + c.islot + t];
            if (a.ys) a.ys[(size_t)b * m + ref_in(k, t)] = YI[k * 8 + c.islot + t];
          }
        }
      }
    }
    int polish_status = 0;
    const bool do_pol = S.polish && status == LPVMPC_SOLVED;
    bool polished_sets = false;
    if (__any_sync(kFull, do_pol)) {
      polish_status = polish<KIND>(c, S, &I, do_pol);
      polished_sets = do_pol;
    }
    // ---- outputs
    if (valid) {
      const double *X = c.S + L.X, *YD = c.S + L.YD, *YI = c.S + L.YI;
      const double *D = c.cd(C_D), *E = c.cd(C_E), *EI = c.cd(C_EI), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
#pragma unroll 1
      for (int k = 0; k <= N; ++k) {
        const int o = k * 8 + r;
        const double v = has_sol ? D[o] * X[o] : nan("");
        if (c.xl) a.x_pred[(size_t)b * nx + k * NX + r] = v;
        else if (c.ul && k < N) a.u_pred[(size_t)b * 2 * N + k * 2 + ucomp] = v;
        if (c.xl) {
          const size_t q = (size_t)b * m + ref_dyn(k);
          const int act = polished_sets ? (int)ACTD[o] : 0;
          if (a.y) a.y[q] = has_sol ? I.cinv * (E[o] * YD[o]) : nan("");
          if (a.active_lo) a.active_lo[q] = act & 1;
          if (a.active_up) a.active_up[q] = (act >> 1) & 1;
        }
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const int oi = k * 8 + c.islot + t;
            const size_t q = (size_t)b * m + ref_in(k, t);
            const int act = polished_sets ? (int)ACTI[oi] : 0;
            if (a.y) a.y[q] = has_sol ? I.cinv * (EI[oi] * YI[oi]) : nan("");
            if (a.active_lo) a.active_lo[q] = act & 1;
            if (a.active_up) a.active_up[q] = (act >> 1) & 1;
          }
        }
      }
      if (r == 0) {
        a.status[b] = status;
        if (a.iters) a.iters[b] = iter_done;
        if (a.rho_updates) a.rho_updates[b] = rho_updates;
        if (a.polish_status) a.polish_status[b] = polish_status;
        if (a.obj) a.obj[b] = I.obj;
        if (a.pri_res) a.pri_res[b] = failed ? nan("") : I.pri_res;
        if (a.dua_res) a.dua_res[b] = failed ? nan("") : I.dua_res;
      }
    }
    __syncwarp();
  }
}

}  // namespace g8
}  // namespace lpv
