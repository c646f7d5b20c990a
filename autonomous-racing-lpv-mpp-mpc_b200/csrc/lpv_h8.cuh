// "H8" kernel: 8 lanes per QP, QPW QPs per warp, compact code, hot state in shared memory, cold state in an L2 slab.
//
// Same algorithm as lpv_qp.cuh / lpv_g8.cuh / oracle/osqp_ref.c (OSQP 0.6: Ruiz scaling, rho vector, relaxed ADMM,
// termination + infeasibility certificates every check_termination iterations, adaptive rho, polish), restricted to
// diagonal Q and R (the reference's tunings, controllerMain.py:139-148, plannerMain.py:96-99) and steering_delay = 0.
//
// Measurements that shaped it (profiles/r1b_*, r1c_*; tools/fp64_latency.cu, tools/smem_probe.cu):
//  * the solve is SHARED-MEMORY-BANDWIDTH bound: one wavefront (<= 128 B of distinct banks) per cycle per SM, and every
//    FMA of a block mat-vec needs one fp64 operand from shared memory.  A 4-way bank conflict on an LDS.128 costs 16
//    cycles, a broadcast LDS.128 (8 lanes of a group read the same 16 B) costs 1 cycle when the four groups of the warp
//    hit different banks, an 8-lanes-consecutive LDS.64 costs 2 cycles when neighbouring groups are 64 B apart mod 128.
//  * G8 issues 4 000 instructions and 18 KB of state per QP.
//
// What H8 does about it:
//  * The dynamics rows are equalities, so their z is pinned to the right-hand side `be` after the first step and their
//    dual only enters the next right-hand side through r = A_dyn'(rho_eq z_dyn - y_dyn).  With S x~ = rhs solved,
//        rho_eq A_dyn'A_dyn x~ = rhs - (P + sigma I + A_in' rho A_in) x~  =: rhs - M x~      (M: diagonal + slew coupling)
//    so r is advanced by an element-wise recursion,
//        r <- r - alpha (rhs - M x~) + c CR,     CR = rho_eq A_dyn' be,   c = 2 on the first step, alpha afterwards,
//    and the ADMM step needs NO product with A_k / B_k at all: forward sweep, backward sweep, element-wise update.
//    The explicit dual y_dyn is recovered every kSyncEvery steps from the running sum XS of x~,
//        y_dyn += rho_eq (alpha A_dyn XS - (alpha n + (1 - alpha) [first step inside]) be),
//    and r is re-projected from it (r = CR - A_dyn' y_dyn - q): the recursion integrates the residual of the block solve
//    (1e-11 relative per step), whose component outside range(A_dyn') nothing else would remove.
//  * A_k / B_k (G), the scalings, q, be, y_dyn and every polish vector live in a per-slot global slab that stays in L2;
//    shared memory holds the block factor (T, K), 6 stage vectors and the single-variable rows: 14.0 KB per controller
//    QP at N = 8 -> 16 QPs per SM.
//  * Layout for the wavefront rules above: T_k / K_{k+1} interleaved per stage with XOR-swizzled rows (row and column
//    reads conflict free), one pointer per logical 16-byte chunk; the 8-lane all-gather goes through a per-warp
//    double-buffered 256-byte buffer laid out [chunk][group] so that a gather is 4 single-wavefront LDS.128; stage
//    vectors interleaved per stage [B X R XS DG CR]; single-variable rows as {z,y} / {s,u} pairs.
//
// Lane r of a group owns component r of every stage variable w_k = [x_k; u_k], the dynamics row (k, r) and the
// single-variable rows on its variable.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "lpv_qp.cuh"

namespace lpv {
namespace h8 {

// -DLPV_H8_PHASE_TIMING (development builds, tools/h8_phase_timing.py): thread 0 of a helper-warp CTA adds up the clock
// cycles it spends in each phase of a QP; read back through lpvmpc_debug_phase_cycles (only exported by such builds)
#ifdef LPV_H8_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define H8_PH(i) do { const long long ph_t1 = clock64(); ph[i] += ph_t1 - ph_t0; ph_t0 = ph_t1; } while (0)
#define H8_PH_PARAMS , long long *ph, long long &ph_t0
#define H8_PH_ARGS , ph, ph_t0
#else
#define H8_PH(i) do { } while (0)
#define H8_PH_PARAMS
#define H8_PH_ARGS
#endif

// Unroll factors of the chain loops of the twisted sweeps, per problem kind (the loops sit in templates on KIND).  Measured with
// reproducible builds on one box: planner (3 CTAs per SM share the instruction cache) 1: 285.8, 2: 280.9, 4: 275.5, 5: 278.9,
// 8: 291.0, 10: 282.5 ms per plan16384; long-horizon controller (1 CTA per SM) 1: 24.72, 2: 24.57, 4: 23.66, 5: 23.48,
// 8: 23.35, 10: 23.27 ms per ctrl1024N100.
#ifndef H8_TW_UNROLL_PLAN
#define H8_TW_UNROLL_PLAN 4
#endif
#ifndef H8_TW_UNROLL_CTRL
#define H8_TW_UNROLL_CTRL 10
#endif
#ifndef H8_COLD_UNROLL
#define H8_COLD_UNROLL 1   // unroll factor of the cold per-stage loops that request their slab operands up front (measured, reproducible builds, same box: 2: ctrl1024N100 23.27 -> 23.64 ms, plan16384 274.7 -> 282.4 ms; 4: 23.85 / 303.9 ms)
#endif
#define H8_STR2(x) #x
#define H8_STR(x) H8_STR2(x)
#define H8_COLD_PRAGMA _Pragma(H8_STR(unroll H8_COLD_UNROLL))
#ifndef H8_HELPER_UNROLL
#define H8_HELPER_UNROLL 1   // unroll factor of the helpers' two-rounds-at-a-time loop (2: ctrl1024N100 23.27 -> 23.15 ms, plan16384 275.3 -> 277.5 ms)
#endif
#define H8_HELPER_PRAGMA _Pragma(H8_STR(unroll H8_HELPER_UNROLL))
#ifndef H8_HELPER_TRIPLE
#define H8_HELPER_TRIPLE 0   // 1: the helpers take three rounds at a time where they can (measured: plan16384 275.7 -> 286.9 ms: the third round delays the start of the other two)
#endif
#ifndef H8_HELPER_SINGLES
#define H8_HELPER_SINGLES 4   // from this many helper warps on, the helpers update one round at a time
#endif
#define H8_TW_PRAGMA _Pragma(H8_STR(unroll (KIND == LPVMPC_PLANNER ? H8_TW_UNROLL_PLAN : H8_TW_UNROLL_CTRL)))
constexpr int TKS = 128;  // doubles per stage of the factor: T_k (64) then K_{k+1} (64)
constexpr int VS = 56;    // doubles per stage of the stage vectors
enum { V_B = 0, V_X = 8, V_R = 16, V_XS = 24, V_DG = 32, V_CR = 40, V_XT = 48 };   // XT: x~ of the backward sweep (helper-warp kernels)

struct Lay {  // per-QP offsets (doubles); computed on the host (lpvmpc.cu: make_h8_layout)
  int N, nsl;                  // horizon; single-variable-row slots per stage (6 controller, 7 planner)
  int is;                      // doubles per stage of the single-variable-row block: {z,y} x nsl, {s,u} x nsl, [l x nsl], pm x 2
  int TK, V, I;                // (N+1) x TKS - 64, (N+1) x VS, (N+2) x is (block N+1: dummy rows, zero coupling)
  int total;                   // doubles per QP in shared memory (== 8 mod 16: neighbouring groups 64 B apart mod 128)
  int cold_total;              // doubles per QP slot in the global slab
  int cG;                      // offset of G (N x NX x 8, row-major rows of -[A_k B_k], scaled) inside the slot
  // streamed factor ("H8S"): the block factor lives in the slab ((N+1) x TKS at cTK) and is staged through a ring of
  // `ring` stage blocks (power of two; 0 = factor resident in shared memory) by 1 KB TMA bulk copies issued `pfd`
  // stage steps ahead; setup's scratch then aliases the slab copy of the factor
  int ring, pfd, cTK;
};
enum { C_D = 0, C_DINV, C_E, C_EINV, C_PD, C_PO, C_EI, C_EIINV, C_Q, C_BE, C_ED, C_YD, C_PVX, C_PVYI, C_DYD, C_PX, C_PYD,
       C_PYI, C_R2D, C_R2I, C_ACTD, C_ACTI, C_ZT, C_COUNT };

struct H8Params {
  Lay L;
  Model M;
  lpvmpc_settings S;
  lpvmpc_args a;
  int B;
  unsigned *queue;
  const int *perm;   // visiting order of the batch (NULL: batch order)
  double *cold;
};

template <int KIND> struct Dims;
template <> struct Dims<LPVMPC_CONTROLLER> { static constexpr int NX = 6, NT = 2, NSL = 6, OLI = 24, OPM = 24, IS = 26; };
template <> struct Dims<LPVMPC_PLANNER> { static constexpr int NX = 5, NT = 1, NSL = 7, OLI = 28, OPM = 36, IS = 38; };

// what Ctx takes for granted about the host's layout
inline bool layout_matches(const Lay &L, int nx) {
  return L.TK == 0 && L.cG == C_COUNT * (L.N + 1) * 8 && L.cTK == L.cG + L.N * nx * 8;
}

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
// reductions of the cold loops: over my group, or over the whole warp when its four groups share ONE QP and split the
// stages between them (`wide`, warp-uniform: 1 QP per warp)
__device__ __forceinline__ double wmax(double v, const bool wide) {
  v = gmax(v);
  if (wide) {
    double w = __shfl_xor_sync(kFull, v, 8); v = (w > v) ? w : v;
    w = __shfl_xor_sync(kFull, v, 16); v = (w > v) ? w : v;
  }
  return v;
}
__device__ __forceinline__ double wsum(double v, const bool wide) {
  v = gsum(v);
  if (wide) { v += __shfl_xor_sync(kFull, v, 8); v += __shfl_xor_sync(kFull, v, 16); }
  return v;
}
__device__ __forceinline__ int wany(int v, const bool wide) {
  const unsigned m = __ballot_sync(kFull, v);
  return wide ? (m != 0u) : (((m >> ((threadIdx.x & 31) & ~7)) & 0xffu) != 0);
}
// Reciprocal and reciprocal square root from the hardware seed (2^-23) and three Newton steps: ~1 ulp, a dozen
// instructions and half the latency of the IEEE division / sqrt sequences (70+ instructions each).
__device__ __forceinline__ double frcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0); r = fma(r, e, r);
  e = fma(-x, r, 1.0); r = fma(r, e, r);
  e = fma(-x, r, 1.0); r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double frsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5); y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5); y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5); y = fma(y, e, y);
  return y;
}
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
// slab (global memory) loads that stay where they are written: the compiler neither sinks them into the conditional block
// that uses the value nor reorders them (the cold loops request every operand of a stage before the first use)
__device__ __forceinline__ double ldg_v(const double *p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(__cvta_generic_to_global(p)));
  return v;
}
__device__ __forceinline__ double2 ldg2_v(const double *p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(__cvta_generic_to_global(p)));
  return v;
}
__device__ __forceinline__ int swz(int rr) { return (rr >> 1) & 3; }
// offset of logical 16-byte chunk j of row rr inside a swizzled 8-wide block
__device__ __forceinline__ int chunk(int rr, int j) { return rr * 8 + ((j ^ swz(rr)) << 1); }

template <int KIND>
struct Ctx {
  static constexpr int NX = Dims<KIND>::NX, NB = NX + 2, NT = Dims<KIND>::NT, NSL = Dims<KIND>::NSL;
  static constexpr int OLI = Dims<KIND>::OLI, OPM = Dims<KIND>::OPM, IS = Dims<KIND>::IS;
  double *S;     // my QP's shared region
  double *cold;  // my QP slot in the global slab
  // The layout by value (lpvmpc.cu: make_h8_layout; checked at create with layout_matches): the factor starts the region, the
  // stage vectors follow at oV, the single-variable rows at oI; the slab holds C_COUNT vectors, then G, then (streamed factor)
  // the factor.  Through a `const Lay *` into the kernel parameters every accessor of a cold routine paid a generic load from
  // parameter memory in front of its address.
  int oV, oI, ring;
  int N, r;
  int kmask;     // stage -> factor block slot: -1 (resident: slot k) or ring - 1
  int k0, ks;    // stages of the cold per-stage loops handled by my group: k0, k0 + ks, ... (0, 1; or group index, 4 when the
                 // warp holds one QP: its four groups then split the stages instead of mirroring each other)
  int ro[4];     // my row of a swizzled block: offset of logical chunk j
  int co[4];     // my column of a swizzled block: offset inside row rr is co[rr >> 1]
  bool xl, ul;   // state lane / input lane (neither: idle lane)
  int islot;     // first single-variable-row slot of my variable inside a stage (t adds 1)
  uint64_t eqm, loosem;  // planner: bit k = my box row at stage k is an equality / both bounds infinite
  // Helper-warp kernels (one QP per CTA): EVERY warp of the CTA walks the cold per-stage loops -- warp w, group g takes the
  // stages 4 w + g, + 4 nw, ... -- so a barrier between two loops is a CTA barrier and a reduction goes through `red`
  // (16 doubles per warp in shared memory).  nw = 1: the warp is on its own (every other kernel).
  int nw;
  double *red;

  __device__ __forceinline__ void sync() const { if (nw > 1) __syncthreads(); else __syncwarp(); }
  // v[0 .. NV-1] reduced over the warps of the CTA (max or sum, in warp order: the same value in every warp); the values
  // come in reduced over the warp
  template <int NV, bool SUM>
  __device__ __forceinline__ void cta_reduce(double (&v)[NV]) const {
    static_assert(NV <= 16, "16 reduction slots per warp");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) red[w * 16 + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double acc = red[i];
      for (int ww = 1; ww < nw; ++ww) { const double o = red[ww * 16 + i]; acc = SUM ? acc + o : ((o > acc) ? o : acc); }
      v[i] = acc;
    }
    __syncthreads();
  }
  __device__ __forceinline__ double rmax(double v) const {
    v = h8::wmax(v, ks > 1);
    if (nw > 1) { double a[1] = {v}; cta_reduce<1, false>(a); v = a[0]; }
    return v;
  }
  __device__ __forceinline__ double rsum(double v) const {
    v = h8::wsum(v, ks > 1);
    if (nw > 1) { double a[1] = {v}; cta_reduce<1, true>(a); v = a[0]; }
    return v;
  }
  __device__ __forceinline__ int rany(int v) const {
    v = h8::wany(v, ks > 1);
    if (nw > 1) v = __syncthreads_or(v);
    return v;
  }

  __device__ __forceinline__ bool var_live(int k) const { return xl || (ul && k < N); }
  __device__ __forceinline__ bool has_in(int k) const {
    if (KIND == LPVMPC_CONTROLLER) return (r == 0 || ul) && k < N;
    return xl || (ul && k < N);
  }
  __device__ __forceinline__ double *cd(int arr) const { return cold + arr * (N + 1) * 8; }
  __device__ __forceinline__ int cG() const { return C_COUNT * (N + 1) * 8; }
  __device__ __forceinline__ int cTK() const { return cG() + N * NX * 8; }
  __device__ __forceinline__ double *Tb(int k) const { return S + (k & kmask) * TKS; }
  __device__ __forceinline__ double *Kb(int k) const { return S + ((k - 1) & kmask) * TKS + 64; }
  __device__ __forceinline__ const double *Gb(int k) const { return cold + cG() + k * (NX * 8); }
  __device__ __forceinline__ double *V(int arr) const { return S + oV + arr; }          // element (k, q) at [k*VS + q]
  // single-variable rows in shared memory: z, y, coefficient, upper (and lower: planner) bound of my row t at stage k
  __device__ __forceinline__ double *Ib(int k) const { return S + oI + k * IS; }
  __device__ __forceinline__ double &zi(int k, int t) const { return Ib(k)[(islot + t) * 2]; }
  __device__ __forceinline__ double &yi(int k, int t) const { return Ib(k)[(islot + t) * 2 + 1]; }
  __device__ __forceinline__ double &si(int k, int t) const { return Ib(k)[NSL * 2 + (islot + t) * 2]; }
  __device__ __forceinline__ double &ui(int k, int t) const { return Ib(k)[NSL * 2 + (islot + t) * 2 + 1]; }
  __device__ __forceinline__ double &li(int k, int t) const { return Ib(k)[OLI + islot + t]; }   // planner only
  __device__ __forceinline__ double lo_of(int k, int t) const { return (KIND == LPVMPC_PLANNER) ? li(k, t) : -kInfty; }
  __device__ __forceinline__ double &pm(int k, int cu) const { return Ib(k)[OPM + cu]; }        // couples u_{k-1}[cu], u_k[cu]
  __device__ __forceinline__ int ci(int k, int t) const { return k * 8 + islot + t; }           // slab slot of my row
  // row-weight classes of my single-variable row (planner); controller rows are always plain inequalities
  __device__ __forceinline__ double rho_of(int k, double rho, double rho_eq) const {
    if (KIND == LPVMPC_CONTROLLER) return rho;
    return ((eqm >> k) & 1ull) ? rho_eq : (((loosem >> k) & 1ull) ? kRhoMin : rho);
  }
};

// my row / my column of a plain row-major 8-wide block (G in the slab, or in shared memory during setup)
__device__ __forceinline__ double prowdot(const double *blk, int r, const double (&g)[8]) {
  const double2 t0 = ld2(blk + r * 8), t1 = ld2(blk + r * 8 + 2), t2 = ld2(blk + r * 8 + 4), t3 = ld2(blk + r * 8 + 6);
  double a0 = t0.x * g[0], a1 = t1.x * g[2], a2 = t2.x * g[4], a3 = t3.x * g[6];
  a0 = fma(t0.y, g[1], a0); a1 = fma(t1.y, g[3], a1); a2 = fma(t2.y, g[5], a2); a3 = fma(t3.y, g[7], a3);
  return (a0 + a1) + (a2 + a3);
}
template <int NR>
__device__ __forceinline__ double pcoldot(const double *blk, int r, const double (&g)[8]) {
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int rr = 0; rr < NR; rr += 2) {
    a0 = fma(blk[rr * 8 + r], g[rr], a0);
    if (rr + 1 < NR) a1 = fma(blk[(rr + 1) * 8 + r], g[rr + 1], a1);
  }
  return a0 + a1;
}

// ---------------------------------------------------------------- block factorisation (cold)
// Row weights: ADMM -> rho_eq on the dynamics rows, rho / rho_eq / rho_min on the single-variable rows;
// polish -> 1/delta on the active rows (C_ACTD / C_ACTI), 0 elsewhere.  In ADMM mode also stores diag(M) in DG.
struct FW { int polish; double rho, rho_eq, idel; };

template <int KIND>
__device__ __noinline__ void factor(const Ctx<KIND> c, const FW fw, const double sigma) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT;
  const int N = c.N, r = c.r;
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI), *ED = c.cd(C_ED);
  double *DG = c.V(V_DG);
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
    const double edk = c.xl ? ED[k * 8 + r] : 0.0;
    {
      double d = PD[k * 8 + r] + sigma;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const double a = c.si(k, t);
          const double w = fw.polish ? ((ACTI[c.ci(k, t)] != 0.0) ? fw.idel : 0.0) : c.rho_of(k, fw.rho, fw.rho_eq);
          d = fma(w * a, a, d);
        }
      }
      if (!fw.polish) DG[k * VS + r] = rowlive ? d : 1.0;
      if (c.xl) d = fma(wdk * edk, edk, d);
      if (!rowlive) d = 1.0;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) if (cc == r) s[cc] = d;
    }
    if (k < N) {  // next-stage dynamics rows: sum_rr w(k+1, rr) G[rr][r] G[rr][cc]
      const double *g = c.Gb(k);
      const double wn = wd(k + 1);
#pragma unroll
      for (int rr = 0; rr < NX; ++rr) {
        const double col = gshfl(wn, rr) * g[rr * 8 + r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 e = ld2(g + rr * 8 + 2 * j);
          s[2 * j] = fma(col, e.x, s[2 * j]);
          s[2 * j + 1] = fma(col, e.y, s[2 * j + 1]);
        }
      }
    }
    if (k > 0) {
      if (c.xl) {  // my row of S_{k,k-1}
        const double f = wdk * edk;
        const double *gp = c.Gb(k - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(gp + r * 8 + 2 * j); so[2 * j] = f * e.x; so[2 * j + 1] = f * e.y; }
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
      for (int j = 0; j < 4; ++j) st2(Kk + c.ro[j], -kr[2 * j], -kr[2 * j + 1]);   // the sweeps add (-K): stored negated
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
        const double piv = frcp(pr[p]);
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
    if (c.ring) {  // streamed factor: block k-1 = [T_{k-1} | K_k] is final, block N = [T_N | -] after the last stage
      double *dst = c.cold + c.cTK();
      if (k > 0) {
        const double *src = c.Tb(k - 1);
        double *d = dst + (size_t)(k - 1) * TKS;
#pragma unroll
        for (int j = 0; j < 8; ++j) { const double2 e = ld2(src + j * 16 + 2 * r); st2(d + j * 16 + 2 * r, e.x, e.y); }
      }
      if (k == N) {
        const double *src = c.Tb(N);
        double *d = dst + (size_t)N * TKS;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(src + j * 16 + 2 * r); st2(d + j * 16 + 2 * r, e.x, e.y); }
      }
    }
  }
  if (c.ring) {  // the slab copy is read by the async proxy (TMA) from here on, and the ring is overwritten by it
    __threadfence_block();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
  }
}

// ---------------------------------------------------------------- hot loop
// Explicit shared-space accesses (32-bit byte addresses, immediate offsets); every access is a volatile asm with a
// memory clobber, so the compiler keeps them in program order with respect to each other and to the plain accesses of
// the cold code.
template <int OFF = 0>
__device__ __forceinline__ double lds(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF = 0>
__device__ __forceinline__ double2 lds2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF = 0>
__device__ __forceinline__ void sts(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
template <int OFF = 0>
__device__ __forceinline__ void sts2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(x), "d"(y) : "memory");
}

constexpr int TKB = TKS * 8, VB = VS * 8;  // bytes per stage

// ---- streamed factor: ring of stage blocks filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx)
// The sweeps touch the factor in a fixed cycle of 2N+1 stage steps (fwd: blocks 0..N, bwd: blocks N-1..0), so the ring
// is a plain FIFO: the load of step t goes to slot t mod kRing and is issued kAhead steps ahead; every load is consumed
// exactly once (blocks around the turning points are simply fetched again from L2).  Streaming kernels are compiled
// separately (template flag ST), hold ONE QP per warp and one warp per CTA, so every address below is warp-uniform.
constexpr int kRing = 4, kAhead = 2;
struct Strm {
  uint32_t t;        // stage steps consumed so far (slot = t % kRing, mbarrier phase parity = (t / kRing) & 1)
  uint32_t ahead;    // loads in flight beyond step t: 0 (cold) or kAhead
  uint32_t rdy;      // the load of step t was seen complete by the probe at the end of step t-1
};
struct StrmC {
  uint32_t ring;     // shared address of ring slot 0
  uint32_t mbar;     // shared address of the mbarriers (8 bytes per slot)
  const double *gsrc;// slab copy of the factor: block k at + k * TKS
  int N;             // horizon
  bool issuer;       // this lane issues the copies
};
__device__ __forceinline__ void strm_wait(const StrmC &sc, const uint32_t t) {
  const uint32_t bar = sc.mbar + 8u * (t & (kRing - 1)), ph = (t / kRing) & 1u;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(ph) : "memory");
}
// non-blocking probe of the load of step t (its latency, ~40 cycles, hides behind the end of the current stage)
__device__ __forceinline__ uint32_t strm_probe(const StrmC &sc, const uint32_t t) {
  const uint32_t bar = sc.mbar + 8u * (t & (kRing - 1)), ph = (t / kRing) & 1u;
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok;
}
// request factor block `blk` for step t
__device__ __forceinline__ void strm_issue(const StrmC &sc, const uint32_t t, const int blk) {
  if (sc.issuer) {
    const uint32_t slot = t & (kRing - 1), bar = sc.mbar + 8u * slot, dst = sc.ring + slot * (uint32_t)(TKS * 8);
    const double *src = sc.gsrc + (size_t)blk * TKS;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TKS * 8) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(TKS * 8), "r"(bar)
                 : "memory");
  }
}
// Entering forward stage k / backward stage k: waits for its block (unless the probe already saw it), requests the
// block kAhead steps further along the cycle, returns the byte offset of the block inside the ring.
__device__ __forceinline__ uint32_t strm_fwd(const StrmC &sc, Strm &sm, const int k) {
  const uint32_t t = sm.t;
  if (!sm.rdy) strm_wait(sc, t);
  strm_issue(sc, t + kAhead, (k + kAhead <= sc.N) ? k + kAhead : 2 * sc.N - k - kAhead);
  sm.t = t + 1u;
  return (t & (kRing - 1)) * (uint32_t)(TKS * 8);
}
__device__ __forceinline__ uint32_t strm_bwd(const StrmC &sc, Strm &sm, const int k) {
  const uint32_t t = sm.t;
  if (!sm.rdy) strm_wait(sc, t);
  strm_issue(sc, t + kAhead, (k >= kAhead) ? k - kAhead : kAhead - 1 - k);
  sm.t = t + 1u;
  return (t & (kRing - 1)) * (uint32_t)(TKS * 8);
}
__device__ __forceinline__ void strm_post(const StrmC &sc, Strm &sm) { sm.rdy = strm_probe(sc, sm.t); }
// before the first stage step after a (re)factorisation: fill the pipeline with blocks 0 .. kAhead-1
__device__ __forceinline__ void strm_warm(const StrmC &sc, Strm &sm) {
  if (sm.ahead == 0u) {
#pragma unroll
    for (int d = 0; d < kAhead; ++d) strm_issue(sc, sm.t + (uint32_t)d, d);
    sm.ahead = kAhead;
    sm.rdy = 0u;
  }
}
// nothing in flight after this (the loads issued ahead are waited for and dropped)
__device__ __forceinline__ void strm_drain(const StrmC &sc, Strm &sm) {
  for (uint32_t d = 0; d < sm.ahead; ++d) strm_wait(sc, sm.t + d);
  sm.t += sm.ahead;
  sm.ahead = 0u;
  sm.rdy = 0u;
}

// Per-lane shared-space addresses of stage 0 (computed once per QP).
template <int KIND>
struct Hot {
  StrmC sc;         // streamed factor (ST kernels)
  uint32_t tk[4];   // logical 16-byte chunk j of my row of T_0 (the same chunk of K_1 at +512)
  uint32_t kc[4];   // my column of K_1 in row 2q (row 2q+1 at +64)
  uint32_t v;       // my element of the stage vectors: B +0, X +64, R +128, XS +192, DG +256, CR +320
  uint32_t ib;      // my first single-variable row: {z,y} of row t at +16t, {s,u} at +NSL*16 + 16t
  uint32_t il;      // planner: lower bound of my row
  uint32_t pm, pm2; // slew coupling with the previous / the next stage (state lanes: the zero pad, stride 0)
  uint32_t istr, pstr;  // bytes per stage of my single-variable rows / my slew coupling (0: dummy row / zero pad)
  uint32_t gpub, ggat;  // all-gather buffer 0 (buffer 1 at ^256): where I publish, where my group's 64 bytes start
};

__device__ __forceinline__ double dot8(const double2 &a0, const double2 &a1, const double2 &a2, const double2 &a3, const double2 &g0,
                                      const double2 &g1, const double2 &g2, const double2 &g3) {
  double w0 = a0.x * g0.x, w1 = a1.x * g1.x, w2 = a2.x * g2.x, w3 = a3.x * g3.x;
  w0 = fma(a0.y, g0.y, w0); w1 = fma(a1.y, g1.y, w1); w2 = fma(a2.y, g2.y, w2); w3 = fma(a3.y, g3.y, w3);
  return (w0 + w1) + (w2 + w3);
}

// forward sweep: B holds the right-hand side on entry, W_k = T_k v_k on exit.  `gsel` toggles the gather buffer.
template <int KIND, bool ST>
__device__ __forceinline__ void sweep_fwd(const Hot<KIND> &h, Strm &sm, const int N, uint32_t &gsel) {
  uint32_t vb = h.v;
  double v = lds(vb);
  if (ST) strm_warm(h.sc, sm);
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const uint32_t so = ST ? strm_fwd(h.sc, sm, k) : (uint32_t)k * (uint32_t)TKB;
    const uint32_t t0 = h.tk[0] + so, t1 = h.tk[1] + so, t2 = h.tk[2] + so, t3 = h.tk[3] + so;
    sts(h.gpub ^ gsel, v);
    const double2 a0 = lds2(t0), a1 = lds2(t1), a2 = lds2(t2), a3 = lds2(t3);
    const double2 b0 = lds2<512>(t0), b1 = lds2<512>(t1), b2 = lds2<512>(t2), b3 = lds2<512>(t3);
    const double bn = lds<VB>(vb);
    __syncwarp();
    const uint32_t gg = h.ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    if (ST) strm_post(h.sc, sm);   // probe the next step's block while the dot products run
    double c0 = fma(b0.x, g0.x, bn), c1 = b1.x * g1.x, c2 = b2.x * g2.x, c3 = b3.x * g3.x;   // b = row of -K_{k+1}
    c0 = fma(b0.y, g0.y, c0); c1 = fma(b1.y, g1.y, c1); c2 = fma(b2.y, g2.y, c2); c3 = fma(b3.y, g3.y, c3);
    v = (c0 + c1) + (c2 + c3);
    sts(vb, dot8(a0, a1, a2, a3, g0, g1, g2, g3));
    vb += VB; gsel ^= 256u;
  }
  {  // stage N: no K
    const uint32_t so = ST ? strm_fwd(h.sc, sm, N) : (uint32_t)N * (uint32_t)TKB;
    const uint32_t t0 = h.tk[0] + so, t1 = h.tk[1] + so, t2 = h.tk[2] + so, t3 = h.tk[3] + so;
    sts(h.gpub ^ gsel, v);
    const double2 a0 = lds2(t0), a1 = lds2(t1), a2 = lds2(t2), a3 = lds2(t3);
    __syncwarp();
    const uint32_t gg = h.ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    if (ST) strm_post(h.sc, sm);
    sts(vb, dot8(a0, a1, a2, a3, g0, g1, g2, g3));
    gsel ^= 256u;
  }
}

// Forward sweep when the warp holds ONE QP: lane group 1 computes W_k = T_k v_k while the other groups run the chain
// v_{k+1} = b_{k+1} + (-K_{k+1}) v_k (mirrored), with one instruction stream: every lane loads ONE matrix row of the stage
// block (T_k or -K_{k+1}), forms base + row . v_k with the same FMA order as sweep_fwd, and stores its result at the top
// of the next stage (chain lanes: into the gather buffer; group 1: W_k into B_k).
template <int KIND, bool ST>
__device__ __forceinline__ void sweep_fwd_w1(const Hot<KIND> &h, Strm &sm, const int N, uint32_t &gsel, const int g) {
  const bool trole = (g == 1);
  const uint32_t roff = trole ? 0u : 512u;
  const uint32_t ggat = trole ? h.ggat - 16u : h.ggat;   // group 1 reads the chain's column of the gather buffer
  uint32_t vb = h.v, wst = h.v;
  double res = lds(vb);   // chain: v_0 = b_0; group 1: b_0 as well, written back to B_0 unchanged by its first store
  if (ST) strm_warm(h.sc, sm);
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const uint32_t so = (ST ? strm_fwd(h.sc, sm, k) : (uint32_t)k * (uint32_t)TKB) + roff;
    sts(trole ? wst : (h.gpub ^ gsel), res);
    const double2 r0 = lds2(h.tk[0] + so), r1 = lds2(h.tk[1] + so), r2 = lds2(h.tk[2] + so), r3 = lds2(h.tk[3] + so);
    const double bn = (k < N) ? lds<VB>(vb) : 0.0;
    __syncwarp();
    const uint32_t gg = ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    if (ST) strm_post(h.sc, sm);
    double c0 = fma(r0.x, g0.x, trole ? 0.0 : bn), c1 = r1.x * g1.x, c2 = r2.x * g2.x, c3 = r3.x * g3.x;
    c0 = fma(r0.y, g0.y, c0); c1 = fma(r1.y, g1.y, c1); c2 = fma(r2.y, g2.y, c2); c3 = fma(r3.y, g3.y, c3);
    res = (c0 + c1) + (c2 + c3);
    wst = vb; vb += VB; gsel ^= 256u;
  }
  if (trole) sts(wst, res);   // W_N (the chain lanes' last result is meaningless: stage N has no K)
  __syncwarp();
}

// x~_k = W_k - K_{k+1}' x~_{k+1} for one stage: returns my component, refreshes gn with the gathered x~_k
__device__ __forceinline__ double bwd_step(const uint32_t k0, const uint32_t k1, const uint32_t k2, const uint32_t k3, const uint32_t vb,
                                          const uint32_t gpub, const uint32_t ggat, uint32_t &gsel, double (&gn)[8]) {
  const double e0 = lds(k0), e1 = lds<64>(k0), e2 = lds(k1), e3 = lds<64>(k1), e4 = lds(k2), e5 = lds<64>(k2), e6 = lds(k3), e7 = lds<64>(k3);
  const double w = lds(vb);
  double a0 = fma(e0, gn[0], w), a1 = e1 * gn[1];   // e = column of -K_{k+1}
  a0 = fma(e2, gn[2], a0); a1 = fma(e3, gn[3], a1);
  a0 = fma(e4, gn[4], a0); a1 = fma(e5, gn[5], a1);
  a0 = fma(e6, gn[6], a0); a1 = fma(e7, gn[7], a1);
  const double xt = a0 + a1;
  sts(gpub ^ gsel, xt);
  return xt;
}
__device__ __forceinline__ void gather_in(const uint32_t ggat, uint32_t &gsel, double (&gn)[8]) {
  __syncwarp();
  const uint32_t gg = ggat ^ gsel;
  const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
  gn[0] = g0.x; gn[1] = g0.y; gn[2] = g1.x; gn[3] = g1.y; gn[4] = g2.x; gn[5] = g2.y; gn[6] = g3.x; gn[7] = g3.y;
  gsel ^= 256u;
}

// backward sweep only (polish): x_k = W_k - K_{k+1}' x_{k+1}, written back into B
template <int KIND, bool ST>
__device__ __forceinline__ void sweep_bwd_plain(const Hot<KIND> &h, Strm &sm, const int N, uint32_t &gsel) {
  double gn[8];
  uint32_t vb = h.v + N * VB;
  {
    const double x = lds(vb);
    sts(h.gpub ^ gsel, x);
    gather_in(h.ggat, gsel, gn);
  }
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    const uint32_t so = ST ? strm_bwd(h.sc, sm, k) : (uint32_t)k * (uint32_t)TKB;
    vb -= VB;
    const double xt = bwd_step(h.kc[0] + so, h.kc[1] + so, h.kc[2] + so, h.kc[3] + so, vb, h.gpub, h.ggat, gsel, gn);
    if (ST) strm_post(h.sc, sm);
    sts(vb, xt);
    gather_in(h.ggat, gsel, gn);
  }
  __syncwarp();
}

// Element-wise ADMM update of my variable at stage j given x~_j (x1), x~_{j-1} (xm) and x~_{j+1} (xp): r recursion,
// z / y of my single-variable rows, relaxed x, next right-hand side, running sum of x~.
// BRANCH FREE: lanes without a live variable work on all-zero vector slots, lanes without single-variable rows (and
// every lane at a stage without rows) on a dummy row {z = y = s = 0, u = +inf} that reproduces itself, state lanes read
// a zero slew coupling.  Split in two parts so that the first fills the publish -> gather latency of the sweep.
template <int KIND>
struct Upd {
  double rho, rho_eq, rinv, rinv_eq, sigma, alpha, oma, cc;
  uint64_t eqm, loosem;
  bool live;
  int N;
};
struct UpdMid { double xo, rr, xs, cr, m, sold, snew; };
struct UpdIn { double xo, rr, xs, cr, dg, pm, pp, li; double2 zy0, su0, zy1, su1; };

// operands of the update of stage j: issued ahead of the sweep's dependent chain so that they arrive behind it
template <int KIND>
__device__ __forceinline__ void update_loads(const uint32_t vj, const uint32_t ij, const uint32_t lj, const uint32_t pj, const uint32_t pj2,
                                             UpdIn &in) {
  constexpr int NT = Dims<KIND>::NT, NSL = Dims<KIND>::NSL;
  in.xo = lds<64>(vj); in.rr = lds<128>(vj); in.xs = lds<192>(vj); in.dg = lds<256>(vj); in.cr = lds<320>(vj);
  in.pm = lds(pj); in.pp = lds(pj2);
  in.zy0 = lds2(ij); in.su0 = lds2<NSL * 16>(ij);
  in.zy1 = make_double2(0.0, 0.0); in.su1 = make_double2(0.0, 0.0);
  if (NT > 1) { in.zy1 = lds2<16>(ij); in.su1 = lds2<NSL * 16 + 16>(ij); }
  in.li = 0.0;
  if (KIND == LPVMPC_PLANNER) in.li = lds(lj);
}
template <int KIND>
__device__ __forceinline__ void update_part1(const Upd<KIND> &u, const int j, const uint32_t ij, const UpdIn &in, const double x1,
                                             const double xm, const double xp, UpdMid &q) {
  constexpr int NT = Dims<KIND>::NT;
  q.xo = in.xo; q.rr = in.rr; q.xs = in.xs; q.cr = in.cr;
  double m = in.dg * x1;
  m = fma(in.pm, xm, m);
  q.m = fma(in.pp, xp, m);
  double rt = u.rho, ri = u.rinv;
  if (KIND == LPVMPC_PLANNER) {
    const bool eq = (u.eqm >> j) & 1ull, lo = (u.loosem >> j) & 1ull;
    rt = eq ? u.rho_eq : (lo ? kRhoMin : u.rho);
    ri = eq ? u.rinv_eq : (lo ? 1.0 / kRhoMin : u.rinv);
  }
  double sold, snew;
  {
    sold = in.su0.x * fma(rt, in.zy0.x, -in.zy0.y);
    const double zr = fma(u.alpha, in.su0.x * x1, u.oma * in.zy0.x);
    double zn = fma(ri, in.zy0.y, zr);
    if (KIND == LPVMPC_PLANNER) zn = (zn > in.li) ? zn : in.li;
    // controller rows: the lower bound is -OSQP_INFTY (times the row scaling): below any iterate, no clamp needed
    zn = (zn < in.su0.y) ? zn : in.su0.y;
    const double yn = fma(rt, zr - zn, in.zy0.y);
    if (u.live) sts2(ij, zn, yn);
    snew = in.su0.x * fma(rt, zn, -yn);
  }
  if (NT > 1) {
    sold = fma(in.su1.x, fma(rt, in.zy1.x, -in.zy1.y), sold);
    const double zr = fma(u.alpha, in.su1.x * x1, u.oma * in.zy1.x);
    double zn = fma(ri, in.zy1.y, zr);
    zn = (zn < in.su1.y) ? zn : in.su1.y;
    const double yn = fma(rt, zr - zn, in.zy1.y);
    if (u.live) sts2<16>(ij, zn, yn);
    snew = fma(in.su1.x, fma(rt, zn, -yn), snew);
  }
  q.sold = sold; q.snew = snew;
}
template <int KIND>
__device__ __forceinline__ void update_part2(const Upd<KIND> &u, const uint32_t vj, const double x1, const UpdMid &q) {
  const double hh = fma(u.sigma, q.xo, q.rr) + q.sold;            // the right-hand side this x~ was solved for
  const double rn = fma(-u.alpha, hh - q.m, fma(u.cc, q.cr, q.rr));
  const double xn = fma(u.alpha, x1, u.oma * q.xo);
  if (u.live) sts<64>(vj, xn);
  sts<128>(vj, rn);
  sts<192>(vj, q.xs + x1);
  sts(vj, fma(u.sigma, xn, rn) + q.snew);
}

// backward sweep fused with the element-wise update of stage k+1 (hot)
template <int KIND, bool ST>
__device__ __forceinline__ void sweep_bwd_admm(const Hot<KIND> &h, Strm &sm, const Upd<KIND> &u, uint32_t &gsel) {
  const int N = u.N;
  double gn[8];
  uint32_t vb = h.v + N * VB, ib = h.ib + N * h.istr, il = h.il + N * h.istr, pb = h.pm + N * h.pstr, pb2 = h.pm2 + N * h.pstr;
  double x1 = lds(vb), x2 = 0.0;   // stage N: x~_N = W_N
  sts(h.gpub ^ gsel, x1);
  gather_in(h.ggat, gsel, gn);
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    UpdIn in;
    update_loads<KIND>(vb, ib, il, pb, pb2, in);   // stage k+1, independent of the chain below
    const uint32_t so = ST ? strm_bwd(h.sc, sm, k) : (uint32_t)k * (uint32_t)TKB;
    const double xt = bwd_step(h.kc[0] + so, h.kc[1] + so, h.kc[2] + so, h.kc[3] + so, vb - VB, h.gpub, h.ggat, gsel, gn);
    if (ST) strm_post(h.sc, sm);
    UpdMid q;
    update_part1<KIND>(u, k + 1, ib, in, x1, xt, x2, q);   // fills the publish -> gather latency
    gather_in(h.ggat, gsel, gn);
    update_part2<KIND>(u, vb, x1, q);
    vb -= VB; ib -= h.istr; il -= h.istr; pb -= h.pstr; pb2 -= h.pstr;
    x2 = x1; x1 = xt;
  }
  {
    UpdIn in;
    update_loads<KIND>(vb, ib, il, pb, pb2, in);
    UpdMid q;
    update_part1<KIND>(u, 0, ib, in, x1, 0.0, x2, q);
    update_part2<KIND>(u, vb, x1, q);
  }
  __syncwarp();
}

// Backward sweep + element-wise updates when the warp holds ONE QP: the chain (x~_k = W_k - K_{k+1}' x~_{k+1}) is serial
// and runs mirrored in all four lane groups, but the updates are independent per stage, so they are issued once per FOUR
// chain steps with group g taking stage jtop - g (the fused loop above issues them once per step with all groups doing
// the same stage).  x~_k is parked in the B slot of its stage between its chain step and its update; the x~ of the
// stage above a batch comes from the previous batch through a shuffle.
template <int KIND, bool ST>
__device__ __forceinline__ void sweep_bwd_admm_w1(const Hot<KIND> &h, Strm &sm, const Upd<KIND> &u, uint32_t &gsel, const int g) {
  const int N = u.N;
  const int lane = threadIdx.x & 31;
  double gn[8];
  {
    const double xN = lds(h.v + N * VB);   // stage N: x~_N = W_N, already in its B slot
    sts(h.gpub ^ gsel, xN);
    gather_in(h.ggat, gsel, gn);
  }
  double xcarry = 0.0;   // x~ of the stage above the batch's top stage
  int k = N - 1;         // next chain stage
  int jtop = N;          // top stage of the next batch of updates
#pragma unroll 1
  while (jtop >= 0) {
#pragma unroll 1
    for (int i = 0; i < 4 && k >= 0; ++i, --k) {
      const uint32_t so = ST ? strm_bwd(h.sc, sm, k) : (uint32_t)k * (uint32_t)TKB;
      const uint32_t vb = h.v + (uint32_t)k * VB;
      const double xt = bwd_step(h.kc[0] + so, h.kc[1] + so, h.kc[2] + so, h.kc[3] + so, vb, h.gpub, h.ggat, gsel, gn);
      if (ST) strm_post(h.sc, sm);
      sts(vb, xt);   // W_k has been consumed: park x~_k here until the update of stage k overwrites it
      gather_in(h.ggat, gsel, gn);
    }
    __syncwarp();
    // stages whose lower neighbour's x~ exists: j >= k + 2, and stage 0 once the chain is done
    const int jmin = (k < 0) ? 0 : k + 2;
    const int jlow = (jtop - 3 > jmin) ? jtop - 3 : jmin;
    const int j = jtop - g;
    double x1 = 0.0;
    if (j >= jlow) {
      const uint32_t vj = h.v + (uint32_t)j * VB, ij = h.ib + (uint32_t)j * h.istr, lj = h.il + (uint32_t)j * h.istr;
      const uint32_t pj = h.pm + (uint32_t)j * h.pstr, pj2 = h.pm2 + (uint32_t)j * h.pstr;
      UpdIn in;
      update_loads<KIND>(vj, ij, lj, pj, pj2, in);
      x1 = lds(vj);
      const double xm = (j > 0) ? lds(vj - VB) : 0.0;
      const double xp = (g > 0) ? lds(vj + VB) : xcarry;
      UpdMid q;
      update_part1<KIND>(u, j, ij, in, x1, xm, xp, q);
      update_part2<KIND>(u, vj, x1, q);
    }
    xcarry = __shfl_sync(kFull, x1, (lane & 7) + 8 * (jtop - jlow));   // x~ of the lowest stage of this batch
    __syncwarp();
    jtop = jlow - 1;
  }
}

// ---------------------------------------------------------------- twisted variants (one QP per warp, even N)
// The block-tridiagonal system is eliminated FROM BOTH ENDS towards the middle stage m = N / 2 (two-sided / "twisted"
// factorisation, the same scheme as lpv_h16t.cuh): lane groups 0 and 2 own the left half (stages 0 .. m-1, ascending),
// groups 1 and 3 the right half (stages N .. m+1, descending); in LOCAL steps j (stage j on the left, N - j on the right)
// both halves run one instruction stream on mirrored addresses, so a sweep is N/2 + 1 dependent stage steps instead of N + 1.
//   left : T_k = S~_k^-1,  K_k = S_{k,k-1} T_{k-1},  S~_k = S_k - K_k S_{k,k-1}'        v_k = b_k - K_k v_{k-1}
//   right: T_k = S^_k^-1,  J_k = S_{k,k+1} T_{k+1},  S^_k = S_k - J_k S_{k,k+1}'        v_k = b_k - J_k v_{k+1}
//   middle: S~_m = S_m - K_m S_{m,m-1}' - J_m S_{m,m+1}',  v_m = b_m - K_m v_{m-1} - J_m v_{m+1},  x_m = T_m v_m
//   back : x_k = T_k v_k - K_{k+1}' x_{k+1} (left),  x_k = T_k v_k - J_{k-1}' x_{k-1} (right)
// Factor slot k holds [T_k | M_k], M_k = -K_{k+1} for k < m and -J_{k-1} for k > m (slot m: T_m only): the multiplier used
// when LEAVING stage k forwards and when ENTERING it backwards, so both sweeps address slot (stage of local step j) in
// both halves.  The middle stage is owned by both halves: identical operands, identical values, the same addresses.
template <int KIND>
__device__ __noinline__ void factor_tw(const Ctx<KIND> c, const FW fw, const double sigma, const int half) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT;
  const int N = c.N, r = c.r, NL = N >> 1;
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI), *ED = c.cd(C_ED);
  double *DG = c.V(V_DG);
  auto wd = [&](int k) -> double {  // weight of my dynamics row (k, r)
    if (!c.xl) return 0.0;
    return fw.polish ? ((ACTD[k * 8 + r] != 0.0) ? fw.idel : 0.0) : fw.rho_eq;
  };
  __syncwarp();
#pragma unroll 1
  for (int j = 0; j <= NL; ++j) {
    const int k = half ? N - j : j;
    const bool mid = (j == NL);
    const int nbk = (k < N) ? NB : NX;
    const bool rowlive = r < nbk;
    double s[8], so[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) { s[cc] = 0.0; so[cc] = 0.0; }
    const double wdk = wd(k);
    const double edk = c.xl ? ED[k * 8 + r] : 0.0;
    const bool diag = !(mid && half);   // the middle stage's own block enters once: through the left half
    {
      double d = PD[k * 8 + r] + sigma;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const double a = c.si(k, t);
          const double w = fw.polish ? ((ACTI[c.ci(k, t)] != 0.0) ? fw.idel : 0.0) : c.rho_of(k, fw.rho, fw.rho_eq);
          d = fma(w * a, a, d);
        }
      }
      if (!fw.polish && diag) DG[k * VS + r] = rowlive ? d : 1.0;
      if (c.xl) d = fma(wdk * edk, edk, d);
      if (!rowlive) d = 1.0;
      if (!diag) d = 0.0;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) if (cc == r) s[cc] = d;
    }
    // weights of the next stage's dynamics rows (every shuffle stays outside the half-dependent branches)
    const double *gk = c.Gb(k < N ? k : N - 1);
    const double wn = (k < N) ? wd(k + 1) : 0.0;
    const double fn = (k < N && c.xl) ? wn * ED[(k + 1) * 8 + r] : 0.0;
#pragma unroll
    for (int rr = 0; rr < NX; ++rr) {  // next-stage dynamics rows: sum_rr w(k+1, rr) G[rr][r] G[rr][cc]
      const double wcol = gshfl(wn, rr);
      if (k < N && diag) {
        const double col = wcol * gk[rr * 8 + r];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 e = ld2(gk + rr * 8 + 2 * q);
          s[2 * q] = fma(col, e.x, s[2 * q]);
          s[2 * q + 1] = fma(col, e.y, s[2 * q + 1]);
        }
      }
    }
    if (j > 0) {
      // my row of the block that couples stage k to the neighbour eliminated before it
      if (!half) {  // left: S_{k,k-1}
        if (c.xl) {
          const double f = wdk * edk;
          const double *gp = c.Gb(k - 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) { const double2 e = ld2(gp + r * 8 + 2 * q); so[2 * q] = f * e.x; so[2 * q + 1] = f * e.y; }
        } else if (c.ul && k < N) {
#pragma unroll
          for (int cc = NX; cc < NB; ++cc) if (cc == r) so[cc] = PO[(k - 1) * 8 + r];
        }
      }
#pragma unroll
      for (int cc = 0; cc < NX; ++cc) {  // right: S_{k,k+1} = S_{k+1,k}': my row is column r of S_{k+1,k}
        const double f = gshfl(fn, cc);
        if (half && c.var_live(k)) so[cc] = f * gk[cc * 8 + r];
      }
      if (half && c.ul && k < N - 1) {
#pragma unroll
        for (int cc = NX; cc < NB; ++cc) if (cc == r) so[cc] = PO[k * 8 + r];
      }
      const int kp = half ? k + 1 : k - 1;
      double *Mk = c.Tb(kp) + 64;
      const double *Tp = c.Tb(kp);
#pragma unroll
      for (int q = 0; q < 4; ++q) st2(Mk + c.ro[q], so[2 * q], so[2 * q + 1]);  // park the block in the multiplier's slot
      __syncwarp();
      double kr[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) kr[cc] = 0.0;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 e = ld2(Tp + chunk(jj, q));
          kr[2 * q] = fma(so[jj], e.x, kr[2 * q]);
          kr[2 * q + 1] = fma(so[jj], e.y, kr[2 * q + 1]);
        }
      }
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        double acc = s[cc];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 e = ld2(Mk + chunk(cc, q));  // block[cc][2q..2q+1]
          acc = fma(-kr[2 * q], e.x, acc);
          acc = fma(-kr[2 * q + 1], e.y, acc);
        }
        s[cc] = acc;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) st2(Mk + c.ro[q], -kr[2 * q], -kr[2 * q + 1]);   // the sweeps add (-K): stored negated
    }
    if (mid) {  // S~_m = (S_m - K_m S_{m,m-1}') + (- J_m S_{m,m+1}'): the halves exchange their terms
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) s[cc] = s[cc] + __shfl_xor_sync(kFull, s[cc], 8);
    }
    if (!rowlive) {
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) s[cc] = (cc == r) ? 1.0 : 0.0;
    }
    // Gauss-Jordan inverse of the pivot block, rows across lanes (no pivoting: the block is SPD)
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      double pr[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) pr[cc] = gshfl(s[cc], p);
      if (p < nbk) {
        const double piv = frcp(pr[p]);
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
    for (int q = 0; q < 4; ++q) st2(Tk + c.ro[q], s[2 * q], s[2 * q + 1]);
    __syncwarp();
  }
}

// Forward sweep: groups 0 / 1 run the chains of their halves (v_next = b_next + M_k v_k), groups 2 / 3 form W_k = T_k v_k of
// their halves one step behind; one instruction stream: every lane loads ONE row of the stage slot (T_k or M_k).
template <int KIND>
__device__ __forceinline__ void sweep_fwd_tw(const Hot<KIND> &h, const int N, uint32_t &gsel, const int g) {
  const int NL = N >> 1;
  const bool half = g & 1, trole = (g >> 1) != 0;
  const uint32_t tstr = half ? (uint32_t)(-TKB) : (uint32_t)TKB, vstr = half ? (uint32_t)(-VB) : (uint32_t)VB;
  const uint32_t ggat = trole ? h.ggat - 32u : h.ggat;   // the T lanes read the chain column of their half
  uint32_t so = (half ? (uint32_t)N * (uint32_t)TKB : 0u) + (trole ? 0u : 512u);
  uint32_t vb = h.v + (half ? (uint32_t)N * (uint32_t)VB : 0u), wst = vb;
  double res = lds(vb);   // chain: v of local step 0 = b; T lanes: b as well, written back unchanged by their first store
  // Software-pipelined: the matrix row and the right-hand side of the NEXT step are requested behind the products of this
  // one and arrive while the sums, the publish and the barrier run (profiles/r4a_*: with the row loads in front of the
  // barrier a step waited 36 of its 146 cycles for them before it could reach the barrier).  What is left is the warp's
  // own shared-memory instruction rate -- one LDS / STS per 8.7 cycles from a single warp whatever its width
  // (profiles/r1_smem_probe.jsonl) -- at 10 such instructions per step.  The last step's prefetch reads the middle stage's
  // slot and is not used.
  double2 r0 = lds2(h.tk[0] + so), r1 = lds2(h.tk[1] + so), r2 = lds2(h.tk[2] + so), r3 = lds2(h.tk[3] + so);
  double bn = lds(vb + vstr);
  if (trole || (half && NL == 1)) bn = 0.0;   // the middle stage's right-hand side enters once: through the left half
H8_TW_PRAGMA
  for (int j = 0; j < NL; ++j) {
    sts(trole ? wst : (h.gpub ^ gsel), res);
    __syncwarp();
    const uint32_t gg = ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    double c0 = fma(r0.x, g0.x, bn), c1 = r1.x * g1.x, c2 = r2.x * g2.x, c3 = r3.x * g3.x;
    c0 = fma(r0.y, g0.y, c0); c1 = fma(r1.y, g1.y, c1); c2 = fma(r2.y, g2.y, c2); c3 = fma(r3.y, g3.y, c3);
    wst = vb; vb += vstr; so += tstr; gsel ^= 256u;
    r0 = lds2(h.tk[0] + so); r1 = lds2(h.tk[1] + so); r2 = lds2(h.tk[2] + so); r3 = lds2(h.tk[3] + so);
    bn = lds(vb + vstr);
    if (trole || (half && j + 1 == NL - 1)) bn = 0.0;
    res = (c0 + c1) + (c2 + c3);
  }
  if (trole) sts(wst, res);   // W of local step NL-1
  const double vm = res + __shfl_xor_sync(kFull, res, 8);   // chain lanes: v_m = (b_m - K_m v_{m-1}) + (- J_m v_{m+1})
  {  // the middle stage: W_m = T_m v_m (every group, same value to the same slot)
    sts(h.gpub ^ gsel, vm);   // T lanes publish into their own, unread column
    const uint32_t sm_ = (uint32_t)NL * (uint32_t)TKB;
    const double2 a0 = lds2(h.tk[0] + sm_), a1 = lds2(h.tk[1] + sm_), a2 = lds2(h.tk[2] + sm_), a3 = lds2(h.tk[3] + sm_);
    __syncwarp();
    const uint32_t gg = ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    sts(vb, dot8(a0, a1, a2, a3, g0, g1, g2, g3));
    gsel ^= 256u;
  }
  __syncwarp();
}

// backward sweep only (polish): x written back into B; groups 2 / 3 mirror groups 0 / 1
template <int KIND>
__device__ __forceinline__ void sweep_bwd_plain_tw(const Hot<KIND> &h, const int N, uint32_t &gsel, const int g) {
  const int NL = N >> 1;
  const bool half = g & 1;
  double gn[8];
  {
    const double x = lds(h.v + (uint32_t)NL * VB);
    sts(h.gpub ^ gsel, x);
    gather_in(h.ggat, gsel, gn);
  }
#pragma unroll 1
  for (int j = NL - 1; j >= 0; --j) {
    const uint32_t k = (uint32_t)(half ? N - j : j), so = k * (uint32_t)TKB, vb = h.v + k * (uint32_t)VB;
    const double xt = bwd_step(h.kc[0] + so, h.kc[1] + so, h.kc[2] + so, h.kc[3] + so, vb, h.gpub, h.ggat, gsel, gn);
    sts(vb, xt);
    gather_in(h.ggat, gsel, gn);
  }
  __syncwarp();
}

// Backward sweep + element-wise updates.  The chains run from the middle outwards (groups 2 / 3 mirror groups 0 / 1), two
// local steps at a time; behind them one update round: group (half, idx) updates the stage of local step jt - idx of its
// half (round jt = NL: both halves update the middle stage, identically).  x~_k is parked in the B slot of its stage
// between its chain step and its update; the x~ of the neighbour TOWARDS THE MIDDLE of an idx-0 stage was overwritten by
// the previous round and comes from that round's idx-1 group through a shuffle.
template <int KIND>
__device__ __forceinline__ void sweep_bwd_admm_tw(const Hot<KIND> &h, const Upd<KIND> &u, uint32_t &gsel, const int g) {
  const int N = u.N, NL = N >> 1;
  const int lane = threadIdx.x & 31;
  const bool half = g & 1;
  const int idx = g >> 1;
  double gn[8];
  {
    const double xm_ = lds(h.v + (uint32_t)NL * VB);   // x~_m = W_m
    sts(h.gpub ^ gsel, xm_);
    gather_in(h.ggat, gsel, gn);
  }
  double carry = 0.0;
  int jc = NL - 1;   // next local step of the chains
  int jt = NL;       // top local step of the next update round
#pragma unroll 1
  while (jt >= 0) {
#pragma unroll 1
    for (int i = 0; i < 2 && jc >= 0; ++i, --jc) {
      const uint32_t k = (uint32_t)(half ? N - jc : jc), so = k * (uint32_t)TKB, vb = h.v + k * (uint32_t)VB;
      const double xt = bwd_step(h.kc[0] + so, h.kc[1] + so, h.kc[2] + so, h.kc[3] + so, vb, h.gpub, h.ggat, gsel, gn);
      sts(vb, xt);   // W_k has been consumed: park x~_k here until the update of stage k overwrites it
      gather_in(h.ggat, gsel, gn);
    }
    __syncwarp();
    const int jl = jt - idx;
    double x1 = 0.0;
    if (jl >= 0) {
      const int k = half ? N - jl : jl;
      const uint32_t vj = h.v + (uint32_t)k * VB, ij = h.ib + (uint32_t)k * h.istr, lj = h.il + (uint32_t)k * h.istr;
      const uint32_t pj = h.pm + (uint32_t)k * h.pstr, pj2 = h.pm2 + (uint32_t)k * h.pstr;
      UpdIn in;
      update_loads<KIND>(vj, ij, lj, pj, pj2, in);
      x1 = lds(vj);
      double xm = (k > 0) ? lds(vj - VB) : 0.0;
      double xp = (k < N) ? lds(vj + VB) : 0.0;
      if (idx == 0 && jl < NL) { if (half) xm = carry; else xp = carry; }
      UpdMid q;
      update_part1<KIND>(u, k, ij, in, x1, xm, xp, q);
      update_part2<KIND>(u, vj, x1, q);
    }
    carry = __shfl_sync(kFull, x1, (lane & 7) + 8 * ((half ? 1 : 0) + 2));   // x~ of my half's idx-1 stage of this round
    __syncwarp();
    jt -= 2;
  }
}

// ---------------------------------------------------------------- helper warps (twisted kernels, one QP per CTA)
// The element-wise ADMM update is independent per stage, yet in the one-warp kernels it sits on the critical path: 11 update
// rounds of 4 stages behind the 20 chain steps of a planner sweep.  Here the CTA has NH more warps: inside an ADMM step
// they do nothing but updates, and in every cold phase (setup, residuals, certificates, re-projection, polish) all
// NH + 1 warps split the stages between them (Ctx::nw, Ctx::sync, Ctx::cta_reduce -- the cold phases are chains of exposed
// L2 round trips per stage, and with few iterations per QP they outweigh the sweeps: profiles/r4a_*).  All warps follow the
// same control flow; what belongs to the chains (factorisation, sweeps) is done by warp 0 between two CTA barriers.
// In an ADMM step the main warp runs the forward sweep and the backward CHAIN (x~_k goes to its own stage vector XT instead of
// being parked in B, so an update never reads a slot another update writes), publishes how far the chain has come, and the
// 4 NH lane groups of the helpers update the stages behind it, in the order in which their neighbours' x~ appear
// (m, m-1, m+1, m-2, ...).  When the chain ends only the last round of updates is still outstanding.
//   main -> helpers: `go` (iteration sequence number), `prog` (x~ known for local
//                    distances 0 .. prog-1 from the middle stage, st.release after the XT stores)
//   helpers -> main: `done` (atom.add.release after a helper's last update of the iteration)
struct HwShared {   // 512 bytes are set aside for it; the CTA's reduction scratch (Ctx::red) follows
  uint32_t go, prog, done, pad;
};
__device__ __forceinline__ uint32_t ld_acq(uint32_t a) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_rel(uint32_t a, uint32_t v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// progress of the backward chain: a plain (relaxed) store.  Every x~ it announces was stored by this warp BEFORE it -- the
// lanes' XT stores, the __syncwarp of the all-gather, then lane 0's store -- and shared-memory accesses of one warp are
// performed in issue order, so a helper that reads the new value reads the new x~ too; a MEMBAR.ALL.CTA per chain step
// (st.release) costs more than the step itself.
__device__ __forceinline__ void st_prog(uint32_t a, uint32_t v) { asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void add_rel(uint32_t a, uint32_t v) {
  asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// Per-lane addresses of a QP region (main warp and helpers alike)
template <int KIND>
__device__ __forceinline__ void init_hot(Hot<KIND> &h, const Lay &L, const uint32_t sq, const uint32_t gbuf, const int N, const int g, const int r,
                                         const int (&ro)[4], const int (&co)[4]) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, OLI = Ctx<KIND>::OLI, OPM = Ctx<KIND>::OPM;
  const bool xl = r < NX, ul = (r >= NX) && (r < NB);
  const int islot = (KIND == LPVMPC_CONTROLLER) ? ((r == 0) ? 0 : ((r >= NX) ? (r - NX + 1) * 2 : 0)) : r;
  const int ucomp = r - NX;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    h.tk[q] = sq + (uint32_t)(L.TK + ro[q]) * 8u;
    h.kc[q] = sq + (uint32_t)(L.TK + 64 + (2 * q) * 8 + co[q]) * 8u;
  }
  h.v = sq + (uint32_t)(L.V + r) * 8u;
  constexpr int ISBk = Ctx<KIND>::IS * 8;
  const bool rows = (KIND == LPVMPC_CONTROLLER) ? (r == 0 || ul) : (xl || ul);
  const uint32_t dm = sq + (uint32_t)(L.I + (N + 1) * Ctx<KIND>::IS) * 8u;   // block N+1: dummy rows, zero coupling
  h.ib = rows ? sq + (uint32_t)(L.I + islot * 2) * 8u : dm;
  h.il = rows ? sq + (uint32_t)(L.I + OLI + islot) * 8u : dm + (uint32_t)OLI * 8u;
  h.istr = rows ? (uint32_t)ISBk : 0u;
  h.pm = ul ? sq + (uint32_t)(L.I + OPM + ucomp) * 8u : dm + (uint32_t)OPM * 8u;
  h.pm2 = ul ? h.pm + (uint32_t)ISBk : h.pm;
  h.pstr = ul ? (uint32_t)ISBk : 0u;
  h.gpub = gbuf + (uint32_t)(64 * (r >> 1) + 16 * g + 8 * (r & 1));
  h.ggat = gbuf + (uint32_t)(16 * g);
}

// main warp: backward chain only; x~ -> XT, progress -> hw.prog; returns when every helper has finished the iteration
template <int KIND>
__device__ __forceinline__ void sweep_bwd_chain_hw(const Hot<KIND> &h, const int N, uint32_t &gsel, const int g, const uint32_t hws, const uint32_t seq,
                                                   const uint32_t nh H8_PH_PARAMS) {
  const int NL = N >> 1;
  const int lane = threadIdx.x & 31;
  const bool half = g & 1;
  double gn[8];
  if (lane == 0) *reinterpret_cast<volatile uint32_t *>(__cvta_shared_to_generic(hws + (uint32_t)offsetof(HwShared, prog))) = 0u;
  __syncwarp();
  {
    const uint32_t vm = h.v + (uint32_t)NL * VB;
    const double xm_ = lds(vm);   // x~_m = W_m
    sts<V_XT * 8>(vm, xm_);
    sts(h.gpub ^ gsel, xm_);
    gather_in(h.ggat, gsel, gn);   // (its __syncwarp orders the XT stores of all lanes before lane 0's release below)
  }
  if (lane == 0) { st_rel(hws + (uint32_t)offsetof(HwShared, go), seq); st_rel(hws + (uint32_t)offsetof(HwShared, prog), 1u); }
  H8_PH(13);
  // software-pipelined: the multiplier column and W of the NEXT chain step are requested a step ahead
  const uint32_t kstr = half ? (uint32_t)TKB : (uint32_t)(-TKB), vstr = half ? (uint32_t)VB : (uint32_t)(-VB);   // towards the ends
  uint32_t so = (uint32_t)(half ? N - (NL - 1) : NL - 1) * (uint32_t)TKB, vb = h.v + (uint32_t)(half ? N - (NL - 1) : NL - 1) * (uint32_t)VB;
  double e0 = lds(h.kc[0] + so), e1 = lds<64>(h.kc[0] + so), e2 = lds(h.kc[1] + so), e3 = lds<64>(h.kc[1] + so);
  double e4 = lds(h.kc[2] + so), e5 = lds<64>(h.kc[2] + so), e6 = lds(h.kc[3] + so), e7 = lds<64>(h.kc[3] + so);
  double w = lds(vb);
H8_TW_PRAGMA
  for (int j = NL - 1; j >= 0; --j) {
    double a0 = fma(e0, gn[0], w), a1 = e1 * gn[1];   // e = column of the (negated) multiplier
    a0 = fma(e2, gn[2], a0); a1 = fma(e3, gn[3], a1);
    a0 = fma(e4, gn[4], a0); a1 = fma(e5, gn[5], a1);
    a0 = fma(e6, gn[6], a0); a1 = fma(e7, gn[7], a1);
    const double xt = a0 + a1;
    sts(h.gpub ^ gsel, xt);
    sts<V_XT * 8>(vb, xt);
    gather_in(h.ggat, gsel, gn);
    // the next step's column goes into the shared-memory pipe behind the all-gather loads; measured equal to requesting it
    // in front of the publish (same-box A/B): the step is bound by the warp's shared-memory instruction rate (8.7 cycles per
    // LDS / STS from one warp, 16 of them here against 10 in the forward sweep -- the column comes in 64-bit pieces), not by
    // the order in the pipe.  The last step re-reads its own column: no branch in the loop body
    if (j > 0) { so += kstr; vb += vstr; }
    e0 = lds(h.kc[0] + so); e1 = lds<64>(h.kc[0] + so); w = lds(vb);
    e2 = lds(h.kc[1] + so); e3 = lds<64>(h.kc[1] + so);
    e4 = lds(h.kc[2] + so); e5 = lds<64>(h.kc[2] + so);
    e6 = lds(h.kc[3] + so); e7 = lds<64>(h.kc[3] + so);
    if (lane == 0) st_prog(hws + (uint32_t)offsetof(HwShared, prog), (uint32_t)(NL - j + 1));
  }
  H8_PH(14);
  const uint32_t want = seq * nh;
  while (ld_acq(hws + (uint32_t)offsetof(HwShared, done)) != want) {}
  __syncwarp();
}

// helper warps, one ADMM iteration: wait for the main warp's `go`, update the stages behind the chain, report
template <int KIND>
__device__ __forceinline__ void helper_iter(const Hot<KIND> &h, const Upd<KIND> &u, const uint32_t hws, const uint32_t seq, const int hw, const int nh) {
  const int lane = threadIdx.x & 31, g = lane >> 3;
  const int N = u.N, NL = N >> 1;
  const int U = 4 * nh;
  while (ld_acq(hws + (uint32_t)offsetof(HwShared, go)) != seq) {}
  // One update is a dependent chain of ~40 fp64 operations (630 cycles per round of 4 stages, measured); two rounds at a
  // time give the warp two independent chains.  t = 0 -> m, 1 -> m-1, 2 -> m+1, 3 -> m-2, ...: the order in which the
  // neighbours' x~ appear.
  auto stage_of = [&](int t) { const int d = (t + 1) >> 1; return (t & 1) ? NL - d : NL + d; };
  auto wait_for = [&](int tl) {   // until the x~ of the neighbours of stage t <= tl are there
    const int dl = (tl + 1) >> 1;
    const uint32_t need = (uint32_t)((dl + 1 < NL ? dl + 1 : NL) + 1);
    while (ld_acq(hws + (uint32_t)offsetof(HwShared, prog)) < need) {}
  };
  int t0 = (hw - 1) * 4;
#if H8_HELPER_TRIPLE
#pragma unroll 1
  for (; nh < H8_HELPER_SINGLES && t0 + 2 * U + 3 <= N; t0 += 3 * U) {   // three full rounds: three independent chains
    wait_for(t0 + 2 * U + 3);
    const int ka = stage_of(t0 + g), kb = stage_of(t0 + U + g), kc = stage_of(t0 + 2 * U + g);
    const uint32_t va = h.v + (uint32_t)ka * VB, ia = h.ib + (uint32_t)ka * h.istr, la = h.il + (uint32_t)ka * h.istr;
    const uint32_t vb = h.v + (uint32_t)kb * VB, ib = h.ib + (uint32_t)kb * h.istr, lb = h.il + (uint32_t)kb * h.istr;
    const uint32_t vc = h.v + (uint32_t)kc * VB, ic = h.ib + (uint32_t)kc * h.istr, lc = h.il + (uint32_t)kc * h.istr;
    UpdIn ina, inb, inc;
    update_loads<KIND>(va, ia, la, h.pm + (uint32_t)ka * h.pstr, h.pm2 + (uint32_t)ka * h.pstr, ina);
    update_loads<KIND>(vb, ib, lb, h.pm + (uint32_t)kb * h.pstr, h.pm2 + (uint32_t)kb * h.pstr, inb);
    update_loads<KIND>(vc, ic, lc, h.pm + (uint32_t)kc * h.pstr, h.pm2 + (uint32_t)kc * h.pstr, inc);
    const double xa1 = lds<V_XT * 8>(va), xam = (ka > 0) ? lds<V_XT * 8>(va - VB) : 0.0, xap = (ka < N) ? lds<V_XT * 8>(va + VB) : 0.0;
    const double xb1 = lds<V_XT * 8>(vb), xbm = (kb > 0) ? lds<V_XT * 8>(vb - VB) : 0.0, xbp = (kb < N) ? lds<V_XT * 8>(vb + VB) : 0.0;
    const double xc1 = lds<V_XT * 8>(vc), xcm = (kc > 0) ? lds<V_XT * 8>(vc - VB) : 0.0, xcp = (kc < N) ? lds<V_XT * 8>(vc + VB) : 0.0;
    UpdMid qa, qb, qc;
    update_part1<KIND>(u, ka, ia, ina, xa1, xam, xap, qa);
    update_part1<KIND>(u, kb, ib, inb, xb1, xbm, xbp, qb);
    update_part1<KIND>(u, kc, ic, inc, xc1, xcm, xcp, qc);
    update_part2<KIND>(u, va, xa1, qa);
    update_part2<KIND>(u, vb, xb1, qb);
    update_part2<KIND>(u, vc, xc1, qc);
  }
#endif
  // (with many helpers -- long horizons, one CTA per SM -- every helper has slack: single rounds, each as soon as its x~ are
  // there, so that after the chain's last step only ONE round is outstanding instead of a pair that waited for its later half)
H8_HELPER_PRAGMA
  for (; nh < H8_HELPER_SINGLES && t0 + U + 3 <= N; t0 += 2 * U) {   // two full rounds
    wait_for(t0 + U + 3);
    const int ka = stage_of(t0 + g), kb = stage_of(t0 + U + g);
    const uint32_t va = h.v + (uint32_t)ka * VB, ia = h.ib + (uint32_t)ka * h.istr, la = h.il + (uint32_t)ka * h.istr;
    const uint32_t vb = h.v + (uint32_t)kb * VB, ib = h.ib + (uint32_t)kb * h.istr, lb = h.il + (uint32_t)kb * h.istr;
    UpdIn ina, inb;
    update_loads<KIND>(va, ia, la, h.pm + (uint32_t)ka * h.pstr, h.pm2 + (uint32_t)ka * h.pstr, ina);
    update_loads<KIND>(vb, ib, lb, h.pm + (uint32_t)kb * h.pstr, h.pm2 + (uint32_t)kb * h.pstr, inb);
    const double xa1 = lds<V_XT * 8>(va), xam = (ka > 0) ? lds<V_XT * 8>(va - VB) : 0.0, xap = (ka < N) ? lds<V_XT * 8>(va + VB) : 0.0;
    const double xb1 = lds<V_XT * 8>(vb), xbm = (kb > 0) ? lds<V_XT * 8>(vb - VB) : 0.0, xbp = (kb < N) ? lds<V_XT * 8>(vb + VB) : 0.0;
    UpdMid qa, qb;
    update_part1<KIND>(u, ka, ia, ina, xa1, xam, xap, qa);
    update_part1<KIND>(u, kb, ib, inb, xb1, xbm, xbp, qb);
    update_part2<KIND>(u, va, xa1, qa);
    update_part2<KIND>(u, vb, xb1, qb);
  }
#pragma unroll 1
  for (; t0 <= N; t0 += U) {   // what is left: single, possibly partial rounds
    wait_for((t0 + 3 <= N) ? t0 + 3 : N);
    const int t = t0 + g;
    if (t <= N) {
      const int k = stage_of(t);
      const uint32_t vj = h.v + (uint32_t)k * VB, ij = h.ib + (uint32_t)k * h.istr, lj = h.il + (uint32_t)k * h.istr;
      const uint32_t pj = h.pm + (uint32_t)k * h.pstr, pj2 = h.pm2 + (uint32_t)k * h.pstr;
      UpdIn in;
      update_loads<KIND>(vj, ij, lj, pj, pj2, in);
      const double x1 = lds<V_XT * 8>(vj);
      const double xm = (k > 0) ? lds<V_XT * 8>(vj - VB) : 0.0;
      const double xp = (k < N) ? lds<V_XT * 8>(vj + VB) : 0.0;
      UpdMid q;
      update_part1<KIND>(u, k, ij, in, x1, xm, xp, q);
      update_part2<KIND>(u, vj, x1, q);
    }
  }
  __syncwarp();
  if (lane == 0) add_rel(hws + (uint32_t)offsetof(HwShared, done), 1u);
}

// ---------------------------------------------------------------- per-QP scalars shared by the cold routines
struct Info {
  double pri_res, dua_res, obj;
  double n_rp, n_z, n_Ax, n_rd, n_q, n_Aty, n_Px;  // scaled-space inf-norms of the last update_info
  double u_z, u_Ax, u_q, u_Aty, u_Px;              // the same, unscaled (termination)
  double csc, cinv;
  int status, unscale;
};

// (P v)_(k, r) for a vector stored [k*vs + r] (diagonal Q, R and the slew-rate coupling of the inputs)
template <int KIND>
__device__ __forceinline__ double rowP(const Ctx<KIND> &c, const double *PD, const double *PO, const double *v, int vs, int k) {
  const int N = c.N, o = k * 8 + c.r, ov = k * vs + c.r;
  double acc = PD[o] * v[ov];
  if (c.ul) {
    if (k > 0 && k < N) acc = fma(PO[o - 8], v[ov - vs], acc);
    if (k < N - 1) acc = fma(PO[o], v[ov + vs], acc);
    if (k == N) acc = 0.0;
  }
  return acc;
}
// (A v) on my dynamics row (k, r): v stored [k*vs + r]
template <int KIND>
__device__ __forceinline__ double rowA_dyn(const Ctx<KIND> &c, const double *ED, const double *v, int vs, int k) {
  double acc = ED[k * 8 + c.r] * v[k * vs + c.r];
  if (k > 0) {
    double g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = v[(k - 1) * vs + q];
    acc += prowdot(c.Gb(k - 1), c.r, g);
  }
  return acc;
}
// (A' t)_(k, r): td on the dynamics rows [k*8 + rr] (slab), ti on the single-variable rows: slab array [k*8 + slot] when
// `tis` is false, the shared-memory duals y when true
template <int KIND>
__device__ __forceinline__ double colA(const Ctx<KIND> &c, const double *ED, const double *td, const double *ti, bool tis, int k) {
  constexpr int NX = Ctx<KIND>::NX, NT = Ctx<KIND>::NT;
  const int o = k * 8 + c.r;
  double acc = c.xl ? ED[o] * td[o] : 0.0;
  if (k < c.N) {
    double g[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? td[(k + 1) * 8 + rr] : 0.0;
    acc += pcoldot<NX>(c.Gb(k), c.r, g);
  }
  if (c.has_in(k)) {
#pragma unroll
    for (int t = 0; t < NT; ++t) acc = fma(c.si(k, t), tis ? c.yi(k, t) : ti[c.ci(k, t)], acc);
  }
  return acc;
}

// Brings the explicit dynamics dual up to date from the running sum of x~ (see the header):
//   y_dyn <- y_dyn + rho_eq (alpha A_dyn XS - (alpha n + (1 - alpha) first) be);  XS <- 0
template <int KIND>
__device__ __noinline__ void sync_yd(const Ctx<KIND> c, const bool live, const double rho_eq, const double alpha, const int n,
                                    const int first) {
  const int N = c.N, r = c.r;
  double *XS = c.V(V_XS);
  double *YD = c.cd(C_YD);
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED);
  const double cb = alpha * n + (first ? (1.0 - alpha) : 0.0);
  if (n > 0) {
    const int rg = c.xl ? r : 0;
H8_COLD_PRAGMA
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, km = k > 0 ? k - 1 : 0;
      // slab operands up front (see update_info); then rowA_dyn(c, ED, XS, VS, k) on them
      const double ed = ldg_v(ED + o), yd = ldg_v(YD + o), be = ldg_v(BE + o);
      const double *gm = c.Gb(km) + rg * 8;
      const double2 t0 = ldg2_v(gm), t1 = ldg2_v(gm + 2), t2 = ldg2_v(gm + 4), t3 = ldg2_v(gm + 6);
      if (c.xl) {
        double ax = ed * XS[k * VS + r];
        if (k > 0) {
          double g[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) g[q] = XS[(k - 1) * VS + q];
          double a0 = t0.x * g[0], a1 = t1.x * g[2], a2 = t2.x * g[4], a3 = t3.x * g[6];
          a0 = fma(t0.y, g[1], a0); a1 = fma(t1.y, g[3], a1); a2 = fma(t2.y, g[5], a2); a3 = fma(t3.y, g[7], a3);
          ax += (a0 + a1) + (a2 + a3);
        }
        if (live) YD[o] = yd + rho_eq * (alpha * ax - cb * be);
      }
    }
    c.sync();
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) XS[k * VS + r] = 0.0;
  }
  c.sync();
}

// Re-projects the recursion state from the explicit iterate:
//   CR = rho_eq A_dyn' be (when `new_cr`);  R = zsel CR - A_dyn' y_dyn - q;  B = sigma x + R + A_in'(rho z - y)
// `doit` guards the groups that keep their state.
template <int KIND>
__device__ __noinline__ void reproject(const Ctx<KIND> c, const bool doit, const double rho, const double rho_eq, const double sigma,
                                      const double zsel, const bool new_cr) {
  constexpr int NT = Ctx<KIND>::NT, NX = Ctx<KIND>::NX;
  const int N = c.N, r = c.r;
  double *BV = c.V(V_B), *R = c.V(V_R), *CR = c.V(V_CR);
  const double *X = c.V(V_X);
  const double *YD = c.cd(C_YD), *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
H8_COLD_PRAGMA
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r, ov = k * VS + r;
    // slab operands of the common path up front (see update_info)
    const int kp = k < N ? k + 1 : N, kg = k < N ? k : N - 1;
    const double ed0 = ldg_v(ED + o), yd0 = ldg_v(YD + o), qv0 = ldg_v(QV + o);
    double gc[NX], ydn[NX];
    {
      const double *gq = c.Gb(kg);
#pragma unroll
      for (int rr = 0; rr < NX; ++rr) { gc[rr] = ldg_v(gq + rr * 8 + r); ydn[rr] = ldg_v(YD + kp * 8 + rr); }
    }
    if (c.var_live(k)) {
      double crd = CR[ov];
      if (new_cr) {
        double acc = c.xl ? ED[o] * BE[o] : 0.0;
        if (k < N) {
          double g[8];
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? BE[(k + 1) * 8 + rr] : 0.0;
          double a0 = 0.0, a1 = 0.0;   // pcoldot<NX>(G_k, r, g) on the column already loaded
#pragma unroll
          for (int rr = 0; rr < NX; rr += 2) {
            a0 = fma(gc[rr], g[rr], a0);
            if (rr + 1 < NX) a1 = fma(gc[rr + 1], g[rr + 1], a1);
          }
          acc += a0 + a1;
        }
        crd = rho_eq * acc;
      }
      double aty = c.xl ? ed0 * yd0 : 0.0;
      if (k < N) {   // pcoldot<NX>(G_k, r, y_dyn of stage k + 1) on the operands above
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int rr = 0; rr < NX; rr += 2) {
          a0 = fma(gc[rr], ydn[rr], a0);
          if (rr + 1 < NX) a1 = fma(gc[rr + 1], ydn[rr + 1], a1);
        }
        aty += a0 + a1;
      }
      double sin = 0.0;
      if (c.has_in(k)) {
        const double rt = c.rho_of(k, rho, rho_eq);
#pragma unroll
        for (int t = 0; t < NT; ++t) sin = fma(c.si(k, t), rt * c.zi(k, t) - c.yi(k, t), sin);
      }
      if (doit) {
        const double rr = (zsel * crd - aty) - qv0;
        CR[ov] = crd; R[ov] = rr;
        BV[ov] = fma(sigma, X[ov], rr) + sin;
      }
    } else if (doit) { CR[ov] = 0.0; R[ov] = 0.0; BV[ov] = 0.0; }
  }
  c.sync();
}

// residual norms at the current iterate (update_info); y_dyn must be in sync.
// Every slab (L2) operand of a stage is requested at the top of the stage's iteration, with clamped indices and no
// condition in front of it: behind the conditions (state lane / row present / k < N ...) the loads sat in different basic
// blocks and a stage cost one exposed L2 round trip per block (six to eight of ~700 cycles; the slab cannot live in L1,
// shared memory takes the SM's array).  The arithmetic is unchanged: the same products in the same order.
template <int KIND>
__device__ __noinline__ void update_info(const Ctx<KIND> c, Info *ip, const double zsel) {
  constexpr int NT = Ctx<KIND>::NT, NX = Ctx<KIND>::NX;
  Info &I = *ip;
  const int N = c.N, r = c.r;
  const double *X = c.V(V_X);
  const double *YD = c.cd(C_YD), *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
  const double *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV), *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double a_rp = 0, a_z = 0, a_Ax = 0, b_rp = 0, b_z = 0, b_Ax = 0;
  double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, b_rd = 0, b_q = 0, b_Aty = 0, b_Px = 0;
  const int rg = c.xl ? r : 0;   // a row of G that exists (only state lanes use theirs)
H8_COLD_PRAGMA
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r, ov = k * VS + r;
    const int km = k > 0 ? k - 1 : 0, kp = k < N ? k + 1 : N, kg = k < N ? k : N - 1;
    // ---- slab operands, up front
    const double ed = ldg_v(ED + o), be = ldg_v(BE + o), einv = ldg_v(EINV + o), pd = ldg_v(PD + o), po0 = ldg_v(PO + o), pom = ldg_v(PO + km * 8 + r);
    const double yd = ldg_v(YD + o), qv = ldg_v(QV + o), dinv = ldg_v(DINV + o);
    const double *gm = c.Gb(km) + rg * 8;
    const double2 t0 = ldg2_v(gm), t1 = ldg2_v(gm + 2), t2 = ldg2_v(gm + 4), t3 = ldg2_v(gm + 6);
    const double *gk = c.Gb(kg);
    double gc[NX], ydn[NX], eiinv[NT];
#pragma unroll
    for (int rr = 0; rr < NX; ++rr) { gc[rr] = ldg_v(gk + rr * 8 + r); ydn[rr] = ldg_v(YD + kp * 8 + rr); }
#pragma unroll
    for (int t = 0; t < NT; ++t) eiinv[t] = ldg_v(EIINV + c.ci(k, t));
    const bool hin = c.has_in(k);
    // ---- the arithmetic of rowA_dyn / rowP / colA on them
    if (c.xl) {
      double Ax = ed * X[ov];
      if (k > 0) {
        double g[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = X[(k - 1) * VS + q];
        double a0 = t0.x * g[0], a1 = t1.x * g[2], a2 = t2.x * g[4], a3 = t3.x * g[6];
        a0 = fma(t0.y, g[1], a0); a1 = fma(t1.y, g[3], a1); a2 = fma(t2.y, g[5], a2); a3 = fma(t3.y, g[7], a3);
        Ax += (a0 + a1) + (a2 + a3);
      }
      const double z = zsel * be, rr = Ax - z, ei = einv;
      a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, Ax);
      b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * Ax);
    }
    if (hin) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const double ax = c.si(k, t) * X[ov], z = c.zi(k, t), rr = ax - z, ei = eiinv[t];
        a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, ax);
        b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * ax);
      }
    }
    if (c.var_live(k)) {
      double Px = pd * X[ov];
      if (c.ul) {
        if (k > 0 && k < N) Px = fma(pom, X[ov - VS], Px);
        if (k < N - 1) Px = fma(po0, X[ov + VS], Px);
        if (k == N) Px = 0.0;
      }
      double Aty = c.xl ? ed * yd : 0.0;
      if (k < N) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int rr = 0; rr < NX; rr += 2) {
          a0 = fma(gc[rr], ydn[rr], a0);
          if (rr + 1 < NX) a1 = fma(gc[rr + 1], ydn[rr + 1], a1);
        }
        Aty += a0 + a1;
      }
      if (hin) {
#pragma unroll
        for (int t = 0; t < NT; ++t) Aty = fma(c.si(k, t), c.yi(k, t), Aty);
      }
      const double rr = (qv + Px) + Aty, di = dinv;
      a_rd = absmax(a_rd, rr); a_q = absmax(a_q, qv); a_Aty = absmax(a_Aty, Aty); a_Px = absmax(a_Px, Px);
      b_rd = absmax(b_rd, di * rr); b_q = absmax(b_q, di * qv); b_Aty = absmax(b_Aty, di * Aty); b_Px = absmax(b_Px, di * Px);
    }
  }
  const bool wd = c.ks > 1;
  if (c.nw > 1) {   // the CTA's warps split the stages: one trip through shared memory for all fourteen norms
    double v[14] = {wmax(a_rp, wd), wmax(a_z, wd), wmax(a_Ax, wd), wmax(a_rd, wd), wmax(a_q, wd), wmax(a_Aty, wd), wmax(a_Px, wd),
                    wmax(b_rp, wd), wmax(b_z, wd), wmax(b_Ax, wd), wmax(b_rd, wd), wmax(b_q, wd), wmax(b_Aty, wd), wmax(b_Px, wd)};
    c.template cta_reduce<14, false>(v);
    I.n_rp = v[0]; I.n_z = v[1]; I.n_Ax = v[2]; I.n_rd = v[3]; I.n_q = v[4]; I.n_Aty = v[5]; I.n_Px = v[6];
    if (I.unscale) { I.pri_res = v[7]; I.u_z = v[8]; I.u_Ax = v[9]; I.dua_res = I.cinv * v[10]; I.u_q = v[11]; I.u_Aty = v[12]; I.u_Px = v[13]; }
    else { I.pri_res = I.n_rp; I.u_z = I.n_z; I.u_Ax = I.n_Ax; I.dua_res = I.n_rd; I.u_q = I.n_q; I.u_Aty = I.n_Aty; I.u_Px = I.n_Px; }
    return;
  }
  I.n_rp = wmax(a_rp, wd); I.n_z = wmax(a_z, wd); I.n_Ax = wmax(a_Ax, wd); I.n_rd = wmax(a_rd, wd); I.n_q = wmax(a_q, wd);
  I.n_Aty = wmax(a_Aty, wd); I.n_Px = wmax(a_Px, wd);
  if (I.unscale) {
    I.pri_res = wmax(b_rp, wd); I.u_z = wmax(b_z, wd); I.u_Ax = wmax(b_Ax, wd);
    I.dua_res = I.cinv * wmax(b_rd, wd); I.u_q = wmax(b_q, wd); I.u_Aty = wmax(b_Aty, wd); I.u_Px = wmax(b_Px, wd);
  } else {
    I.pri_res = I.n_rp; I.u_z = I.n_z; I.u_Ax = I.n_Ax; I.dua_res = I.n_rd; I.u_q = I.n_q; I.u_Aty = I.n_Aty; I.u_Px = I.n_Px;
  }
}

// ---------------------------------------------------------------- infeasibility certificates (rare)
// delta_y / delta_x of the last ADMM step against the iterate saved before it (C_PVX, C_PVYI).  The dynamics part of
// delta_y is rho_eq (alpha A_dyn x~ - c be) with x~ = (x - (1 - alpha) x_prev) / alpha of that step (c = 1 on the very
// first step, alpha afterwards).  The projected delta_y goes to C_DYD / C_PYI (free while ADMM runs).
template <int KIND>
__device__ __noinline__ bool primal_infeasible(const Ctx<KIND> c, const Info *ip, const double eps, const double rho_eq,
                                              const double alpha, const int last_was_first) {
  constexpr int NT = Ctx<KIND>::NT;
  const bool unscale = ip->unscale;
  const int N = c.N, r = c.r;
  const double *X = c.V(V_X);
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED), *PVX = c.cd(C_PVX);
  const double *PVYI = c.cd(C_PVYI), *E = c.cd(C_E), *EI = c.cd(C_EI), *DINV = c.cd(C_DINV);
  double *DYD = c.cd(C_DYD), *DYI = c.cd(C_PYI);
  double *XT = c.V(V_XT);   // [k*VS + q]: the stage vector x~ of the backward sweep, free between two ADMM steps (it was a slab vector: an L2 round trip per read)
  const double ia = 1.0 / alpha, oma = 1.0 - alpha, cb = last_was_first ? 1.0 : alpha;
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) { const int o = k * 8 + r; XT[k * VS + r] = c.var_live(k) ? (X[k * VS + r] - oma * PVX[o]) * ia : 0.0; }
  c.sync();
  double nrm = 0.0, lhs = 0.0;
  const int rg = c.xl ? r : 0;   // a row of G that exists (only state lanes use theirs)
H8_COLD_PRAGMA
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r;
    const int km = k > 0 ? k - 1 : 0;
    // slab operands up front (see update_info)
    const double ed = ldg_v(ED + o), be = ldg_v(BE + o), e = ldg_v(E + o);
    const double *gm = c.Gb(km) + rg * 8;
    const double2 t0 = ldg2_v(gm), t1 = ldg2_v(gm + 2), t2 = ldg2_v(gm + 4), t3 = ldg2_v(gm + 6);
    double pvyi[NT], ei[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) { pvyi[t] = ldg_v(PVYI + c.ci(k, t)); ei[t] = ldg_v(EI + c.ci(k, t)); }
    double d = 0.0;
    if (c.xl) {  // equality rows: no projection.  rowA_dyn(c, ED, XT, VS, k) on the operands above
      double ax = ed * XT[k * VS + r];
      if (k > 0) {
        double g[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = XT[(k - 1) * VS + q];
        double a0 = t0.x * g[0], a1 = t1.x * g[2], a2 = t2.x * g[4], a3 = t3.x * g[6];
        a0 = fma(t0.y, g[1], a0); a1 = fma(t1.y, g[3], a1); a2 = fma(t2.y, g[5], a2); a3 = fma(t3.y, g[7], a3);
        ax += (a0 + a1) + (a2 + a3);
      }
      d = rho_eq * (alpha * ax - cb * be);
    }
    DYD[o] = d;
    if (c.xl) {
      nrm = absmax(nrm, unscale ? e * d : d);
      lhs += be * ((d > 0) ? d : 0) + be * ((d < 0) ? d : 0);
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oc = c.ci(k, t);
        double di = c.yi(k, t) - pvyi[t];
        const double lo = c.lo_of(k, t), up = c.ui(k, t);
        if (up > kInfty * kMinScaling) {
          if (lo < -kInfty * kMinScaling) di = 0.0;
          else di = (di < 0.0) ? di : 0.0;
        } else if (lo < -kInfty * kMinScaling) di = (di > 0.0) ? di : 0.0;
        DYI[oc] = di;
        nrm = absmax(nrm, unscale ? ei[t] * di : di);
        lhs += up * ((di > 0) ? di : 0) + lo * ((di < 0) ? di : 0);
      }
    }
  }
  nrm = c.rmax(nrm);
  lhs = c.rsum(lhs);
  c.sync();
  // the product with A' is only needed when the first two conditions of the certificate hold for some group
  if (!__any_sync(kFull, (nrm > eps) && (lhs < -eps * nrm))) return false;
  double mx = 0.0;
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    if (c.var_live(k)) {
      const double at = colA<KIND>(c, ED, DYD, DYI, false, k);
      mx = absmax(mx, unscale ? DINV[k * 8 + r] * at : at);
    }
  }
  mx = c.rmax(mx);
  return (nrm > eps) && (lhs < -eps * nrm) && (mx < eps * nrm);
}

template <int KIND>
__device__ __noinline__ bool dual_infeasible(const Ctx<KIND> c, const Info *ip, const double eps) {
  constexpr int NT = Ctx<KIND>::NT;
  const bool unscale = ip->unscale;
  const int N = c.N, r = c.r;
  const double *X = c.V(V_X);
  const double *QV = c.cd(C_Q), *ED = c.cd(C_ED);
  const double *PVX = c.cd(C_PVX), *D = c.cd(C_D), *DINV = c.cd(C_DINV), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV);
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double *DX = c.V(V_XT);   // [k*VS + q], as in primal_infeasible (which is done with it: its reductions end in a barrier)
  double nrm = 0.0, qdx = 0.0;
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r;
    const double dx = c.var_live(k) ? X[k * VS + r] - PVX[o] : 0.0;
    DX[k * VS + r] = dx;
    nrm = absmax(nrm, unscale ? D[o] * dx : dx);
    qdx += QV[o] * dx;
  }
  nrm = c.rmax(nrm); qdx = c.rsum(qdx);
  c.sync();
  const double cs = unscale ? ip->csc : 1.0;
  // the products with P and A are only needed when the first two conditions of the certificate hold for some group
  if (!__any_sync(kFull, (nrm > eps) && (qdx < -cs * eps * nrm))) return false;
  double mx = 0.0;
  int viol = 0;
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r;
    if (c.var_live(k)) {
      const double Pdx = rowP<KIND>(c, PD, PO, DX, VS, k);
      mx = absmax(mx, unscale ? DINV[o] * Pdx : Pdx);
    }
    if (c.xl) {
      double v = rowA_dyn<KIND>(c, ED, DX, VS, k);
      if (unscale) v = EINV[o] * v;
      if (v > eps * nrm || v < -eps * nrm) viol = 1;
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        double v = c.si(k, t) * DX[k * VS + r];
        if (unscale) v = EIINV[c.ci(k, t)] * v;
        if (((c.ui(k, t) < kInfty * kMinScaling) && (v > eps * nrm)) || ((c.lo_of(k, t) > -kInfty * kMinScaling) && (v < -eps * nrm))) viol = 1;
      }
    }
  }
  mx = c.rmax(mx);
  viol = c.rany(viol);
  return (nrm > eps) && (qdx < -cs * eps * nrm) && (mx < cs * eps * nrm) && !viol;
}

// returns 1 when a termination status was set for my group (check_termination)
template <int KIND>
__device__ __noinline__ int check_termination(const Ctx<KIND> c, const lpvmpc_settings &S, Info *ip, const bool live, const int approximate,
                                             const double rho_eq, const int last_was_first, const bool have_prev) {
  Info &I = *ip;
  double eps_abs = S.eps_abs, eps_rel = S.eps_rel, eps_pi = S.eps_prim_inf, eps_di = S.eps_dual_inf;
  const bool ncvx = !(I.pri_res <= kInfty) || !(I.dua_res <= kInfty);   // also NaN
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; eps_di *= 10; }
  const double eps_prim = eps_abs + eps_rel * (I.u_z > I.u_Ax ? I.u_z : I.u_Ax);
  const bool prim_ok = I.pri_res < eps_prim;
  double mr = I.u_q; mr = (I.u_Aty > mr) ? I.u_Aty : mr; mr = (I.u_Px > mr) ? I.u_Px : mr;
  if (I.unscale) mr *= I.cinv;
  const double eps_dual = eps_abs + eps_rel * mr;
  const bool dual_ok = I.dua_res < eps_dual;
  bool prim_inf = false, dual_inf = false;
  if (have_prev) {
    if (__any_sync(kFull, live && !ncvx && !prim_ok)) prim_inf = primal_infeasible<KIND>(c, ip, eps_pi, rho_eq, S.alpha, last_was_first) && !prim_ok;
    if (__any_sync(kFull, live && !ncvx && !dual_ok)) dual_inf = dual_infeasible<KIND>(c, ip, eps_di) && !dual_ok;
  }
  if (!live) return 0;
  if (ncvx) { I.status = LPVMPC_NON_CVX; I.obj = nan(""); return 1; }
  if (prim_ok && dual_ok) { I.status = approximate ? LPVMPC_SOLVED_INACCURATE : LPVMPC_SOLVED; return 1; }
  if (prim_inf) { I.status = approximate ? LPVMPC_PRIMAL_INFEASIBLE_INACCURATE : LPVMPC_PRIMAL_INFEASIBLE; I.obj = kInfty; return 1; }
  if (dual_inf) { I.status = approximate ? LPVMPC_DUAL_INFEASIBLE_INACCURATE : LPVMPC_DUAL_INFEASIBLE; I.obj = -kInfty; return 1; }
  return 0;
}

template <int KIND>
__device__ __noinline__ double objective(const Ctx<KIND> c, const double *xv, const int vs, const double scale) {
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *QV = c.cd(C_Q);
  double acc = 0.0;
#pragma unroll 1
  for (int k = c.k0; k <= c.N; k += c.ks)
    if (c.var_live(k)) acc += (0.5 * rowP<KIND>(c, PD, PO, xv, vs, k) + QV[k * 8 + c.r]) * xv[k * vs + c.r];
  return c.rsum(acc) * scale;
}

// ---------------------------------------------------------------- setup: schedule + build + Ruiz (cold, once per QP)
// Works in shared memory (scratch in the factor area and in the R / CR / XS vectors), then moves the cold data to the
// slab.  Returns flags: bit 0 = Curvature() failed, bit 1 = l > u somewhere.
template <int KIND>
__device__ __noinline__ int setup(const Ctx<KIND> c, const H8Params &p, const int b, const bool valid, double *csc_out,
                                 uint64_t *eqm_out, uint64_t *loosem_out) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT, GS = NX * 8;
  const int N = c.N, r = c.r;
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &St = p.S;
  double *S = c.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N, NS8 = (N + 1) * 8;
  (void)nx;
  const int ucomp = r - NX;
  int sched_err = 0, data_err = 0;
  double x0r = 0.0;
  // scratch in the factor area ((N+1)*128 - 64 doubles): 8 arrays of NS8, then G (N x GS, plain rows), then nz doubles
  double *sD = c.ring ? c.cold + c.cTK() : S, *sE = sD + NS8, *sEI = sE + NS8, *sDt = sEI + NS8, *sEt = sDt + NS8, *sEti = sEt + NS8;
  double *sPD = sEti + NS8, *sPO = sPD + NS8;
  double *Gs = sPO + NS8;
  // ---- schedule: G_k = -[A_k B_k] (unscaled), my row
  if (a.sched_mode == LPVMPC_SCHED_GIVEN) {
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      if (c.xl) {
        const double *Ar = a.A + ((size_t)b * N + k) * NX * NX + r * NX, *Br = a.Bm + ((size_t)b * N + k) * NX * 2 + r * 2;
        double row[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) row[cc] = (cc < NX) ? -Ar[cc < NX ? cc : 0] : ((cc < NB) ? -Br[cc - NX < 2 ? cc - NX : 0] : 0.0);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) if (notfinite(row[cc])) data_err = 1;
        double *gk = Gs + k * GS + r * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) st2(gk + 2 * j, row[2 * j], row[2 * j + 1]);
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
        if (notfinite(v)) data_err = 1;   // e.g. vx = 0 or 1 - ey kappa = 0 in the stage matrices
      }
      if (c.xl) {
        double *gk = Gs + k * GS + r * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) st2(gk + 2 * j, -row[2 * j], -row[2 * j + 1]);
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
  c.sync();
  sched_err = gany(sched_err);

  // ---- build (PathFollowingLPVMPC.py:334-348, 397-464; LPV_MPC_Planner.py:145-181)
  double *X = c.V(V_X), *BV = c.V(V_B), *QV = c.V(V_R), *BE = c.V(V_CR), *ED = c.V(V_XS);   // q, be, ed: scratch homes [k*VS + r]
  {
    const double Qrr = c.xl ? M.Q[r * NX + r] : 0.0;
    const double Q0r = c.xl ? M.Q[r] : 0.0;
    const double Rcc = c.ul ? M.R[ucomp * 2 + ucomp] : 0.0;
    const double dRc = c.ul ? M.dR[ucomp] : 0.0;
    const double uold = (c.ul && a.u_old) ? a.u_old[(size_t)b * 2 + ucomp] : 0.0;
    const double mey = (KIND == LPVMPC_PLANNER) ? a.max_ey[b] : 0.0;
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, ov = k * VS + r;
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
      if (notfinite(q) || notfinite(be)) data_err = 1;   // NaN / inf in x0, C, vel_ref or u_old
      sPD[o] = pd; sPO[o] = po; QV[ov] = q; BE[ov] = be; ED[ov] = ed;
      sD[o] = 1.0; sE[o] = 1.0; sEI[o] = 1.0; sEti[o] = 1.0;
      X[ov] = 0.0; BV[ov] = 0.0;
    }
    // every single-variable-row slot starts as a dummy row; block N+1 and the dummy area are all dummies
    {
      constexpr int NSLc = Ctx<KIND>::NSL, OLIc = Ctx<KIND>::OLI, OPMc = Ctx<KIND>::OPM;
#pragma unroll 1
      for (int k = c.k0; k <= N + 1; k += c.ks) {
        double *ibk = c.Ib(k);
        if (r < NSLc) {
          ibk[r * 2] = 0.0; ibk[r * 2 + 1] = 0.0; ibk[NSLc * 2 + r * 2] = 0.0; ibk[NSLc * 2 + r * 2 + 1] = kInfty * kInfty;
          if (KIND == LPVMPC_PLANNER) ibk[OLIc + r] = -kInfty * kInfty;
        }
        if (r < 2) ibk[OPMc + r] = 0.0;
      }
    }
    c.sync();
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
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
          lo = (lo > -kInfty || lo != lo) ? lo : -kInfty;  // python wrapper: l = max(l, -OSQP_INFTY), u = min(u, OSQP_INFTY); NaN stays NaN
          up = (up < kInfty || up != up) ? up : kInfty;
          if (!(lo <= up)) data_err = 1;   // l > u or a NaN bound
          c.si(k, t) = si; c.ui(k, t) = up; c.zi(k, t) = 0.0; c.yi(k, t) = 0.0;
          if (KIND == LPVMPC_PLANNER) c.li(k, t) = lo;
        }
      }
    }
    c.sync();
  }
  data_err = c.rany(data_err);

  // ---- Ruiz equilibration (OSQP scale_data)
  double csc = 1.0;
#pragma unroll 1
  for (int it = 0; it < St.scaling; ++it) {
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, ov = k * VS + r;
      double pa = fabs(sPD[o]);
      if (c.ul) {
        if (k < N - 1) pa = absmax(pa, sPO[o]);
        if (k > 0 && k < N) pa = absmax(pa, sPO[o - 8]);
      }
      double qa = c.xl ? fabs(ED[ov]) : 0.0;
      if (k < N) {
        const double *gk = Gs + k * GS;
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) qa = absmax(qa, gk[rr * 8 + r]);
      }
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          qa = absmax(qa, c.si(k, t));
          sEti[c.ci(k, t)] = frsqrt(limit_scaling(fabs(c.si(k, t))));
        }
      }
      sDt[o] = frsqrt(limit_scaling(pa > qa ? pa : qa));
      double ea = c.xl ? fabs(ED[ov]) : 0.0;
      if (k > 0 && c.xl) {
        const double *gp = Gs + (k - 1) * GS + r * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(gp + 2 * j); ea = absmax(ea, e.x); ea = absmax(ea, e.y); }
      }
      sEt[o] = frsqrt(limit_scaling(ea));
    }
    c.sync();
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, ov = k * VS + r;
      const double dt = sDt[o];
      if (k < N) {
        double *gk = Gs + k * GS;
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) { double *e = gk + rr * 8 + r; *e = (*e * sEt[(k + 1) * 8 + rr]) * dt; }
        if (k < N - 1 && c.ul) sPO[o] = (sPO[o] * dt) * sDt[o + 8];
      }
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oc = c.ci(k, t);
          c.si(k, t) = (c.si(k, t) * sEti[oc]) * dt;
          sEI[oc] = sEI[oc] * sEti[oc];
        }
      }
      if (c.xl) ED[ov] = (ED[ov] * sEt[o]) * dt;
      sPD[o] = (sPD[o] * dt) * dt;
      QV[ov] = dt * QV[ov];
      sD[o] = sD[o] * dt;
      sE[o] = sE[o] * sEt[o];
    }
    c.sync();
    // cost scaling: mean of the column norms of P (summed per lane, then across the group)
    double qn = 0.0, ct = 0.0;
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r;
      double pa = fabs(sPD[o]);
      if (c.ul) {
        if (k < N - 1) pa = absmax(pa, sPO[o]);
        if (k > 0 && k < N) pa = absmax(pa, sPO[o - 8]);
      }
      if (c.var_live(k)) { ct += pa; qn = absmax(qn, QV[k * VS + r]); }
    }
    qn = c.rmax(qn);
    ct = c.rsum(ct) / nz;
    qn = limit_scaling(qn);
    ct = ct > qn ? ct : qn;
    ct = limit_scaling(ct);
    ct = frcp(ct);
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) { const int o = k * 8 + r; sPD[o] *= ct; QV[k * VS + r] *= ct; sPO[o] *= ct; }
    csc *= ct;
    c.sync();
  }
  *csc_out = csc;
  // ---- scaled bounds, constraint classes, cold data to the slab, G to the slab
  {
    double *cD = c.cd(C_D), *cDI = c.cd(C_DINV), *cE = c.cd(C_E), *cEI = c.cd(C_EINV), *cPD = c.cd(C_PD), *cPO = c.cd(C_PO);
    double *cEi = c.cd(C_EI), *cEiI = c.cd(C_EIINV), *cQ = c.cd(C_Q), *cBE = c.cd(C_BE), *cED = c.cd(C_ED), *cYD = c.cd(C_YD);
    uint64_t eqm = 0, loosem = 0;
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, ov = k * VS + r;
      cD[o] = sD[o]; cDI[o] = 1.0 / sD[o]; cE[o] = sE[o]; cEI[o] = 1.0 / sE[o]; cPD[o] = sPD[o]; cPO[o] = sPO[o];
      cEi[o] = sEI[o]; cEiI[o] = 1.0 / sEI[o];
      cQ[o] = c.var_live(k) ? QV[ov] : 0.0; cBE[o] = sE[o] * BE[ov]; cED[o] = ED[ov]; cYD[o] = 0.0;
      if (c.ul) c.pm(k, ucomp) = (k > 0 && k < N) ? sPO[o - 8] : 0.0;   // couples u_{k-1}, u_k
    }
    double *Gd = c.cold + c.cG();
#pragma unroll 1
    for (int k = c.k0; k < N; k += c.ks) {
      if (c.xl) {
        const double *gs = Gs + k * GS + r * 8;
        double *gd = Gd + k * GS + r * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(gs + 2 * j); st2(gd + 2 * j, e.x, e.y); }
      }
    }
    c.sync();
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oc = c.ci(k, t);
          const double up = sEI[oc] * c.ui(k, t);
          c.ui(k, t) = up;
          if (KIND == LPVMPC_PLANNER) {
            const double lo = sEI[oc] * c.li(k, t);
            c.li(k, t) = lo;
            if ((lo < -kInfty * kMinScaling) && (up > kInfty * kMinScaling)) loosem |= 1ull << k;
            else if (up - lo < kRhoTol) eqm |= 1ull << k;
          }
        }
      }
    }
    if (c.ks > 1) {  // the groups split the stages: merge their bit masks (same lane r in every group)
      eqm |= __shfl_xor_sync(kFull, eqm, 8); eqm |= __shfl_xor_sync(kFull, eqm, 16);
      loosem |= __shfl_xor_sync(kFull, loosem, 8); loosem |= __shfl_xor_sync(kFull, loosem, 16);
    }
    if (c.nw > 1) {   // ... and the warps': lane r of a warp's first group leaves its masks in the reduction scratch
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      uint64_t *mr = reinterpret_cast<uint64_t *>(c.red);
      if (lane < 8) { mr[w * 16 + lane] = eqm; mr[w * 16 + 8 + lane] = loosem; }
      __syncthreads();
      eqm = 0; loosem = 0;
      for (int ww = 0; ww < c.nw; ++ww) { eqm |= mr[ww * 16 + r]; loosem |= mr[ww * 16 + 8 + r]; }
      __syncthreads();
    }
    *eqm_out = eqm; *loosem_out = loosem;
    c.sync();
    // the scratch homes become hot vectors: XS = 0 (R, CR, B are set by refresh, DG by factor)
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) c.V(V_XS)[k * VS + r] = 0.0;
    c.sync();
  }
  return (sched_err ? 1 : 0) | (data_err ? 2 : 0);
}

// ---------------------------------------------------------------- polish (cold, once per QP)
// Reduced KKT system of the active rows solved in condensed form (row weights 1/delta) with iterative refinement.
// The work vectors sit in the stage vectors ADMM no longer needs (x -> R, y_dyn -> XS, r2_dyn -> DG); on success the
// polished (x, z, y) replace the iterate.  Each refinement step is two passes over the stages around one solve:
//   rhs = -q - P x - A_red'(y - r2 / delta)                                  (one product with the columns of G)
//   dx  = S^-1 rhs;  A dx gives dy = (A dx - r2) / delta and r2 <- r2 - A dx   (one product with the rows of G)
template <int KIND, bool ST, bool TW>
__device__ __noinline__ int polish(const Ctx<KIND> c, const Hot<KIND> h, Strm *smp, const lpvmpc_settings &St, Info *ip, const bool do_pol, uint32_t gsel) {
  Strm &sm = *smp;
  const int g = (threadIdx.x & 31) >> 3;
  constexpr int NT = Ctx<KIND>::NT, NX = Ctx<KIND>::NX;
  Info &I = *ip;
  const bool unscale = I.unscale;
  const int N = c.N, r = c.r;
  double *X = c.V(V_X), *BV = c.V(V_B);
  double *PX = c.V(V_R), *PYD = c.V(V_XS), *R2D = c.V(V_DG);       // [k*VS + q]
  double *YD = c.cd(C_YD);
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
  double *PYI = c.cd(C_PYI), *R2I = c.cd(C_R2I), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV);
  const double delta = St.delta, idel = 1.0 / St.delta;
  // active-set guess (form_Ared): 1 = lower, 2 = upper, 3 = both
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r;
    double ad = 0.0;
    // equality row (z == l == u): lower / upper by the sign of the dual.  Upstream drops the row when the dual is exactly
    // 0.0; round-off makes that impossible there, but y_dyn recovered from the running sum can be exactly 0.0 on the rows
    // of the unweighted state `s`, and the polish system needs every dynamics row: keep it (as "lower").
    if (c.xl && do_pol) ad = (0.0 < YD[o]) ? 2.0 : 1.0;
    ACTD[o] = ad;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        double ai = 0.0;
        if (do_pol) { if (c.zi(k, t) - c.lo_of(k, t) < -c.yi(k, t)) ai += 1.0; if (c.ui(k, t) - c.zi(k, t) < c.yi(k, t)) ai += 2.0; }
        ACTI[c.ci(k, t)] = ai;
      }
    }
  }
  c.sync();
  FW fw; fw.polish = 1; fw.rho = 0.0; fw.rho_eq = 0.0; fw.idel = idel;
  if (ST) strm_drain(h.sc, sm);
  if (TW && c.nw > 1) {   // the chains belong to warp 0
    if (threadIdx.x < 32) factor_tw<KIND>(c, fw, delta, g & 1);
    __syncthreads();
  } else if (TW) factor_tw<KIND>(c, fw, delta, g & 1);
  else factor<KIND>(c, fw, delta);
  auto bred_i = [&](int k, int t) { const double a = ACTI[c.ci(k, t)]; return (a == 1.0 || a == 3.0) ? c.lo_of(k, t) : c.ui(k, t); };
  // (A' t)_(k, r) with t_dyn stored [k*VS + q] and t_in = (ti0, ti1) of my single-variable rows
  auto colAt2 = [&](const double *td, const double ti0, const double ti1, int k) {
    double acc = c.xl ? ED[k * 8 + r] * td[k * VS + r] : 0.0;
    if (k < N) {
      double g[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? td[(k + 1) * VS + rr] : 0.0;
      acc += pcoldot<NX>(c.Gb(k), r, g);
    }
    if (c.has_in(k)) {
      acc = fma(c.si(k, 0), ti0, acc);
      if (NT > 1) acc = fma(c.si(k, NT - 1), ti1, acc);
    }
    return acc;
  };
  auto colAt = [&](const double *td, const double *ti, int k) {
    double t2[2] = {0.0, 0.0};
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) t2[t] = ti[c.ci(k, t)];
    }
    return colAt2(td, t2[0], t2[1], k);
  };
  // Upstream's refinement, pass for pass (see polish() in lpv_h8t.cuh for the scheme): t = (PX, PYD, PYI) is the tentative
  // iterate, DX the x part of the increment of the running K_reg solve, TMP a row temporary.
  double *DX = R2D, *TMP = c.V(V_CR);
  auto solve = [&]() {
    c.sync();
    if (TW && c.nw > 1) {
      if (threadIdx.x < 32) { sweep_fwd_tw<KIND>(h, N, gsel, g); sweep_bwd_plain_tw<KIND>(h, N, gsel, g); }
      __syncthreads();
    } else if (TW) {
      sweep_fwd_tw<KIND>(h, N, gsel, g);
      sweep_bwd_plain_tw<KIND>(h, N, gsel, g);
    } else {
      sweep_fwd<KIND, ST>(h, sm, N, gsel);
      sweep_bwd_plain<KIND, ST>(h, sm, N, gsel);
    }
  };
  auto inner = [&]() {
#pragma unroll 1
    for (int ii = 0; ii < kPolishInner; ++ii) {
#pragma unroll 1
      for (int k = c.k0; k <= N; k += c.ks) {
        const int o = k * 8 + r, ov = k * VS + r;
        double b = 0.0;
        if (c.var_live(k)) b = ((-QV[o] - rowP<KIND>(c, PD, PO, PX, VS, k)) - colAt(PYD, PYI, k)) - delta * DX[ov];
        BV[ov] = b;
      }
      solve();
#pragma unroll 1
      for (int k = c.k0; k <= N; k += c.ks) {
        const int o = k * 8 + r, ov = k * VS + r;
        const double ddx = BV[ov];
        if (c.xl && ACTD[o] != 0.0) PYD[ov] += idel * rowA_dyn<KIND>(c, ED, BV, VS, k);
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) { const int oc = c.ci(k, t); if (ACTI[oc] != 0.0) PYI[oc] += idel * (c.si(k, t) * ddx); }
        }
        PX[ov] += ddx; DX[ov] += ddx;
      }
      c.sync();
    }
  };
  // ---- s_0 = K_reg^-1 (-q, b_red): rhs = -q + A_red'(b_red / delta)
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) TMP[k * VS + r] = (c.xl && ACTD[k * 8 + r] != 0.0) ? idel * BE[k * 8 + r] : 0.0;
  c.sync();
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    double t2[2] = {0.0, 0.0};
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) t2[t] = (ACTI[c.ci(k, t)] != 0.0) ? idel * bred_i(k, t) : 0.0;
    }
    BV[k * VS + r] = c.var_live(k) ? (-QV[k * 8 + r] + colAt2(TMP, t2[0], t2[1], k)) : 0.0;
  }
  solve();
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r, ov = k * VS + r;
    const double xk = BV[ov];
    PX[ov] = xk; DX[ov] = xk;
    PYD[ov] = (c.xl && ACTD[o] != 0.0) ? idel * (rowA_dyn<KIND>(c, ED, BV, VS, k) - BE[o]) : 0.0;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) { const int oc = c.ci(k, t); PYI[oc] = (ACTI[oc] != 0.0) ? idel * (c.si(k, t) * xk - bred_i(k, t)) : 0.0; }
    }
  }
  c.sync();
  inner();
#pragma unroll 1
  for (int it = 0; it < St.polish_refine_iter; ++it) {
    // s_{j+1} = s_j + K_reg^-1 (rhs - K s_j): condensed right-hand side -q - P tx - A'(ty - r2 / delta), r2 = b_red - A tx
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, ov = k * VS + r;
      TMP[ov] = (c.xl && ACTD[o] != 0.0) ? fma(-idel, BE[o] - rowA_dyn<KIND>(c, ED, PX, VS, k), PYD[ov]) : 0.0;
    }
    c.sync();
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, ov = k * VS + r;
      double b = 0.0;
      if (c.var_live(k)) {
        double t2[2] = {0.0, 0.0};
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) { const int oc = c.ci(k, t); t2[t] = (ACTI[oc] != 0.0) ? fma(-idel, bred_i(k, t) - c.si(k, t) * PX[ov], PYI[oc]) : 0.0; }
        }
        b = (-QV[o] - rowP<KIND>(c, PD, PO, PX, VS, k)) - colAt2(TMP, t2[0], t2[1], k);
      }
      BV[ov] = b;
    }
    solve();
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) { const int ov = k * VS + r; const double ddx = BV[ov]; DX[ov] = ddx; PX[ov] += ddx; }
    c.sync();
    // ty += (A ddx - r2) / delta = (A tx_new - b_red) / delta
#pragma unroll 1
    for (int k = c.k0; k <= N; k += c.ks) {
      const int o = k * 8 + r, ov = k * VS + r;
      if (c.xl && ACTD[o] != 0.0) PYD[ov] += idel * (rowA_dyn<KIND>(c, ED, PX, VS, k) - BE[o]);
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) { const int oc = c.ci(k, t); if (ACTI[oc] != 0.0) PYI[oc] += idel * (c.si(k, t) * PX[ov] - bred_i(k, t)); }
      }
    }
    c.sync();
    inner();
  }
  // pol z = A x, normal-cone projection, residuals, acceptance.  Polished z of the single-variable rows -> R2I.
  double a_rp = 0, a_rd = 0;
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r, ov = k * VS + r;
    if (c.xl) {
      const double Ax = rowA_dyn<KIND>(c, ED, PX, VS, k), t = Ax + PYD[ov];
      PYD[ov] = t - BE[o];
      const double rr = Ax - BE[o];
      a_rp = absmax(a_rp, unscale ? EINV[o] * rr : rr);
    } else PYD[ov] = 0.0;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oc = c.ci(k, t);
        const double ax = c.si(k, t) * PX[ov], tt = ax + PYI[oc];
        const double zc = clampd(tt, c.lo_of(k, t), c.ui(k, t));
        R2I[oc] = zc; PYI[oc] = tt - zc;
        const double rr = ax - zc;
        a_rp = absmax(a_rp, unscale ? EIINV[oc] * rr : rr);
      }
    }
  }
  c.sync();
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r;
    if (c.var_live(k)) {
      const double rr = (QV[o] + rowP<KIND>(c, PD, PO, PX, VS, k)) + colAt(PYD, PYI, k);
      a_rd = absmax(a_rd, unscale ? DINV[o] * rr : rr);
    }
  }
  const double pol_pri = c.rmax(a_rp), pol_dua = (unscale ? I.cinv : 1.0) * c.rmax(a_rd);
  const double pol_obj = objective<KIND>(c, PX, VS, St.scaling ? I.cinv : 1.0);
  const bool ok = (pol_pri < I.pri_res && pol_dua < I.dua_res) || (pol_pri < I.pri_res && I.dua_res < 1e-10) ||
                  (pol_dua < I.dua_res && I.pri_res < 1e-10);
  if (!do_pol) return 0;
  if (!ok) return -1;
  I.obj = pol_obj; I.pri_res = pol_pri; I.dua_res = pol_dua;
#pragma unroll 1
  for (int k = c.k0; k <= N; k += c.ks) {
    const int o = k * 8 + r, ov = k * VS + r;
    X[ov] = PX[ov]; YD[o] = PYD[ov];
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) { c.zi(k, t) = R2I[c.ci(k, t)]; c.yi(k, t) = PYI[c.ci(k, t)]; }
    }
  }
  return 1;
}

// ---------------------------------------------------------------- persistent warps, QPW QPs at a time each
constexpr int kSyncEvery = 25;  // y_dyn is brought up to date and r re-projected at least every kSyncEvery steps

// ST: factor streamed from the slab (one QP per warp, one warp per CTA: every staging address is warp-uniform)
// TW: twisted factorisation (one QP per warp, resident factor, even N: checked by the host)
// NH: helper warps (TW only; one QP per CTA: warp 0 runs the chains, warps 1 .. NH the element-wise updates of the ADMM step, all of them the cold phases)
template <int KIND, int QPW, bool ST, bool TW = false, int NH = 0>
__global__ void __launch_bounds__(NH ? 32 * (NH + 1) : 64, 1) lpv_solve_h8_kernel(const __grid_constant__ H8Params p) {
  static_assert(!ST || QPW == 1, "the streamed kernel holds one QP per warp");
  static_assert(!TW || (QPW == 1 && !ST), "the twisted kernel holds one QP per warp with the factor resident");
  static_assert(NH == 0 || TW, "helper warps come with the twisted kernel");
  extern __shared__ __align__(16) double smem[];
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT, NSL = Ctx<KIND>::NSL;
  const int lane = threadIdx.x & 31, warp = (ST || NH) ? 0 : (int)(threadIdx.x >> 5), wpc = (ST || NH) ? 1 : (int)(blockDim.x >> 5);
  const int g = lane >> 3, r = lane & 7;
  const Lay &L = p.L;
  const int N = L.N;
  Ctx<KIND> c;
  c.S = smem; c.cold = p.cold;
  c.oV = L.V; c.oI = L.I; c.ring = L.ring; c.N = N; c.r = r;
  c.kmask = ST ? (kRing - 1) : -1;
  const int wid = NH ? (int)(threadIdx.x >> 5) : 0;   // helper-warp kernels: my warp among the CTA's NH + 1
  c.nw = NH + 1;
  c.k0 = (QPW == 1) ? 4 * wid + g : 0; c.ks = (QPW == 1) ? 4 * (NH + 1) : 1;
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
  double *wsm = smem + (size_t)warp * QPW * L.total;
  const size_t wslot = (size_t)(blockIdx.x * wpc + warp) * QPW;
  // all-gather buffers: two 256-byte buffers per warp, 512-byte aligned, after the QP regions; then 8 mbarriers per QP
  const uint32_t smem_a = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t gbuf0 = (smem_a + (uint32_t)(wpc * QPW * L.total * 8) + 511u) & ~511u;
  const uint32_t gbuf = gbuf0 + (uint32_t)warp * 512u;
  const uint32_t mbar0 = gbuf0 + (uint32_t)wpc * 512u + (uint32_t)(warp * QPW) * 64u;
  uint32_t gsel = 0;
  Strm sm;
  sm.t = 0u; sm.ahead = 0u; sm.rdy = 0u;
  // helper warps: the mailbox sits behind the gather buffers, the reduction scratch of the cold phases behind the mailbox
  const uint32_t hws = gbuf0 + (uint32_t)wpc * 512u + 256u;
  uint32_t hw_seq = 0;
  c.red = NH ? reinterpret_cast<double *>(__cvta_shared_to_generic(hws + 512u)) : nullptr;
  __shared__ unsigned base_s;
  if (NH) {
    if (threadIdx.x < (unsigned)(sizeof(HwShared) / 4)) reinterpret_cast<volatile uint32_t *>(__cvta_shared_to_generic(hws))[threadIdx.x] = 0u;
    __syncthreads();
  }
  if (ST) {  // this warp's mbarriers (one arrival: the issuing lane's expect_tx)
    if (lane < kRing) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8u * (uint32_t)lane) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
  }

#ifdef LPV_H8_PHASE_TIMING
  long long ph[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ph[i] = 0;
  long long ph_t0 = clock64();
#endif
  for (;;) {
    unsigned base = 0;
    if (NH) {   // one QP per CTA (the barriers of the QP's cold phases separate this write from the previous round's reads)
      if (threadIdx.x == 0) base_s = atomicAdd(p.queue, 1u);
      __syncthreads();
      base = base_s;
    } else {
      if (lane == 0) base = atomicAdd(p.queue, (unsigned)QPW);
      base = __shfl_sync(kFull, base, 0);
    }
    if ((int)base >= p.B) break;
    // Groups without a problem of their own (batch tail, or g >= QPW) mirror group 0 exactly: same problem, same
    // shared / scratch region, same values written by the same instruction; only user-visible outputs are guarded.
    const bool valid = (g < QPW) && ((int)(base + g) < p.B);
    const int gq = (QPW == 1) ? 0 : (valid ? g : 0);
    // one QP per warp / CTA: its groups (and warps) split the stages of the output loops like those of every cold loop
    const bool vout = (QPW == 1) ? ((int)base < p.B) : valid;
    const int b = p.perm ? p.perm[(int)base + gq] : (int)base + gq;
    c.S = wsm + gq * L.total;
    c.cold = p.cold + (wslot + gq) * L.cold_total;

    Hot<KIND> h;
    {
      const uint32_t sq = smem_a + (uint32_t)((warp * QPW + gq) * L.total) * 8u;
      h.sc.issuer = ST && lane == 0;
      h.sc.ring = sq + (uint32_t)L.TK * 8u; h.sc.mbar = mbar0;
      h.sc.gsrc = c.cold + L.cTK; h.sc.N = N;
      init_hot<KIND>(h, L, sq, gbuf, N, g, r, c.ro, c.co);
    }

    Info I;
    double csc = 1.0;
    int flags = 0;
    {
      uint64_t eqm = 0, loosem = 0;
      flags = setup<KIND>(c, p, b, valid, &csc, &eqm, &loosem);
      c.eqm = eqm; c.loosem = loosem;
    }
    H8_PH(0);
    I.csc = csc; I.cinv = 1.0 / csc;
    I.unscale = (S.scaling && !S.scaled_termination) ? 1 : 0;
    I.pri_res = 0.0; I.dua_res = 0.0; I.obj = nan("");
    I.n_rp = I.n_z = I.n_Ax = I.n_rd = I.n_q = I.n_Aty = I.n_Px = 0.0;
    I.u_z = I.u_Ax = I.u_q = I.u_Aty = I.u_Px = 0.0;
    I.status = (flags & 1) ? LPVMPC_SCHEDULE_ERROR : ((flags & 2) ? LPVMPC_DATA_ERROR : LPVMPC_UNSOLVED);

    const double sigma = S.sigma, alpha = S.alpha;
    double rho = fmin(fmax(S.rho, kRhoMin), kRhoMax);
    double rho_eq = kRhoEqOverIneq * rho;
    {
      FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0;
      if (NH) { if (wid == 0) factor_tw<KIND>(c, fw, sigma, g & 1); __syncthreads(); }
      else if (TW) factor_tw<KIND>(c, fw, sigma, g & 1);
      else factor<KIND>(c, fw, sigma);
    }
    H8_PH(1);
    reproject<KIND>(c, true, rho, rho_eq, sigma, 0.0, true);
    H8_PH(2);
    bool live = (flags == 0);
    const bool failed = flags != 0;
    int iter_done = 0, rho_updates = 0;
    int adapt_interval = S.adaptive_rho_interval;
    if (S.adaptive_rho && !adapt_interval) adapt_interval = S.check_termination ? 4 * S.check_termination : 100;
    const int ct = S.check_termination, ai = S.adaptive_rho ? adapt_interval : 0;

    Upd<KIND> u;
    u.sigma = sigma; u.alpha = alpha; u.oma = 1.0 - alpha; u.eqm = c.eqm; u.loosem = c.loosem; u.N = N;
    int iter = 0, nsync = 0, first_in = 0;   // steps since the last y_dyn sync; whether step 0 is among them
    double rho_eq_last = rho_eq;             // rho_eq of the last executed step (delta_y of the certificates)
    bool checked_last = false;
    double zsel = 0.0;
    while (iter < S.max_iter && __any_sync(kFull, live)) {
      int stop = S.max_iter;
      if (ct) { const int nxt = (iter / ct + 1) * ct; stop = nxt < stop ? nxt : stop; }
      if (ai) { const int nxt = (iter / ai + 1) * ai; stop = nxt < stop ? nxt : stop; }
      { const int nxt = iter + kSyncEvery; stop = nxt < stop ? nxt : stop; }
      u.rho = rho; u.rho_eq = rho_eq; u.rinv = 1.0 / rho; u.rinv_eq = 1.0 / rho_eq; u.live = live;
      H8_PH(9);
#pragma unroll 1
      for (; iter < stop; ++iter) {
        if (iter == stop - 1) {  // keep the iterate before the last step of the chunk: delta_x, delta_y
          // (helper-warp kernels: the copy is split by stage over the CTA's warps, the updates of the previous step were split
          // differently and so are those of this step -- a barrier on either side)
          if (NH) __syncthreads();
          if (live) {
            double *PVX = c.cd(C_PVX), *PVYI = c.cd(C_PVYI);
            const double *X = c.V(V_X);
#pragma unroll 1
            for (int k = c.k0; k <= N; k += c.ks) {
              PVX[k * 8 + r] = X[k * VS + r];
              if (c.has_in(k)) {
#pragma unroll
                for (int t = 0; t < NT; ++t) PVYI[c.ci(k, t)] = c.yi(k, t);
              }
            }
          }
          if (NH) __syncthreads();
          H8_PH(3);
        }
        u.cc = (iter == 0) ? 2.0 : alpha;
        if (NH) {
          ++hw_seq;
          if (wid == 0) { sweep_fwd_tw<KIND>(h, N, gsel, g); H8_PH(4); sweep_bwd_chain_hw<KIND>(h, N, gsel, g, hws, hw_seq, NH H8_PH_ARGS); H8_PH(5); }
          else helper_iter<KIND>(h, u, hws, hw_seq, wid, NH);
        }
        else if (TW) { sweep_fwd_tw<KIND>(h, N, gsel, g); sweep_bwd_admm_tw<KIND>(h, u, gsel, g); }
        else if (QPW == 1) sweep_fwd_w1<KIND, ST>(h, sm, N, gsel, g);
        else sweep_fwd<KIND, ST>(h, sm, N, gsel);
        if (NH || TW) {}
        else if (QPW == 1) sweep_bwd_admm_w1<KIND, ST>(h, sm, u, gsel, g);
        else sweep_bwd_admm<KIND, ST>(h, sm, u, gsel);
        if (iter == 0) first_in = 1;
        ++nsync;
        zsel = 1.0;
      }
      c.sync();
      H8_PH(9);
      const int last_was_first = (iter == 1);
      rho_eq_last = rho_eq;
      sync_yd<KIND>(c, live, rho_eq, alpha, nsync, first_in);
      H8_PH(6);
      nsync = 0; first_in = 0;
      const bool can_check = ct && (iter % ct == 0);
      const bool can_adapt = ai && (iter % ai == 0);
      bool new_cr = false;
      checked_last = can_check;
      if (can_check || can_adapt) {
        Info J = I;
        update_info<KIND>(c, &J, zsel);
        if (live) { I = J; iter_done = iter; }
        H8_PH(7);
        if (can_check) {
          if (check_termination<KIND>(c, S, &I, live, 0, rho_eq_last, last_was_first, true)) live = false;  // frozen: stores are predicated on `live`
          H8_PH(8);
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
            if (upd) { rho = rho_new; rho_eq = kRhoEqOverIneq * rho; ++rho_updates; }
            FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0;
            c.sync();
            if (ST) strm_drain(h.sc, sm);
            if (NH) { if (wid == 0) factor_tw<KIND>(c, fw, sigma, g & 1); __syncthreads(); }
            else if (TW) factor_tw<KIND>(c, fw, sigma, g & 1);
            else factor<KIND>(c, fw, sigma);
            H8_PH(1);
            new_cr = true;
          }
        }
      }
      // re-project r (and the pending right-hand side) from the explicit iterate: removes the drift of the recursion
      if (iter < S.max_iter && __any_sync(kFull, live)) reproject<KIND>(c, live, rho, rho_eq, sigma, zsel, new_cr);
      H8_PH(2);
    }
    if (!checked_last && __any_sync(kFull, live)) {
      Info J = I;
      update_info<KIND>(c, &J, zsel);
      if (live) { I = J; iter_done = iter; }
      if (check_termination<KIND>(c, S, &I, live, 0, rho_eq_last, iter == 1, iter > 0)) live = false;
    }
    {
      const bool unsolved = (I.status == LPVMPC_UNSOLVED);
      if (__any_sync(kFull, unsolved)) {
        if (!check_termination<KIND>(c, S, &I, unsolved, 1, rho_eq_last, iter == 1, iter > 0) && unsolved) I.status = LPVMPC_MAX_ITER_REACHED;
      }
    }
    const int status = I.status;
    const bool has_sol = !(status == LPVMPC_PRIMAL_INFEASIBLE || status == LPVMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_DUAL_INFEASIBLE || status == LPVMPC_DUAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_NON_CVX || status == LPVMPC_SCHEDULE_ERROR || status == LPVMPC_DATA_ERROR);
    {
      const double o = objective<KIND>(c, c.V(V_X), VS, S.scaling ? I.cinv : 1.0);
      if (has_sol) I.obj = o;
    }
    // row / variable indices in the reference order
    auto ref_dyn = [&](int k) { return (KIND == LPVMPC_CONTROLLER) ? (6 * N + k * NX + r) : (k * NX + r); };
    auto ref_in = [&](int k, int t) {
      if (KIND == LPVMPC_CONTROLLER) return (r == 0) ? (2 * k + t) : (2 * N + 4 * k + 2 * ucomp + t);
      return nx + (c.xl ? (k * NX + r) : (nx + k * 2 + ucomp));
    };
    auto ref_var = [&](int k) { return c.xl ? (k * NX + r) : (nx + k * 2 + ucomp); };
    if (vout && (a.xs || a.zs || a.ys)) {
      const double *X = c.V(V_X);
      const double *YD = c.cd(C_YD), *BE = c.cd(C_BE);
#pragma unroll 1
      for (int k = c.k0; k <= N; k += c.ks) {
        const int o = k * 8 + r;
        if (c.var_live(k) && a.xs) a.xs[(size_t)b * nz + ref_var(k)] = X[k * VS + r];
        if (c.xl) {
          if (a.zs) a.zs[(size_t)b * m + ref_dyn(k)] = (iter > 0 && !failed) ? BE[o] : 0.0;
          if (a.ys) a.ys[(size_t)b * m + ref_dyn(k)] = YD[o];
        }
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (a.zs) a.zs[(size_t)b * m + ref_in(k, t)] = c.zi(k, t);
            if (a.ys) a.ys[(size_t)b * m + ref_in(k, t)] = c.yi(k, t);
          }
        }
      }
    }
    H8_PH(10);
    int polish_status = 0;
    const bool do_pol = S.polish && status == LPVMPC_SOLVED;
    bool polished_sets = false;
    if (__any_sync(kFull, do_pol)) {
      polish_status = polish<KIND, ST, TW>(c, h, &sm, S, &I, do_pol, gsel);
      polished_sets = do_pol;
    }
    H8_PH(11);
    // ---- outputs
    if (vout) {
      const double *X = c.V(V_X);
      const double *YD = c.cd(C_YD), *D = c.cd(C_D), *E = c.cd(C_E), *EI = c.cd(C_EI), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
#pragma unroll 1
      for (int k = c.k0; k <= N; k += c.ks) {
        const int o = k * 8 + r;
        const double v = has_sol ? D[o] * X[k * VS + r] : nan("");
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
            const int oc = c.ci(k, t);
            const size_t q = (size_t)b * m + ref_in(k, t);
            const int act = polished_sets ? (int)ACTI[oc] : 0;
            if (a.y) a.y[q] = has_sol ? I.cinv * (EI[oc] * c.yi(k, t)) : nan("");
            if (a.active_lo) a.active_lo[q] = act & 1;
            if (a.active_up) a.active_up[q] = (act >> 1) & 1;
          }
        }
      }
      if (r == 0 && valid) {
        a.status[b] = status;
        if (a.iters) a.iters[b] = iter_done;
        if (a.rho_updates) a.rho_updates[b] = rho_updates;
        if (a.polish_status) a.polish_status[b] = polish_status;
        if (a.obj) a.obj[b] = I.obj;
        if (a.pri_res) a.pri_res[b] = failed ? nan("") : I.pri_res;
        if (a.dua_res) a.dua_res[b] = failed ? nan("") : I.dua_res;
      }
    }
    if (ST) strm_drain(h.sc, sm);   // nothing may be in flight into the ring when the next QP's setup / factor starts
    c.sync();
    H8_PH(12);
  }
#ifdef LPV_H8_PHASE_TIMING
  if (NH && threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)ph[i]);
  }
#endif
  (void)NSL; (void)NB; (void)hw_seq; (void)wid;
}

}  // namespace h8
}  // namespace lpv
