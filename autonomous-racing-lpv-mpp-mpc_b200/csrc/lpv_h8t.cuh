// "H8T" kernel: 8 lanes per QP, 4 QPs per warp, 4 warps per CTA, one CTA per SM; the block factor lives in TENSOR MEMORY.
//
// Same algorithm and the same ADMM formulation as lpv_h8.cuh (implicit dynamics dual, see there); what changes is where
// the data lives.  Measurements (profiles/r1d_*, tools/tmem_probe.cu, tools/smem_probe.cu): with the factor in shared
// memory only 16 QPs fit an SM, every FMA of a block mat-vec pulls an operand through the 128 B/clk shared-memory pipe,
// and everything else (A_k/B_k, scalings, duals, polish vectors) has to sit in an L2 slab whose latency the single warp
// per SM sub-partition cannot hide (40 % of all stall cycles).
//
// Tensor memory (256 KB per SM, 128 lanes x 512 32-bit columns) is reachable from ordinary warps with tcgen05.st /
// tcgen05.ld .32x32b: thread i of warp w owns lane 32 (w % 4) + i, i.e. 2 KB of private, dynamically indexable storage
// that moves 8 doubles per thread per instruction at 195 B/clk per warp with no bank conflicts and no contention between
// the four warps.  Lane r of a QP group keeps there, for every stage k: row r of T_k, row r of K_k and column r of K_k
// (48 N + 16 columns <= 512, so N <= 10).  The loads of the next stage are issued before the all-gather of the current
// one, so their latency hides behind it.
//
// Shared memory (14 KB per QP, 16 QPs per SM as before) now holds: the 6 stage vectors, the single-variable rows, G
// (A_k / B_k, XOR-swizzled rows), the eight most used "cold" vectors (P diagonal / slew coupling, q, be, ed, y_dyn, 1/D,
// 1/E) and a 128-double scratch used by the factorisation; only the rarely used vectors stay in the L2 slab.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "lpv_qp.cuh"

namespace lpv {
namespace h8t {

constexpr int VS = 48;    // doubles per stage of the stage vectors
enum { V_B = 0, V_X = 8, V_R = 16, V_XS = 24, V_DG = 32, V_CR = 40 };

struct Lay {  // per-QP offsets (doubles); computed on the host (lpvmpc.cu: make_h8t_layout)
  int N, nsl;                  // horizon; single-variable-row slots per stage (6 controller, 7 planner)
  int is;                      // doubles per stage of the single-variable-row block: {z,y} x nsl, {s,u} x nsl, [l x nsl], pm x 2
  int V, I;                    // (N+1) x VS, (N+2) x is (block N+1: dummy rows {z = y = s = 0, u = +inf, l = -inf}, zero coupling)
  int G;                       // N x NX x 8: rows of -[A_k B_k] (scaled), chunks XOR-swizzled
  int CS;                      // C_NSMEM x (N+1) x 8: the cold vectors kept in shared memory
  int FS;                      // 128: factorisation scratch (previous pivot inverse, parked off-diagonal block)
  int total;                   // doubles per QP in shared memory (== 8 mod 16: neighbouring groups 64 B apart mod 128)
  int cold_total;              // doubles per QP slot in the global slab
};
enum { C_PD = 0, C_PO, C_Q, C_BE, C_ED, C_YD, C_DINV, C_EINV, C_NSMEM,                       // shared memory
       C_D = C_NSMEM, C_E, C_EI, C_EIINV, C_PVX, C_PVYI, C_DYD, C_PX, C_PYD, C_PYI, C_R2D, C_R2I, C_ACTD, C_ACTI, C_ZT, C_COUNT };  // slab

struct H8Params {
  Lay L;
  Model M;
  lpvmpc_settings S;
  lpvmpc_args a;
  int B;
  unsigned *queue;
  const int *perm;   // visiting order of the batch (NULL: batch order)
  double *cold;
  int cta_rounds;    // H16T: the CTA's warps take their QPs from the queue TOGETHER (one round = 2 QPs per warp) and meet again
                     // before the next round, so that they run the same phase of the kernel at the same time (instruction cache)
};

template <int KIND> struct Dims;
template <> struct Dims<LPVMPC_CONTROLLER> { static constexpr int NX = 6, NT = 2, NSL = 6, OLI = 24, OPM = 24, IS = 26; };
template <> struct Dims<LPVMPC_PLANNER> { static constexpr int NX = 5, NT = 1, NSL = 7, OLI = 28, OPM = 36, IS = 38; };

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
__device__ __forceinline__ int swz(int rr) { return (rr >> 1) & 3; }
// offset of logical 16-byte chunk j of row rr inside a swizzled 8-wide block
__device__ __forceinline__ int chunk(int rr, int j) { return rr * 8 + ((j ^ swz(rr)) << 1); }

template <int KIND>
struct Ctx {
  static constexpr int NX = Dims<KIND>::NX, NB = NX + 2, NT = Dims<KIND>::NT, NSL = Dims<KIND>::NSL;
  static constexpr int OLI = Dims<KIND>::OLI, OPM = Dims<KIND>::OPM, IS = Dims<KIND>::IS;
  double *S;     // my QP's shared region
  double *cold;  // my QP slot in the global slab
  const Lay *L;
  int N, r;
  int ro[4];     // my row of a swizzled block: offset of logical chunk j
  int co[4];     // my column of a swizzled block: offset inside row rr is co[rr >> 1]
  bool xl, ul;   // state lane / input lane (neither: idle lane)
  int islot;     // first single-variable-row slot of my variable inside a stage (t adds 1)
  uint64_t eqm, loosem;  // planner: bit k = my box row at stage k is an equality / both bounds infinite

  __device__ __forceinline__ bool var_live(int k) const { return xl || (ul && k < N); }
  __device__ __forceinline__ bool has_in(int k) const {
    if (KIND == LPVMPC_CONTROLLER) return (r == 0 || ul) && k < N;
    return xl || (ul && k < N);
  }
  uint32_t tm;   // tensor memory: my lane, column 0
  __device__ __forceinline__ double *cd(int arr) const {
    return (arr < C_NSMEM) ? (S + L->CS + arr * (N + 1) * 8) : (cold + (arr - C_NSMEM) * (N + 1) * 8);
  }
  __device__ __forceinline__ double *F0() const { return S + L->FS; }        // previous pivot inverse (swizzled rows)
  __device__ __forceinline__ double *F1() const { return S + L->FS + 64; }   // parked S_{k,k-1} / K_k (swizzled rows)
  __device__ __forceinline__ double *Gb(int k) const { return S + L->G + k * (NX * 8); }
  // tensor-memory columns of my row of T_k, my row of K_k and my column of K_k (k >= 1)
  __device__ __forceinline__ uint32_t tT(int k) const { return tm + (uint32_t)(k * 16); }
  __device__ __forceinline__ uint32_t tKr(int k) const { return tm + (uint32_t)(16 * (N + 1) + (k - 1) * 16); }
  __device__ __forceinline__ uint32_t tKc(int k) const { return tm + (uint32_t)(16 * (2 * N + 1) + (k - 1) * 16); }
  // polish: dual and residual target of my (up to two) single-variable rows at stage k: 8 columns behind the factor
  // (48 N + 16 + 8 (N + 1) <= 512 for N <= 8)
  __device__ __forceinline__ uint32_t tP(int k) const { return tm + (uint32_t)(48 * N + 16 + 8 * k); }
  // certificates: my component of the iterate saved before the last step of a chunk (2 columns) and of the scaling D
  // (2 columns) at stage k, behind the polish columns (60 N + 28 <= 512 for N <= 8)
  __device__ __forceinline__ uint32_t tQ(int k) const { return tm + (uint32_t)(56 * N + 24 + 4 * k); }
  __device__ __forceinline__ double *V(int arr) const { return S + L->V + arr; }          // element (k, q) at [k*VS + q]
  // single-variable rows in shared memory: z, y, coefficient, upper (and lower: planner) bound of my row t at stage k
  __device__ __forceinline__ double *Ib(int k) const { return S + L->I + k * IS; }
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

// dot product of my row of a swizzled block with a vector gathered in logical order
__device__ __forceinline__ double rowdot(const double *blk, const int (&ro)[4], const double (&g)[8]) {
  const double2 t0 = ld2(blk + ro[0]), t1 = ld2(blk + ro[1]), t2 = ld2(blk + ro[2]), t3 = ld2(blk + ro[3]);
  double a0 = t0.x * g[0], a1 = t1.x * g[2], a2 = t2.x * g[4], a3 = t3.x * g[6];
  a0 = fma(t0.y, g[1], a0); a1 = fma(t1.y, g[3], a1); a2 = fma(t2.y, g[5], a2); a3 = fma(t3.y, g[7], a3);
  return (a0 + a1) + (a2 + a3);
}
// dot product of my column of a swizzled block (rows 0..NR-1) with the vector g
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

// ---------------------------------------------------------------- tensor memory as a per-thread scratchpad
// .32x32b: thread i of the warp accesses TMEM lane (lane field of the address) + i; x16 = 16 consecutive 32-bit columns
// = 8 doubles.  All four instructions are warp-collective (.sync.aligned): call them from converged code only.
__device__ __forceinline__ void tm_st8(uint32_t ta, const double (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta),
               "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
               "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
               "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
               "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7]))
               : "memory");
}
struct TmRow { uint32_t w[16]; };  // 8 doubles as loaded (lo, hi pairs); valid after tm_wait_ld()
__device__ __forceinline__ void tm_ld8(uint32_t ta, TmRow &t) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(t.w[0]), "=r"(t.w[1]), "=r"(t.w[2]), "=r"(t.w[3]), "=r"(t.w[4]), "=r"(t.w[5]), "=r"(t.w[6]), "=r"(t.w[7]),
                 "=r"(t.w[8]), "=r"(t.w[9]), "=r"(t.w[10]), "=r"(t.w[11]), "=r"(t.w[12]), "=r"(t.w[13]), "=r"(t.w[14]), "=r"(t.w[15])
               : "r"(ta) : "memory");
}
// x8 = 8 consecutive 32-bit columns = 4 doubles
__device__ __forceinline__ void tm_st4(uint32_t ta, double v0, double v1, double v2, double v3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "r"(__double2loint(v0)),
               "r"(__double2hiint(v0)), "r"(__double2loint(v1)), "r"(__double2hiint(v1)), "r"(__double2loint(v2)), "r"(__double2hiint(v2)),
               "r"(__double2loint(v3)), "r"(__double2hiint(v3))
               : "memory");
}
// x2 = one double, x4 = two doubles
__device__ __forceinline__ void tm_st1(uint32_t ta, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(ta), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ void tm_ld2(uint32_t ta, double &v0, double &v1) {
  uint32_t w0, w1, w2, w3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(ta) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v0 = __hiloint2double((int)w1, (int)w0); v1 = __hiloint2double((int)w3, (int)w2);
}
struct TmQuad { uint32_t w[8]; };
__device__ __forceinline__ void tm_ld4(uint32_t ta, TmQuad &t) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(t.w[0]), "=r"(t.w[1]), "=r"(t.w[2]), "=r"(t.w[3]), "=r"(t.w[4]), "=r"(t.w[5]), "=r"(t.w[6]), "=r"(t.w[7])
               : "r"(ta) : "memory");
}
__device__ __forceinline__ double tm_getq(const TmQuad &t, int i) { return __hiloint2double((int)t.w[2 * i + 1], (int)t.w[2 * i]); }
__device__ __forceinline__ double tm_get(const TmRow &t, int i) { return __hiloint2double((int)t.w[2 * i + 1], (int)t.w[2 * i]); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- block factorisation (cold)
// Row weights: ADMM -> rho_eq on the dynamics rows, rho / rho_eq / rho_min on the single-variable rows;
// polish -> 1/delta on the active rows (C_ACTD / C_ACTI), 0 elsewhere.  In ADMM mode also stores diag(M) in DG.
// Lane r ends up with row r of T_k, row r of K_k and column r of K_k in tensor memory; the shared-memory scratch holds
// the previous pivot inverse (all lanes read all of its rows) and the block being transposed.
struct FW { int polish; double rho, rho_eq, idel; };

template <int KIND>
__device__ __noinline__ void factor(const Ctx<KIND> c, const FW fw, const double sigma) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT;
  const int N = c.N, r = c.r;
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI), *ED = c.cd(C_ED);
  double *DG = c.V(V_DG);
  double *Tp = c.F0(), *Kk = c.F1();
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
        const double f = wdk * edk;
        const double *gp = c.Gb(k - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(gp + c.ro[j]); so[2 * j] = f * e.x; so[2 * j + 1] = f * e.y; }
      } else if (c.ul && k < N) {
#pragma unroll
        for (int cc = NX; cc < NB; ++cc) if (cc == r) so[cc] = PO[(k - 1) * 8 + r];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) st2(Kk + c.ro[j], so[2 * j], so[2 * j + 1]);  // park S_{k,k-1}
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
      tm_st8(c.tKr(k), kr);                                                      // my row of K_k
#pragma unroll
      for (int j = 0; j < 4; ++j) st2(Kk + c.ro[j], kr[2 * j], kr[2 * j + 1]);   // transpose through the scratch
      __syncwarp();
      double kc[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) kc[rr] = Kk[rr * 8 + c.co[rr >> 1]];
      tm_st8(c.tKc(k), kc);                                                      // my column of K_k
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
    tm_st8(c.tT(k), s);                                                          // my row of T_k
    __syncwarp();  // everybody is done with the previous pivot inverse and with the parked block
#pragma unroll
    for (int j = 0; j < 4; ++j) st2(Tp + c.ro[j], s[2 * j], s[2 * j + 1]);
    __syncwarp();
  }
  tm_wait_st();
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

constexpr int VB = VS * 8;  // bytes per stage of the stage vectors

// Per-lane addresses of stage 0 (computed once per QP).
template <int KIND>
struct Hot {
  uint32_t tT, tKr, tKc;  // tensor memory: my row of T_0, my row of K_1, my column of K_1 (16 columns per stage)
  uint32_t v;       // my element of the stage vectors: B +0, X +64, R +128, XS +192, DG +256, CR +320
  uint32_t ib;      // my first single-variable row: {z,y} of row t at +16t, {s,u} at +NSL*16 + 16t
  uint32_t il;      // planner: lower bound of my row
  uint32_t pm, pm2; // slew coupling with the previous / the next stage (state lanes: the zero pad, stride 0)
  uint32_t istr, pstr;  // bytes per stage of my single-variable rows / my slew coupling (0: dummy row / zero pad)
  uint32_t gpub, ggat;  // all-gather buffer 0 (buffer 1 at ^256): where I publish, where my group's 64 bytes start
};

__device__ __forceinline__ double dot8(const TmRow &a, const double2 &g0, const double2 &g1, const double2 &g2, const double2 &g3) {
  double w0 = tm_get(a, 0) * g0.x, w1 = tm_get(a, 2) * g1.x, w2 = tm_get(a, 4) * g2.x, w3 = tm_get(a, 6) * g3.x;
  w0 = fma(tm_get(a, 1), g0.y, w0); w1 = fma(tm_get(a, 3), g1.y, w1); w2 = fma(tm_get(a, 5), g2.y, w2); w3 = fma(tm_get(a, 7), g3.y, w3);
  return (w0 + w1) + (w2 + w3);
}

// forward sweep: B holds the right-hand side on entry, W_k = T_k v_k on exit.  `gsel` toggles the gather buffer.
// The rows of the next stage are requested from tensor memory before the all-gather of the current one.
template <int KIND>
__device__ __forceinline__ void sweep_fwd(const Hot<KIND> &h, const int N, uint32_t &gsel) {
  uint32_t tt = h.tT, tk = h.tKr, vb = h.v;
  TmRow a, b;
  tm_ld8(tt, a); tm_ld8(tk, b);
  double v = lds(vb);
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    sts(h.gpub ^ gsel, v);
    const double bn = lds<VB>(vb);
    __syncwarp();
    const uint32_t gg = h.ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    tm_wait_ld();
    double c0 = fma(-tm_get(b, 0), g0.x, bn), c1 = -(tm_get(b, 2) * g1.x), c2 = -(tm_get(b, 4) * g2.x), c3 = -(tm_get(b, 6) * g3.x);
    c0 = fma(-tm_get(b, 1), g0.y, c0); c1 = fma(-tm_get(b, 3), g1.y, c1); c2 = fma(-tm_get(b, 5), g2.y, c2); c3 = fma(-tm_get(b, 7), g3.y, c3);
    v = (c0 + c1) + (c2 + c3);
    const double w = dot8(a, g0, g1, g2, g3);
    tt += 16; tk += 16;
    tm_ld8(tt, a); tm_ld8(tk, b);   // next stage (the last K request runs into the column area: harmless, never used)
    sts(vb, w);
    vb += VB; gsel ^= 256u;
  }
  {  // stage N: no K
    sts(h.gpub ^ gsel, v);
    __syncwarp();
    const uint32_t gg = h.ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    tm_wait_ld();
    sts(vb, dot8(a, g0, g1, g2, g3));
    gsel ^= 256u;
  }
}

// x~_k = W_k - K_{k+1}' x~_{k+1}: my component, from my column of K_{k+1} (e, already requested) and the gathered x~_{k+1}
__device__ __forceinline__ double bwd_step(const TmRow &e, const double w, const double (&gn)[8]) {
  double a0 = fma(-tm_get(e, 0), gn[0], w), a1 = -(tm_get(e, 1) * gn[1]);
  a0 = fma(-tm_get(e, 2), gn[2], a0); a1 = fma(-tm_get(e, 3), gn[3], a1);
  a0 = fma(-tm_get(e, 4), gn[4], a0); a1 = fma(-tm_get(e, 5), gn[5], a1);
  a0 = fma(-tm_get(e, 6), gn[6], a0); a1 = fma(-tm_get(e, 7), gn[7], a1);
  return a0 + a1;
}
__device__ __forceinline__ void gather_in(const uint32_t ggat, uint32_t &gsel, double (&gn)[8]) {
  __syncwarp();
  const uint32_t gg = ggat ^ gsel;
  const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
  gn[0] = g0.x; gn[1] = g0.y; gn[2] = g1.x; gn[3] = g1.y; gn[4] = g2.x; gn[5] = g2.y; gn[6] = g3.x; gn[7] = g3.y;
  gsel ^= 256u;
}

// backward sweep only (polish): x_k = W_k - K_{k+1}' x_{k+1}, written back into B
template <int KIND>
__device__ __forceinline__ void sweep_bwd_plain(const Hot<KIND> &h, const int N, uint32_t &gsel) {
  double gn[8];
  uint32_t vb = h.v + N * VB, tc = h.tKc + (uint32_t)((N - 1) * 16);
  TmRow e;
  tm_ld8(tc, e);
  {
    const double x = lds(vb);
    sts(h.gpub ^ gsel, x);
    gather_in(h.ggat, gsel, gn);
  }
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    vb -= VB;
    const double w = lds(vb);
    tm_wait_ld();
    const double xt = bwd_step(e, w, gn);
    tc -= 16;
    tm_ld8(k > 0 ? tc : h.tKc, e);   // column of K_k for the next stage (k = 0: dummy request, keeps the code converged)
    sts(h.gpub ^ gsel, xt);
    sts(vb, xt);
    gather_in(h.ggat, gsel, gn);
  }
  tm_wait_ld();
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
template <int KIND>
__device__ __forceinline__ void sweep_bwd_admm(const Hot<KIND> &h, const Upd<KIND> &u, uint32_t &gsel) {
  const int N = u.N;
  double gn[8];
  uint32_t vb = h.v + N * VB, ib = h.ib + N * h.istr, il = h.il + N * h.istr, pb = h.pm + N * h.pstr, pb2 = h.pm2 + N * h.pstr;
  uint32_t tc = h.tKc + (uint32_t)((N - 1) * 16);
  TmRow e;
  tm_ld8(tc, e);
  double x1 = lds(vb), x2 = 0.0;   // stage N: x~_N = W_N
  sts(h.gpub ^ gsel, x1);
  gather_in(h.ggat, gsel, gn);
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    UpdIn in;
    update_loads<KIND>(vb, ib, il, pb, pb2, in);   // stage k+1, independent of the chain below
    const double w = lds(vb - VB);
    tm_wait_ld();
    const double xt = bwd_step(e, w, gn);
    sts(h.gpub ^ gsel, xt);
    tc -= 16;
    tm_ld8(k > 0 ? tc : h.tKc, e);   // column of K_k for the next stage (k = 0: dummy request, keeps the code converged)
    UpdMid q;
    update_part1<KIND>(u, k + 1, ib, in, x1, xt, x2, q);   // fills the publish -> gather latency
    gather_in(h.ggat, gsel, gn);
    update_part2<KIND>(u, vb, x1, q);
    vb -= VB; ib -= h.istr; il -= h.istr; pb -= h.pstr; pb2 -= h.pstr;
    x2 = x1; x1 = xt;
  }
  tm_wait_ld();
  {
    UpdIn in;
    update_loads<KIND>(vb, ib, il, pb, pb2, in);
    UpdMid q;
    update_part1<KIND>(u, 0, ib, in, x1, 0.0, x2, q);
    update_part2<KIND>(u, vb, x1, q);
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
    acc += rowdot(c.Gb(k - 1), c.ro, g);
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
    acc += coldot<NX>(c.Gb(k), c.co, g);
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
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      if (c.xl) {
        const int o = k * 8 + r;
        const double ax = rowA_dyn<KIND>(c, ED, XS, VS, k);
        if (live) YD[o] = YD[o] + rho_eq * (alpha * ax - cb * BE[o]);
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) XS[k * VS + r] = 0.0;
  }
  __syncwarp();
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
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r, ov = k * VS + r;
    if (c.var_live(k)) {
      const double *gk = c.Gb(k);
      double crd = CR[ov];
      if (new_cr) {
        double acc = c.xl ? ED[o] * BE[o] : 0.0;
        if (k < N) {
          double g[8];
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? BE[(k + 1) * 8 + rr] : 0.0;
          acc += coldot<NX>(gk, c.co, g);
        }
        crd = rho_eq * acc;
      }
      double aty = c.xl ? ED[o] * YD[o] : 0.0;
      if (k < N) {
        double g[8];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? YD[(k + 1) * 8 + rr] : 0.0;
        aty += coldot<NX>(gk, c.co, g);
      }
      double sin = 0.0;
      if (c.has_in(k)) {
        const double rt = c.rho_of(k, rho, rho_eq);
#pragma unroll
        for (int t = 0; t < NT; ++t) sin = fma(c.si(k, t), rt * c.zi(k, t) - c.yi(k, t), sin);
      }
      if (doit) {
        const double rr = (zsel * crd - aty) - QV[o];
        CR[ov] = crd; R[ov] = rr;
        BV[ov] = fma(sigma, X[ov], rr) + sin;
      }
    } else if (doit) { CR[ov] = 0.0; R[ov] = 0.0; BV[ov] = 0.0; }
  }
  __syncwarp();
}

// residual norms at the current iterate (update_info); y_dyn must be in sync
template <int KIND>
__device__ __noinline__ void update_info(const Ctx<KIND> c, Info *ip, const double zsel) {
  constexpr int NT = Ctx<KIND>::NT;
  Info &I = *ip;
  const int N = c.N, r = c.r;
  const double *X = c.V(V_X);
  const double *YD = c.cd(C_YD), *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
  const double *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV), *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double a_rp = 0, a_z = 0, a_Ax = 0, b_rp = 0, b_z = 0, b_Ax = 0;
  double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, b_rd = 0, b_q = 0, b_Aty = 0, b_Px = 0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    if (c.xl) {
      const double Ax = rowA_dyn<KIND>(c, ED, X, VS, k), z = zsel * BE[o], rr = Ax - z, ei = EINV[o];
      a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, Ax);
      b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * Ax);
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const double ax = c.si(k, t) * X[k * VS + r], z = c.zi(k, t), rr = ax - z, ei = EIINV[c.ci(k, t)];
        a_rp = absmax(a_rp, rr); a_z = absmax(a_z, z); a_Ax = absmax(a_Ax, ax);
        b_rp = absmax(b_rp, ei * rr); b_z = absmax(b_z, ei * z); b_Ax = absmax(b_Ax, ei * ax);
      }
    }
    if (c.var_live(k)) {
      const double Px = rowP<KIND>(c, PD, PO, X, VS, k), Aty = colA<KIND>(c, ED, YD, nullptr, true, k);
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
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED);
  const double *PVYI = c.cd(C_PVYI), *E = c.cd(C_E), *EI = c.cd(C_EI), *DINV = c.cd(C_DINV);
  double *DYD = c.cd(C_DYD), *DYI = c.cd(C_PYI), *XT = c.cd(C_ZT);
  const double ia = 1.0 / alpha, oma = 1.0 - alpha, cb = last_was_first ? 1.0 : alpha;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    double pvx, dsc;
    tm_ld2(c.tQ(k), pvx, dsc);   // the saved iterate lives in tensor memory (collective load: outside any branch)
    XT[o] = c.var_live(k) ? (X[k * VS + r] - oma * pvx) * ia : 0.0;
    (void)dsc;
  }
  __syncwarp();
  double nrm = 0.0, lhs = 0.0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    double d = 0.0;
    if (c.xl) d = rho_eq * (alpha * rowA_dyn<KIND>(c, ED, XT, 8, k) - cb * BE[o]);  // equality rows: no projection
    DYD[o] = d;
    if (c.xl) {
      nrm = absmax(nrm, unscale ? E[o] * d : d);
      lhs += BE[o] * ((d > 0) ? d : 0) + BE[o] * ((d < 0) ? d : 0);
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oc = c.ci(k, t);
        double di = c.yi(k, t) - PVYI[oc];
        const double lo = c.lo_of(k, t), up = c.ui(k, t);
        if (up > kInfty * kMinScaling) {
          if (lo < -kInfty * kMinScaling) di = 0.0;
          else di = (di < 0.0) ? di : 0.0;
        } else if (lo < -kInfty * kMinScaling) di = (di > 0.0) ? di : 0.0;
        DYI[oc] = di;
        nrm = absmax(nrm, unscale ? EI[oc] * di : di);
        lhs += up * ((di > 0) ? di : 0) + lo * ((di < 0) ? di : 0);
      }
    }
  }
  nrm = gmax(nrm);
  lhs = gsum(lhs);
  __syncwarp();
  // the product with A' is only needed when the first two conditions of the certificate hold for some group
  if (!__any_sync(kFull, (nrm > eps) && (lhs < -eps * nrm))) return false;
  double mx = 0.0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    if (c.var_live(k)) {
      const double at = colA<KIND>(c, ED, DYD, DYI, false, k);
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
  const double *X = c.V(V_X);
  const double *QV = c.cd(C_Q), *ED = c.cd(C_ED);
  const double *DINV = c.cd(C_DINV), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV);
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double *DX = c.cd(C_PX);
  double nrm = 0.0, qdx = 0.0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    double pvx, dsc;
    tm_ld2(c.tQ(k), pvx, dsc);   // saved iterate and scaling D from tensor memory: no L2 round trip in this loop
    const double dx = c.var_live(k) ? X[k * VS + r] - pvx : 0.0;
    DX[o] = dx;
    nrm = absmax(nrm, unscale ? dsc * dx : dx);
    qdx += QV[o] * dx;
  }
  nrm = gmax(nrm); qdx = gsum(qdx);
  __syncwarp();
  const double cs = unscale ? ip->csc : 1.0;
  // the products with P and A are only needed when the first two conditions of the certificate hold for some group
  if (!__any_sync(kFull, (nrm > eps) && (qdx < -cs * eps * nrm))) return false;
  double mx = 0.0;
  int viol = 0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    if (c.var_live(k)) {
      const double Pdx = rowP<KIND>(c, PD, PO, DX, 8, k);
      mx = absmax(mx, unscale ? DINV[o] * Pdx : Pdx);
    }
    if (c.xl) {
      double v = rowA_dyn<KIND>(c, ED, DX, 8, k);
      if (unscale) v = EINV[o] * v;
      if (v > eps * nrm || v < -eps * nrm) viol = 1;
    }
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        double v = c.si(k, t) * DX[o];
        if (unscale) v = EIINV[c.ci(k, t)] * v;
        if (((c.ui(k, t) < kInfty * kMinScaling) && (v > eps * nrm)) || ((c.lo_of(k, t) > -kInfty * kMinScaling) && (v < -eps * nrm))) viol = 1;
      }
    }
  }
  mx = gmax(mx);
  viol = gany(viol);
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
  for (int k = 0; k <= c.N; ++k)
    if (c.var_live(k)) acc += (0.5 * rowP<KIND>(c, PD, PO, xv, vs, k) + QV[k * 8 + c.r]) * xv[k * vs + c.r];
  return gsum(acc) * scale;
}

// ---------------------------------------------------------------- setup: schedule + build + Ruiz (cold, once per QP)
// Works in shared memory (scratch in the factor area and in the R / CR / XS vectors), then moves the cold data to the
// slab.  Returns flags: bit 0 = Curvature() failed, bit 1 = l > u somewhere.
template <int KIND>
__device__ __noinline__ int setup(const Ctx<KIND> c, const H8Params &p, const int b, const bool valid, double *csc_out,
                                 uint64_t *eqm_out, uint64_t *loosem_out) {
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT, GS = NX * 8;
  const int N = c.N, r = c.r;
  const Lay &L = *c.L;
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &St = p.S;
  double *S = c.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N, NS8 = (N + 1) * 8;
  (void)nx;
  const int ucomp = r - NX;
  int sched_err = 0, data_err = 0;
  double x0r = 0.0;
  // scratch: the Ruiz work vectors sit in the slots of the cold vectors that are only filled at the end, the column
  // norms of P in the factorisation scratch; G is built in place (swizzled rows)
  double *sPD = c.cd(C_PD), *sPO = c.cd(C_PO), *sD = c.cd(C_Q), *sE = c.cd(C_BE), *sEI = c.cd(C_ED), *sDt = c.cd(C_YD);
  double *sEt = c.cd(C_DINV), *sEti = c.cd(C_EINV);
  double *Gs = S + L.G;
  (void)NS8;
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
        double *gk = Gs + k * GS;
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
        if (notfinite(v)) data_err = 1;   // e.g. vx = 0 or 1 - ey kappa = 0 in the stage matrices
      }
      if (c.xl) {
        double *gk = Gs + k * GS;
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
    for (int k = 0; k <= N; ++k) {
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
      for (int k = 0; k <= N + 1; ++k) {
        double *ibk = c.Ib(k);
        if (r < NSLc) {
          ibk[r * 2] = 0.0; ibk[r * 2 + 1] = 0.0; ibk[NSLc * 2 + r * 2] = 0.0; ibk[NSLc * 2 + r * 2 + 1] = kInfty * kInfty;
          if (KIND == LPVMPC_PLANNER) ibk[OLIc + r] = -kInfty * kInfty;
        }
        if (r < 2) ibk[OPMc + r] = 0.0;
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
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
    __syncwarp();
  }
  data_err = gany(data_err);

  // ---- Ruiz equilibration (OSQP scale_data)
  // The cost scale of a pass (OSQP: c = 1 / max(mean column norm of P, |q|_inf)) is not applied by a loop of its own: it
  // stays pending in `cp` and is folded into the next pass's reads (x * 1.0 is exact, and (x * cp) is the value the
  // separate loop used to store, so every product is formed in the same order); the cost norms are taken from the
  // registers of the scaling loop.  Both loops were chains of exposed shared-memory round trips (3 % of the kernel).
  double csc = 1.0, cp = 1.0;
  // Likewise the column norms of a pass are taken while the previous pass scales the columns (every entry of column
  // (k, r) is written by lane r in the scaling loop): max |P column| (before the pending cost scale) and max |A column|
  // wait in two free stage vectors (DG: set by factor; B: set by reproject, re-zeroed below), and the norms of the
  // single-variable rows go straight to sEti.  Only the first pass reads the columns.
  double *sPC = c.V(V_DG), *sQA = c.V(V_B);
#pragma unroll 1
  for (int it = 0; it < St.scaling; ++it) {
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r, ov = k * VS + r;
      double pa, qa;
      if (it == 0) {
        pa = fabs(sPD[o]);
        if (c.ul) {
          if (k < N - 1) pa = absmax(pa, sPO[o]);
          if (k > 0 && k < N) pa = absmax(pa, sPO[o - 8]);
        }
        qa = c.xl ? fabs(ED[ov]) : 0.0;
        if (k < N) {
          const double *gk = Gs + k * GS;
#pragma unroll
          for (int rr = 0; rr < NX; ++rr) qa = absmax(qa, gk[rr * 8 + c.co[rr >> 1]]);
        }
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            qa = absmax(qa, c.si(k, t));
            sEti[c.ci(k, t)] = frsqrt(limit_scaling(fabs(c.si(k, t))));
          }
        }
      } else {
        pa = sPC[ov] * cp;   // max |x| * cp == max |x * cp| (cp > 0, rounding is monotone)
        qa = sQA[ov];
      }
      sDt[o] = frsqrt(limit_scaling(pa > qa ? pa : qa));
      double ea = c.xl ? fabs(ED[ov]) : 0.0;
      if (k > 0 && c.xl) {
        const double *gp = Gs + (k - 1) * GS;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double2 e = ld2(gp + c.ro[j]); ea = absmax(ea, e.x); ea = absmax(ea, e.y); }
      }
      sEt[o] = frsqrt(limit_scaling(ea));
    }
    __syncwarp();
    double qn = 0.0, ct = 0.0, npo_prev = 0.0;
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r, ov = k * VS + r;
      const double dt = sDt[o];
      double npo = 0.0, qan = 0.0;   // qan: max |entry| of my column of A after this pass
      if (k < N) {
        double *gk = Gs + k * GS;
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) {
          double *e = gk + rr * 8 + c.co[rr >> 1];
          const double g = (*e * sEt[(k + 1) * 8 + rr]) * dt;
          *e = g;
          qan = absmax(qan, g);
        }
        if (k < N - 1 && c.ul) { npo = ((sPO[o] * cp) * dt) * sDt[o + 8]; sPO[o] = npo; }
      }
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oc = c.ci(k, t);
          const double et = sEti[oc], sn = (c.si(k, t) * et) * dt;
          c.si(k, t) = sn;
          sEI[oc] = sEI[oc] * et;
          qan = absmax(qan, sn);
          sEti[oc] = frsqrt(limit_scaling(fabs(sn)));   // next pass's norm of this one-entry row
        }
      }
      if (c.xl) { const double en = (ED[ov] * sEt[o]) * dt; ED[ov] = en; qan = absmax(qan, en); }
      sQA[ov] = qan;
      const double npd = ((sPD[o] * cp) * dt) * dt;
      sPD[o] = npd;
      const double nq = dt * (QV[ov] * cp);
      QV[ov] = nq;
      sD[o] = sD[o] * dt;
      sE[o] = sE[o] * sEt[o];
      // cost scaling: mean of the column norms of P (summed per lane, then across the group)
      double pc = fabs(npd);
      if (c.ul) {
        if (k < N - 1) pc = absmax(pc, npo);
        if (k > 0 && k < N) pc = absmax(pc, npo_prev);
      }
      if (c.var_live(k)) { ct += pc; qn = absmax(qn, nq); }
      sPC[ov] = pc;
      npo_prev = npo;
    }
    __syncwarp();
    qn = gmax(qn);
    ct = gsum(ct) / nz;
    qn = limit_scaling(qn);
    ct = ct > qn ? ct : qn;
    ct = limit_scaling(ct);
    ct = frcp(ct);
    cp = ct;
    csc *= ct;
  }
  if (St.scaling > 0) {
#pragma unroll 1
    for (int k = 0; k <= N; ++k) { const int o = k * 8 + r; sPD[o] *= cp; QV[k * VS + r] *= cp; sPO[o] *= cp; }
    __syncwarp();
  }
  *csc_out = csc;
  // ---- scaled bounds, constraint classes; then the work vectors give way to the cold vectors they were parked in
  {
    uint64_t eqm = 0, loosem = 0;
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
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
    *eqm_out = eqm; *loosem_out = loosem;
    __syncwarp();
    double *cD = c.cd(C_D), *cE = c.cd(C_E), *cEi = c.cd(C_EI), *cEiI = c.cd(C_EIINV);
    double *cQ = c.cd(C_Q), *cBE = c.cd(C_BE), *cED = c.cd(C_ED), *cYD = c.cd(C_YD), *cDI = c.cd(C_DINV), *cEI = c.cd(C_EINV);
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r, ov = k * VS + r;
      // element o of every work vector is mine alone: read them all, then overwrite their slots
      const double d = sD[o], e = sE[o], ei = sEI[o], q = QV[ov], be = BE[ov], ed = ED[ov];
      cD[o] = d; cE[o] = e; cEi[o] = ei; cEiI[o] = 1.0 / ei;
      tm_st1(c.tQ(k) + 2u, d);   // copy of D for the dual-infeasibility certificate
      cQ[o] = c.var_live(k) ? q : 0.0; cBE[o] = e * be; cED[o] = ed; cYD[o] = 0.0; cDI[o] = 1.0 / d; cEI[o] = 1.0 / e;
      if (c.ul) c.pm(k, ucomp) = (k > 0 && k < N) ? sPO[o - 8] : 0.0;   // couples u_{k-1}, u_k
      c.V(V_XS)[ov] = 0.0;   // the scratch homes become hot vectors (R, CR, B are set by reproject, DG by factor)
      c.V(V_B)[ov] = 0.0;
    }
    __syncwarp();
  }
  return (sched_err ? 1 : 0) | (data_err ? 2 : 0);
}

// ---------------------------------------------------------------- polish (cold, once per QP)
// OSQP's polish (upstream polish.c; oracle/osqp_ref.c:polish): the reduced KKT system of the active rows,
//   [P A_red'; A_red 0] (x, y) = (-q, b_red),
// by iterative refinement with the regularised matrix K_reg = [P + delta I, A_red'; A_red, -delta I]:
//   s_0 = K_reg^-1 rhs;   s_{j+1} = s_j + K_reg^-1 (rhs - K s_j),   j < polish_refine_iter        (exactly as upstream)
// A solve with K_reg is done in condensed form, (P + delta I + A_red' A_red / delta) dx = r1 + A_red' r2 / delta,
// dy = (A_red dx - r2) / delta, with the block factor of the ADMM steps (row weights 1 / delta).  That solve loses
// about 6 digits to the 1 / delta weights, where upstream's quasi-definite LDL' is accurate to round-off, so every
// K_reg solve is followed by kPolishInner refinement passes ON THE REGULARISED SYSTEM (residual
// r1 - (P + delta I) dx - A_red' dy, dy slaved to dx): the outer iteration then is upstream's, pass for pass, and so are
// the polished iterates of the slowly converging (ill-conditioned) problems.  (Round 1 ran polish_refine_iter + 2 outer
// passes instead: 2 of 16,384 planner QPs ended 6e-3 away from the oracle)
// State: t = s + w (tentative iterate: x part TX, dynamics duals TY, duals of my single-variable rows in tensor memory),
// the x part DX of the increment w of the running K_reg solve; residuals are formed from t and DX alone:
//   r1 - (P + delta I) dx - A' dy = -q - P tx - A' ty - delta dx.
template <int KIND>
__device__ __noinline__ int polish(const Ctx<KIND> c, const Hot<KIND> h, const lpvmpc_settings &St, Info *ip, const bool do_pol, uint32_t gsel) {
  constexpr int NT = Ctx<KIND>::NT, NX = Ctx<KIND>::NX;
  Info &I = *ip;
  const bool unscale = I.unscale;
  const int N = c.N, r = c.r;
  double *X = c.V(V_X), *BV = c.V(V_B);
  double *PX = c.V(V_R), *PYD = c.V(V_XS), *DX = c.V(V_DG), *TMP = c.V(V_CR);       // [k*VS + q]: tx, ty (dynamics rows), dx, row temporary
  double *YD = c.cd(C_YD);
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
  double *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);   // slab copies for factor() and the outputs; the loops below use the register copies
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV);
  const double delta = St.delta, idel = 1.0 / St.delta;
  // The duals of the single-variable rows (and at the end their polished z) live in the tensor-memory columns behind the
  // factor (c.tP(k): 4 doubles per lane and stage), the active-set flags in registers.  The tcgen05 loads / stores are
  // warp-collective: they sit at the top / bottom of the stage loops, outside any branch.
  uint64_t acti = 0;   // 2 bits per row (k, t) of mine: 1 = lower, 2 = upper, 3 = both
  uint32_t actd = 0;   // bit k: my dynamics row (k, r) is in the polish system
  auto ai_of = [&](int k, int t) { return (int)((acti >> (2 * (k * NT + t))) & 3ull); };
  // active-set guess (form_Ared)
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    double ad = 0.0;
    // equality row (z == l == u): lower / upper by the sign of the dual.  Upstream drops the row when the dual is exactly
    // 0.0 (its polish system is then singular and the polish is rejected: a floating-point accident that happens to 4 of the
    // 4,096 QPs of configs[1] in the oracle, on the initial-condition row of the unweighted state `s`).  y_dyn recovered from
    // the running sum is exactly 0.0 on such rows far more often, and the polish system needs every dynamics row: keep it
    // (as "lower").
    if (c.xl && do_pol) ad = (0.0 < YD[o]) ? 2.0 : 1.0;
    ACTD[o] = ad;
    if (ad != 0.0) actd |= 1u << k;
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        int ai = 0;
        if (do_pol) { if (c.zi(k, t) - c.lo_of(k, t) < -c.yi(k, t)) ai += 1; if (c.ui(k, t) - c.zi(k, t) < c.yi(k, t)) ai += 2; }
        ACTI[c.ci(k, t)] = (double)ai;
        acti |= (uint64_t)ai << (2 * (k * NT + t));
      }
    }
  }
  __syncwarp();
  FW fw; fw.polish = 1; fw.rho = 0.0; fw.rho_eq = 0.0; fw.idel = idel;
  factor<KIND>(c, fw, delta);
  auto bred_i = [&](int k, int t) { const int a = ai_of(k, t); return (a == 1 || a == 3) ? c.lo_of(k, t) : c.ui(k, t); };
  // (A' t)_(k, r) with t_dyn stored [k*VS + q] and t_in = (ti0, ti1) of my single-variable rows
  auto colAt = [&](const double *td, const double ti0, const double ti1, int k) {
    double acc = c.xl ? ED[k * 8 + r] * td[k * VS + r] : 0.0;
    if (k < N) {
      double g[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? td[(k + 1) * VS + rr] : 0.0;
      acc += coldot<NX>(c.Gb(k), c.co, g);
    }
    if (c.has_in(k)) {
      acc = fma(c.si(k, 0), ti0, acc);
      if (NT > 1) acc = fma(c.si(k, NT - 1), ti1, acc);
    }
    return acc;
  };
  // one K_reg solve in condensed form: BV holds the right-hand side on entry, ddx on exit
  auto solve = [&]() {
    __syncwarp();
    sweep_fwd<KIND>(h, N, gsel);
    sweep_bwd_plain<KIND>(h, N, gsel);
  };
  // refinement passes of the running K_reg solve: rhs = -q - P tx - A' ty - delta dx; tx, dx += ddx; ty += A ddx / delta
  auto inner = [&]() {
#pragma unroll 1
    for (int ii = 0; ii < kPolishInner; ++ii) {
#pragma unroll 1
      for (int k = 0; k <= N; ++k) {
        const int o = k * 8 + r, ov = k * VS + r;
        TmQuad pq;
        tm_ld4(c.tP(k), pq);
        tm_wait_ld();
        double b = 0.0;
        if (c.var_live(k)) {
          const double Px = rowP<KIND>(c, PD, PO, PX, VS, k);
          const double Aty = colAt(PYD, tm_getq(pq, 0), tm_getq(pq, 1), k);
          b = ((-QV[o] - Px) - Aty) - delta * DX[ov];
        }
        BV[ov] = b;
      }
      solve();
#pragma unroll 1
      for (int k = 0; k <= N; ++k) {
        const int ov = k * VS + r;
        TmQuad pq;
        tm_ld4(c.tP(k), pq);
        tm_wait_ld();
        const double ddx = BV[ov];
        if (c.xl && ((actd >> k) & 1u)) PYD[ov] += idel * rowA_dyn<KIND>(c, ED, BV, VS, k);
        double py[2] = {tm_getq(pq, 0), tm_getq(pq, 1)};
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) if (ai_of(k, t) != 0) py[t] += idel * (c.si(k, t) * ddx);
        }
        tm_st4(c.tP(k), py[0], py[1], 0.0, 0.0);
        PX[ov] += ddx; DX[ov] += ddx;
      }
      tm_wait_st();
      __syncwarp();
    }
  };
  // ---- s_0 = K_reg^-1 (-q, b_red): rhs = -q + A_red'(b_red / delta)
#pragma unroll 1
  for (int k = 0; k <= N; ++k) TMP[k * VS + r] = (c.xl && ((actd >> k) & 1u)) ? idel * BE[k * 8 + r] : 0.0;
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    double t2[2] = {0.0, 0.0};
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) t2[t] = (ai_of(k, t) != 0) ? idel * bred_i(k, t) : 0.0;
    }
    BV[k * VS + r] = c.var_live(k) ? (-QV[k * 8 + r] + colAt(TMP, t2[0], t2[1], k)) : 0.0;
  }
  solve();
  // tx = dx = ddx;  ty = (A tx - b_red) / delta on the active rows
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r, ov = k * VS + r;
    const double xk = BV[ov];
    PX[ov] = xk; DX[ov] = xk;
    PYD[ov] = (c.xl && ((actd >> k) & 1u)) ? idel * (rowA_dyn<KIND>(c, ED, BV, VS, k) - BE[o]) : 0.0;
    double py[2] = {0.0, 0.0};
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) py[t] = (ai_of(k, t) != 0) ? idel * (c.si(k, t) * xk - bred_i(k, t)) : 0.0;
    }
    tm_st4(c.tP(k), py[0], py[1], 0.0, 0.0);
  }
  tm_wait_st();
  __syncwarp();
  inner();
#pragma unroll 1
  for (int it = 0; it < St.polish_refine_iter; ++it) {
    // s_{j+1} = s_j + K_reg^-1 (rhs - K s_j): r1 = -q - P tx - A' ty, r2 = b_red - A tx;
    // condensed right-hand side r1 + A' r2 / delta = -q - P tx - A'(ty - r2 / delta)
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r, ov = k * VS + r;
      TMP[ov] = (c.xl && ((actd >> k) & 1u)) ? fma(-idel, BE[o] - rowA_dyn<KIND>(c, ED, PX, VS, k), PYD[ov]) : 0.0;
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r, ov = k * VS + r;
      TmQuad pq;
      tm_ld4(c.tP(k), pq);
      tm_wait_ld();
      double b = 0.0;
      if (c.var_live(k)) {
        const double Px = rowP<KIND>(c, PD, PO, PX, VS, k);
        double t2[2] = {0.0, 0.0};
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t)
            t2[t] = (ai_of(k, t) != 0) ? fma(-idel, bred_i(k, t) - c.si(k, t) * PX[ov], tm_getq(pq, t)) : 0.0;
        }
        b = (-QV[o] - Px) - colAt(TMP, t2[0], t2[1], k);
      }
      BV[ov] = b;
    }
    solve();
    // dx = ddx; tx += ddx (every stage, before the rows of A are applied to the new tx)
#pragma unroll 1
    for (int k = 0; k <= N; ++k) { const int ov = k * VS + r; const double ddx = BV[ov]; DX[ov] = ddx; PX[ov] += ddx; }
    __syncwarp();
    // ty += (A ddx - r2) / delta = (A tx_new - b_red) / delta
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      const int o = k * 8 + r, ov = k * VS + r;
      TmQuad pq;
      tm_ld4(c.tP(k), pq);
      tm_wait_ld();
      if (c.xl && ((actd >> k) & 1u)) PYD[ov] += idel * (rowA_dyn<KIND>(c, ED, PX, VS, k) - BE[o]);
      double py[2] = {tm_getq(pq, 0), tm_getq(pq, 1)};
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) if (ai_of(k, t) != 0) py[t] += idel * (c.si(k, t) * PX[ov] - bred_i(k, t));
      }
      tm_st4(c.tP(k), py[0], py[1], 0.0, 0.0);
    }
    tm_wait_st();
    __syncwarp();
    inner();
  }
  // pol z = A x, normal-cone projection, residuals, acceptance.  Polished z of the single-variable rows -> columns 2, 3.
  double a_rp = 0, a_rd = 0;
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r, ov = k * VS + r;
    TmQuad pq;
    tm_ld4(c.tP(k), pq);
    tm_wait_ld();
    if (c.xl) {
      const double Ax = rowA_dyn<KIND>(c, ED, PX, VS, k), t = Ax + PYD[ov];
      TMP[ov] = t - BE[o];       // projected dual (PYD keeps feeding rowA-free reads of this loop's neighbours: none, but keep the pass pure)
      const double rr = Ax - BE[o];
      a_rp = absmax(a_rp, unscale ? EINV[o] * rr : rr);
    } else TMP[ov] = 0.0;
    double py[2] = {tm_getq(pq, 0), tm_getq(pq, 1)}, zz[2] = {0.0, 0.0};
    if (c.has_in(k)) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int oc = c.ci(k, t);
        const double ax = c.si(k, t) * PX[ov], tt = ax + py[t];
        const double zc = clampd(tt, c.lo_of(k, t), c.ui(k, t));
        zz[t] = zc; py[t] = tt - zc;
        const double rr = ax - zc;
        a_rp = absmax(a_rp, unscale ? EIINV[oc] * rr : rr);
      }
    }
    tm_st4(c.tP(k), py[0], py[1], zz[0], zz[1]);
  }
  tm_wait_st();
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {
    const int o = k * 8 + r;
    TmQuad pq;
    tm_ld4(c.tP(k), pq);
    tm_wait_ld();
    if (c.var_live(k)) {
      const double rr = (QV[o] + rowP<KIND>(c, PD, PO, PX, VS, k)) + colAt(TMP, tm_getq(pq, 0), tm_getq(pq, 1), k);
      a_rd = absmax(a_rd, unscale ? DINV[o] * rr : rr);
    }
  }
  const double pol_pri = gmax(a_rp), pol_dua = (unscale ? I.cinv : 1.0) * gmax(a_rd);
  const double pol_obj = objective<KIND>(c, PX, VS, St.scaling ? I.cinv : 1.0);
  const bool ok = (pol_pri < I.pri_res && pol_dua < I.dua_res) || (pol_pri < I.pri_res && I.dua_res < 1e-10) ||
                  (pol_dua < I.dua_res && I.pri_res < 1e-10);
  const bool take = do_pol && ok;
  if (take) { I.obj = pol_obj; I.pri_res = pol_pri; I.dua_res = pol_dua; }
#pragma unroll 1
  for (int k = 0; k <= N; ++k) {   // the tensor-memory loads are collective: every group walks the loop, only `take` groups store
    const int o = k * 8 + r, ov = k * VS + r;
    TmQuad pq;
    tm_ld4(c.tP(k), pq);
    tm_wait_ld();
    if (take) {
      X[ov] = PX[ov]; YD[o] = TMP[ov];
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) { c.zi(k, t) = tm_getq(pq, 2 + t); c.yi(k, t) = tm_getq(pq, t); }
      }
    }
  }
  if (!do_pol) return 0;
  return ok ? 1 : -1;
}

// ---------------------------------------------------------------- persistent warps, QPW QPs at a time each
constexpr int kSyncEvery = 25;  // y_dyn is brought up to date and r re-projected at least every kSyncEvery steps

template <int KIND, int QPW>
__global__ void __launch_bounds__(128, 1) lpv_solve_h8t_kernel(const __grid_constant__ H8Params p) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NX = Ctx<KIND>::NX, NB = Ctx<KIND>::NB, NT = Ctx<KIND>::NT, NSL = Ctx<KIND>::NSL;
  constexpr int OLI = Ctx<KIND>::OLI, OPM = Ctx<KIND>::OPM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
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
  double *wsm = smem + (size_t)warp * QPW * L.total;
  const size_t wslot = (size_t)(blockIdx.x * wpc + warp) * QPW;
  // all-gather buffers: two 256-byte buffers per warp, 512-byte aligned, after the QP regions
  const uint32_t smem_a = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t gbuf = ((smem_a + (uint32_t)(wpc * QPW * L.total * 8) + 511u) & ~511u) + (uint32_t)warp * 512u;
  uint32_t gsel = 0;
  // tensor memory: the whole SM's 512 columns (one CTA per SM); warp w owns lanes 32 (w % 4) .. +31
  __shared__ uint32_t tmem_base_s;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  c.tm = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16);

  for (;;) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(p.queue, (unsigned)QPW);
    base = __shfl_sync(kFull, base, 0);
    if ((int)base >= p.B) break;
    // Groups without a problem of their own (batch tail, or g >= QPW) mirror group 0 exactly: same problem, same
    // shared / scratch region, same values written by the same instruction; only user-visible outputs are guarded.
    const bool valid = (g < QPW) && ((int)(base + g) < p.B);
    const int gq = valid ? g : 0;
    const int b = p.perm ? p.perm[(int)base + gq] : (int)base + gq;
    c.S = wsm + gq * L.total;
    c.cold = p.cold + (wslot + gq) * L.cold_total;

    Hot<KIND> h;
    {
      const uint32_t sq = smem_a + (uint32_t)((warp * QPW + gq) * L.total) * 8u;
      h.tT = c.tT(0); h.tKr = c.tKr(1); h.tKc = c.tKc(1);
      h.v = sq + (uint32_t)(L.V + r) * 8u;
      constexpr int ISBk = Ctx<KIND>::IS * 8;
      const bool rows = (KIND == LPVMPC_CONTROLLER) ? (r == 0 || c.ul) : (c.xl || c.ul);
      const uint32_t dm = sq + (uint32_t)(L.I + (N + 1) * Ctx<KIND>::IS) * 8u;   // block N+1: dummy rows, zero coupling
      h.ib = rows ? sq + (uint32_t)(L.I + c.islot * 2) * 8u : dm;
      h.il = rows ? sq + (uint32_t)(L.I + OLI + c.islot) * 8u : dm + (uint32_t)OLI * 8u;
      h.istr = rows ? (uint32_t)ISBk : 0u;
      h.pm = c.ul ? sq + (uint32_t)(L.I + OPM + ucomp) * 8u : dm + (uint32_t)OPM * 8u;
      h.pm2 = c.ul ? h.pm + (uint32_t)ISBk : h.pm;
      h.pstr = c.ul ? (uint32_t)ISBk : 0u;
      h.gpub = gbuf + (uint32_t)(64 * (r >> 1) + 16 * g + 8 * (r & 1));
      h.ggat = gbuf + (uint32_t)(16 * g);
    }

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
    double rho_eq = kRhoEqOverIneq * rho;
    {
      FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0;
      factor<KIND>(c, fw, sigma);
    }
    reproject<KIND>(c, true, rho, rho_eq, sigma, 0.0, true);
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
#pragma unroll 1
      for (; iter < stop; ++iter) {
        if (iter == stop - 1) {  // keep the iterate before the last step of the chunk: delta_x, delta_y
          double *PVYI = c.cd(C_PVYI);
          const double *X = c.V(V_X);
#pragma unroll 1
          for (int k = 0; k <= N; ++k) {
            tm_st1(c.tQ(k), X[k * VS + r]);   // collective: every group stores (a frozen group's copy is never read again)
            if (live && c.has_in(k)) {
#pragma unroll
              for (int t = 0; t < NT; ++t) PVYI[c.ci(k, t)] = c.yi(k, t);
            }
          }
          tm_wait_st();
        }
        u.cc = (iter == 0) ? 2.0 : alpha;
        sweep_fwd<KIND>(h, N, gsel);
        sweep_bwd_admm<KIND>(h, u, gsel);
        if (iter == 0) first_in = 1;
        ++nsync;
        zsel = 1.0;
      }
      __syncwarp();
      const int last_was_first = (iter == 1);
      rho_eq_last = rho_eq;
      sync_yd<KIND>(c, live, rho_eq, alpha, nsync, first_in);
      nsync = 0; first_in = 0;
      const bool can_check = ct && (iter % ct == 0);
      const bool can_adapt = ai && (iter % ai == 0);
      bool new_cr = false;
      checked_last = can_check;
      if (can_check || can_adapt) {
        Info J = I;
        update_info<KIND>(c, &J, zsel);
        if (live) { I = J; iter_done = iter; }
        if (can_check) {
          if (check_termination<KIND>(c, S, &I, live, 0, rho_eq_last, last_was_first, true)) live = false;  // frozen: stores are predicated on `live`
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
            __syncwarp();
            factor<KIND>(c, fw, sigma);
            new_cr = true;
          }
        }
      }
      // re-project r (and the pending right-hand side) from the explicit iterate: removes the drift of the recursion
      if (iter < S.max_iter && __any_sync(kFull, live)) reproject<KIND>(c, live, rho, rho_eq, sigma, zsel, new_cr);
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
    if (valid && (a.xs || a.zs || a.ys)) {
      const double *X = c.V(V_X);
      const double *YD = c.cd(C_YD), *BE = c.cd(C_BE);
#pragma unroll 1
      for (int k = 0; k <= N; ++k) {
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
    int polish_status = 0;
    const bool do_pol = S.polish && status == LPVMPC_SOLVED;
    bool polished_sets = false;
    if (__any_sync(kFull, do_pol)) {
      polish_status = polish<KIND>(c, h, S, &I, do_pol, gsel);
      polished_sets = do_pol;
    }
    // ---- outputs
    if (valid) {
      const double *X = c.V(V_X);
      const double *YD = c.cd(C_YD), *D = c.cd(C_D), *E = c.cd(C_E), *EI = c.cd(C_EI), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
#pragma unroll 1
      for (int k = 0; k <= N; ++k) {
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
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
  (void)NSL; (void)NB;
}

}  // namespace h8t
}  // namespace lpv
