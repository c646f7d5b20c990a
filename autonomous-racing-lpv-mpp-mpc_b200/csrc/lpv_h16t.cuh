// "H16T" kernel: 16 lanes per QP, 2 QPs per warp, 8 warps per CTA, one CTA per SM; TWISTED block factorisation.
//
// Same ADMM formulation and the same data homes as lpv_h8t.cuh (block factor in tensor memory, stage vectors / single-
// variable rows / G / the hot cold vectors in shared memory); what changes is how a QP is spread over lanes.  H8T gives a
// QP 8 lanes and walks the block-tridiagonal system end to end: 2 (N + 1) dependent stage steps per ADMM step, 16 QPs =
// 4 warps per SM, one warp per scheduler, every dependent-issue stall exposed (profiles/r2u_ncu_h8t_*: warps active 6 %,
// issue active 23 %).  Here the two 8-lane halves of a QP eliminate the system FROM BOTH ENDS towards the middle stage
// m = N / 2 ("twisted" / two-sided factorisation):
//
//   left  half: stages 0 .. m-1 as before:           T_k = S~_k^-1,  K_k = S_{k,k-1} T_{k-1},  S~_k = S_k - K_k S_{k,k-1}'
//   right half: stages N .. m+1, mirrored:            T_k = S^_k^-1,  J_k = S_{k,k+1} T_{k+1},  S^_k = S_k - J_k S_{k,k+1}'
//   middle:     S~_m = S_m - K_m S_{m,m-1}' - J_m S_{m,m+1}',  T_m = S~_m^-1   (the two halves add their Schur terms)
//
//   forward:  v_k = b_k - K_k v_{k-1} (left),  v_k = b_k - J_k v_{k+1} (right),  v_m = b_m - K_m v_{m-1} - J_m v_{m+1}
//   backward: x_m = T_m v_m;  x_k = T_k v_k - K_{k+1}' x_{k+1} (left),  x_k = T_k v_k - J_{k-1}' x_{k-1} (right)
//
// In "local steps" j (stage j for the left half, stage N - j for the right half) both halves run the SAME instruction
// stream on mirrored data, so one ADMM step is N + 2 dependent stage steps instead of 2 (N + 1), the cold per-stage loops
// (Ruiz, re-projection, residuals, certificates, polish) take half their trips, and the SM holds the same 16 QPs in 8
// warps: two warps per scheduler.  The middle stage is owned by both halves: they compute identical values from identical
// operands (a + b == b + a) and store them to the same addresses.
//
// Tensor memory: warp w reaches lanes 32 (w % 4) .. +31; warps w and w + 4 share a lane quarter and take 256 columns each.
// tcgen05.ld / st address ONE column range per instruction for the whole warp, so both halves use the same layout
// (N = 8, NL = 4):  T rows of local steps 0..4 (80 columns) | K / J rows of local steps 1..4 (64) | K / J columns (64) |
// 8 columns per cold stage slot for the polish duals, aliased by the certificates' saved iterate and scaling copy (40):
// 248 of 256.
//
// Cold per-stage loops: the left half owns stages 0 .. NL-1, the right half stages NL .. N (one more); loops that touch
// tensor memory run NL + 1 trips in both halves (the left half's last trip is a guarded dummy).
// The factorisation scratch (previous pivot inverse, parked off-diagonal block: 2 x 64 doubles per half) lives in stage
// vector slots that are dead while a factorisation runs (B, R for the left half, CR, XS for the right half; XS is
// re-zeroed), which frees the 1 KB per QP that H8T spends on it.
//
// Controller problems only (the launch configuration of the reference: N = 8, diagonal Q / R, no steering delay).
#pragma once
#include "lpv_h8t.cuh"

namespace lpv {
namespace h16t {

using h8t::VS;
using h8t::VB;
using h8t::V_B; using h8t::V_X; using h8t::V_R; using h8t::V_XS; using h8t::V_DG; using h8t::V_CR;
using h8t::C_PD; using h8t::C_PO; using h8t::C_Q; using h8t::C_BE; using h8t::C_ED; using h8t::C_YD; using h8t::C_DINV; using h8t::C_EINV;
using h8t::C_NSMEM; using h8t::C_D; using h8t::C_E; using h8t::C_EI; using h8t::C_EIINV; using h8t::C_PVYI; using h8t::C_DYD; using h8t::C_PX;
using h8t::C_PYI; using h8t::C_ACTD; using h8t::C_ACTI; using h8t::C_ZT; using h8t::C_COUNT;
using h8t::Lay; using h8t::H8Params; using h8t::Info; using h8t::FW; using h8t::TmRow; using h8t::TmQuad;
using h8t::ld2; using h8t::st2; using h8t::frcp; using h8t::frsqrt;
using h8t::tm_st8; using h8t::tm_ld8; using h8t::tm_st4; using h8t::tm_st1; using h8t::tm_ld2; using h8t::tm_ld4; using h8t::tm_get; using h8t::tm_getq;
using h8t::tm_wait_ld; using h8t::tm_wait_st;
using h8t::lds; using h8t::lds2; using h8t::sts; using h8t::sts2;
using h8t::dot8; using h8t::bwd_step; using h8t::gather_in;
using h8t::Upd; using h8t::UpdIn; using h8t::UpdMid; using h8t::update_loads; using h8t::update_part1; using h8t::update_part2;
using h8t::kSyncEvery;

constexpr int KIND = LPVMPC_CONTROLLER;
constexpr int NX = 6, NB = 8, NT = 2, NSL = 6, OPM = 24, IS = 26, GS = NX * 8;
// Offsets (doubles) of the blocks of a QP's shared-memory region: the layout lpvmpc.cu's make_h16t_layout computes for the
// one horizon this kernel is built for, as compile-time constants (lpvmpc_create checks them against the host's layout with
// layout_matches).  Through the `Lay` in the kernel parameters every accessor of a cold routine paid a generic load from
// parameter memory in front of its address (121 LD.E in the kernel's SASS, reloaded behind every asm volatile).
constexpr int kN8 = 8;
constexpr int kLV = 0, kLI = kLV + (kN8 + 1) * h8t::VS, kLG = kLI + (kN8 + 2) * IS, kLCS = kLG + kN8 * NX * 8;
static_assert(((kN8 + 1) * h8t::VS) % 2 == 0 && ((kN8 + 2) * IS) % 2 == 0 && (kN8 * NX * 8) % 2 == 0, "blocks are padded to even sizes on the host");
inline bool layout_matches(const Lay &L) { return L.N == kN8 && L.V == kLV && L.I == kLI && L.G == kLG && L.CS == kLCS; }

// reductions over the 8 lanes of a half (pivot rows) and over the 16 lanes of a QP
__device__ __forceinline__ double h8shfl(double v, int src) { return __shfl_sync(kFull, v, src, 8); }
__device__ __forceinline__ double xhalf(double v) { return __shfl_xor_sync(kFull, v, 8, 16); }   // the same lane of the other half
__device__ __forceinline__ double qmax(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) { const double w = __shfl_xor_sync(kFull, v, o, 16); v = (w > v) ? w : v; }
  return v;
}
__device__ __forceinline__ double qsum(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o, 16);
  return v;
}
__device__ __forceinline__ int qany(int v) {
  const unsigned m = __ballot_sync(kFull, v);
  return ((m >> ((threadIdx.x & 31) & ~15)) & 0xffffu) != 0;
}

struct Ctx {
  double *S;     // my QP's shared region
  double *cold;  // my QP slot in the global slab
  static constexpr int N = 8, NL = 4;   // horizon, local steps before the middle stage (N = 2 NL): this kernel is N = 8 only (lpvmpc_create)
  int r, h;      // component, half (0: left, stages ascending; 1: right, stages descending)
  int kc0, nck;  // cold ownership: stages kc0 .. kc0 + nck - 1 (left: 0 .. NL-1, right: NL .. N)
  int co[4];     // my column of a swizzled G block: offset inside row rr is co[rr >> 1]
  int ro[4];     // my row of a swizzled G block: offset of logical chunk j
  bool xl, ul;   // state lane / input lane
  int islot;     // first single-variable-row slot of my variable inside a stage (t adds 1)
  uint32_t tm;   // tensor memory: my lane quarter, first column of my warp's 256-column window

  __device__ __forceinline__ bool var_live(int k) const { return xl || (ul && k < N); }
  __device__ __forceinline__ bool has_in(int k) const { return (r == 0 || ul) && k < N; }
  __device__ __forceinline__ bool owns(int k) const { return k >= kc0 && k < kc0 + nck; }
  __device__ __forceinline__ int kstage(int j) const { return h ? N - j : j; }   // stage of local step j
  __device__ __forceinline__ double *cd(int arr) const {
    return (arr < C_NSMEM) ? (S + kLCS + arr * (N + 1) * 8) : (cold + (arr - C_NSMEM) * (N + 1) * 8);
  }
  __device__ __forceinline__ double *Gb(int k) const { return S + kLG + k * GS; }
  __device__ __forceinline__ double *V(int arr) const { return S + kLV + arr; }
  // factorisation scratch rows (8 doubles each) in dead stage-vector slots: previous pivot inverse / parked block
  __device__ __forceinline__ double *Tp(int row) const { return S + kLV + row * VS + (h ? V_CR : V_B); }
  __device__ __forceinline__ double *Ob(int row) const { return S + kLV + row * VS + (h ? V_XS : V_R); }
  // tensor-memory columns (the same for every lane of the warp): local step j, cold slot i
  __device__ __forceinline__ uint32_t tT(int j) const { return tm + (uint32_t)(16 * j); }
  __device__ __forceinline__ uint32_t tKr(int j) const { return tm + (uint32_t)(16 * (NL + 1) + 16 * (j - 1)); }
  __device__ __forceinline__ uint32_t tKc(int j) const { return tm + (uint32_t)(16 * (2 * NL + 1) + 16 * (j - 1)); }
  __device__ __forceinline__ uint32_t tP(int i) const { return tm + (uint32_t)(16 * (3 * NL + 1) + 8 * i); }
  __device__ __forceinline__ uint32_t tQ(int i) const { return tP(i); }   // certificates (ADMM) / polish duals (afterwards)
  __device__ __forceinline__ double *Ib(int k) const { return S + kLI + k * IS; }
  __device__ __forceinline__ double &zi(int k, int t) const { return Ib(k)[(islot + t) * 2]; }
  __device__ __forceinline__ double &yi(int k, int t) const { return Ib(k)[(islot + t) * 2 + 1]; }
  __device__ __forceinline__ double &si(int k, int t) const { return Ib(k)[NSL * 2 + (islot + t) * 2]; }
  __device__ __forceinline__ double &ui(int k, int t) const { return Ib(k)[NSL * 2 + (islot + t) * 2 + 1]; }
  __device__ __forceinline__ double &pm(int k, int cu) const { return Ib(k)[OPM + cu]; }        // couples u_{k-1}[cu], u_k[cu]
  __device__ __forceinline__ int ci(int k, int t) const { return k * 8 + islot + t; }           // slab slot of my row
};

// Iterates a cold per-stage loop over my stages; `kv` is false on the left half's extra (dummy) trip, where k repeats my
// first stage so that every address stays valid (stores and accumulations are guarded by kv).
#ifndef H16_COLD_UNROLL
#define H16_COLD_UNROLL 1
#endif
#define H16_STR2(x) #x
#define H16_STR(x) H16_STR2(x)
#define H16_COLD_LOOP(i, k, kv)                                   \
  _Pragma(H16_STR(unroll H16_COLD_UNROLL)) for (int i = 0; i <= c.NL; ++i)            \
    if (const bool kv = (i < c.nck); true)                        \
      if (const int k = c.kc0 + (kv ? i : 0); true)
// the two loops of a Ruiz pass (10 passes per QP: the bulk of setup): independent per stage, but unrolling them is slower
#ifndef H16_RUIZ_UNROLL
#define H16_RUIZ_UNROLL 1   // measured (reproducible builds, same box): 5 (full) 1.003 ms against 0.986 ms at 1; H16_COLD_UNROLL 2: 0.995 ms
#endif
#define H16_RUIZ_LOOP(i, k, kv)                                   \
  _Pragma(H16_STR(unroll H16_RUIZ_UNROLL)) for (int i = 0; i <= c.NL; ++i)            \
    if (const bool kv = (i < c.nck); true)                        \
      if (const int k = c.kc0 + (kv ? i : 0); true)

// (P v)_(k, r), (A v) on my dynamics row, (A' t)_(k, r): as in lpv_h8t.cuh
__device__ __forceinline__ double rowP(const Ctx &c, const double *PD, const double *PO, const double *v, int vs, int k) {
  const int N = c.N, o = k * 8 + c.r, ov = k * vs + c.r;
  double acc = PD[o] * v[ov];
  if (c.ul) {
    if (k > 0 && k < N) acc = fma(PO[o - 8], v[ov - vs], acc);
    if (k < N - 1) acc = fma(PO[o], v[ov + vs], acc);
    if (k == N) acc = 0.0;
  }
  return acc;
}
__device__ __forceinline__ double rowA_dyn(const Ctx &c, const double *ED, const double *v, int vs, int k) {
  double acc = ED[k * 8 + c.r] * v[k * vs + c.r];
  if (k > 0) {
    double g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = v[(k - 1) * vs + q];
    acc += h8t::rowdot(c.Gb(k - 1), c.ro, g);
  }
  return acc;
}
__device__ __forceinline__ double colA(const Ctx &c, const double *ED, const double *td, const double *ti, bool tis, int k) {
  const int o = k * 8 + c.r;
  double acc = c.xl ? ED[o] * td[o] : 0.0;
  if (k < c.N) {
    double g[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? td[(k + 1) * 8 + rr] : 0.0;
    acc += h8t::coldot<NX>(c.Gb(k), c.co, g);
  }
  if (c.has_in(k)) {
#pragma unroll
    for (int t = 0; t < NT; ++t) acc = fma(c.si(k, t), tis ? c.yi(k, t) : ti[c.ci(k, t)], acc);
  }
  return acc;
}

// ---------------------------------------------------------------- twisted block factorisation (cold)
// Row weights as in lpv_h8t.cuh (ADMM: rho_eq / rho; polish: 1/delta on the active rows).  Lane (h, r) ends up with row r
// of T, row r and column r of the multiplier (K for the left half, J for the right half) of every local step in tensor
// memory.  In ADMM mode diag(M) goes to DG.
__device__ __noinline__ void factor(const Ctx c, const FW fw, const double sigma) {
  const int N = c.N, NL = c.NL, r = c.r, h = c.h;
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI), *ED = c.cd(C_ED);
  double *DG = c.V(V_DG);
  auto wd = [&](int k) -> double {  // weight of my dynamics row (k, r)
    if (!c.xl) return 0.0;
    return fw.polish ? ((ACTD[k * 8 + r] != 0.0) ? fw.idel : 0.0) : fw.rho_eq;
  };
  __syncwarp();
#pragma unroll 1
  for (int j = 0; j <= NL; ++j) {
    const int k = c.kstage(j);
    const bool mid = (j == NL);
    const int nbk = (k < N) ? NB : NX;
    const bool rowlive = r < nbk;
    double s[8], so[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) { s[cc] = 0.0; so[cc] = 0.0; }
    const double wdk = wd(k);
    const double edk = c.xl ? ED[k * 8 + r] : 0.0;
    const bool diag = !(mid && h);   // the middle stage's own block enters once: through the left half
    {
      double d = PD[k * 8 + r] + sigma;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const double a = c.si(k, t);
          const double w = fw.polish ? ((ACTI[c.ci(k, t)] != 0.0) ? fw.idel : 0.0) : fw.rho;
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
    // weights of the next stage's dynamics rows, one per lane of my half (needed by the diagonal block and, on the
    // right half, by the off-diagonal block)
    const double wn = (k < N) ? wd(k + 1) : 0.0;
    const double fn = (k < N && c.xl) ? wn * ED[(k + 1) * 8 + r] : 0.0;
    {  // next-stage dynamics rows: sum_rr w(k+1, rr) G[rr][r] G[rr][cc]  (the shuffles stay outside the branch)
      const double *g = c.Gb(k < N ? k : N - 1);
#pragma unroll
      for (int rr = 0; rr < NX; ++rr) {
        const double wcol = h8shfl(wn, rr);
        if (k < N && diag) {
          const double col = wcol * g[rr * 8 + c.co[rr >> 1]];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 e = ld2(g + h8t::chunk(rr, q));
            s[2 * q] = fma(col, e.x, s[2 * q]);
            s[2 * q + 1] = fma(col, e.y, s[2 * q + 1]);
          }
        }
      }
    }
    if (j > 0) {
      // my row of the block that couples stage k to the neighbour eliminated before it
      if (!h) {  // left: S_{k,k-1}
        if (c.xl) {
          const double f = wdk * edk;
          const double *gp = c.Gb(k - 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) { const double2 e = ld2(gp + c.ro[q]); so[2 * q] = f * e.x; so[2 * q + 1] = f * e.y; }
        } else if (c.ul && k < N) {
#pragma unroll
          for (int cc = NX; cc < NB; ++cc) if (cc == r) so[cc] = PO[(k - 1) * 8 + r];
        }
      }
      {  // right: S_{k,k+1} = S_{k+1,k}': my row is column r of S_{k+1,k}  (shuffles outside the branch)
        const double *gk = c.Gb(k < N ? k : N - 1);
#pragma unroll
        for (int cc = 0; cc < NX; ++cc) {
          const double f = h8shfl(fn, cc);
          if (h && c.var_live(k)) so[cc] = f * gk[cc * 8 + c.co[cc >> 1]];
        }
        if (h && c.ul && k < N - 1) {
#pragma unroll
          for (int cc = NX; cc < NB; ++cc) if (cc == r) so[cc] = PO[k * 8 + r];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) st2(c.Ob(r) + 2 * q, so[2 * q], so[2 * q + 1]);   // park the block
      __syncwarp();
      double kr[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) kr[cc] = 0.0;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const double *tp = c.Tp(jj);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 e = ld2(tp + 2 * q);
          kr[2 * q] = fma(so[jj], e.x, kr[2 * q]);
          kr[2 * q + 1] = fma(so[jj], e.y, kr[2 * q + 1]);
        }
      }
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        double acc = s[cc];
        const double *ob = c.Ob(cc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 e = ld2(ob + 2 * q);  // block[cc][2q..2q+1]
          acc = fma(-kr[2 * q], e.x, acc);
          acc = fma(-kr[2 * q + 1], e.y, acc);
        }
        s[cc] = acc;
      }
      __syncwarp();
      tm_st8(c.tKr(j), kr);                                                          // my row of the multiplier
#pragma unroll
      for (int q = 0; q < 4; ++q) st2(c.Ob(r) + 2 * q, kr[2 * q], kr[2 * q + 1]);    // transpose through the scratch
      __syncwarp();
      double kc[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) kc[rr] = c.Ob(rr)[r];
      tm_st8(c.tKc(j), kc);                                                          // my column of the multiplier
    }
    if (mid) {  // S~_m = (S_m - K_m S_{m,m-1}') + (- J_m S_{m,m+1}'): the halves exchange their terms
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) s[cc] = s[cc] + xhalf(s[cc]);
    }
    if (!rowlive) {
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) s[cc] = (cc == r) ? 1.0 : 0.0;
    }
    // Gauss-Jordan inverse of the pivot block, rows across the lanes of my half (no pivoting: the block is SPD)
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      double pr[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) pr[cc] = h8shfl(s[cc], p);
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
    tm_st8(c.tT(j), s);                                                              // my row of T
    __syncwarp();  // everybody is done with the previous pivot inverse and with the parked block
#pragma unroll
    for (int q = 0; q < 4; ++q) st2(c.Tp(r) + 2 * q, s[2 * q], s[2 * q + 1]);
    __syncwarp();
  }
  tm_wait_st();
  // the scratch rows of the right half sat in XS (the running sum of x~, zero whenever a factorisation runs in ADMM mode)
  if (!fw.polish && h) {
#pragma unroll 1
    for (int row = 0; row < 8; ++row) c.V(V_XS)[row * VS + r] = 0.0;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- hot loop
// Per-lane addresses of local step 0 and signed strides per local step (computed once per QP).
struct Hot {
  uint32_t tT, tKr, tKc;  // tensor memory: T row of local step 0, multiplier row / column of local step 1 (16 columns per step)
  uint32_t v;             // my element of the stage vectors at local step 0: B +0, X +64, R +128, XS +192, DG +256, CR +320
  uint32_t ib;            // my first single-variable row at local step 0 ({z,y} of row t at +16t, {s,u} at +NSL*16 + 16t)
  uint32_t pm, pm2;       // slew coupling with the stage of local step j+1 / j-1 ... see the kernel (state lanes: the zero pad)
  int vstr, istr, pstr;   // bytes per local step (negative on the right half; 0: dummy row / zero pad)
  uint32_t gpub, ggat;    // all-gather buffer 0 (buffer 1 at ^256): where I publish, where my half's 64 bytes start
  uint32_t vmid, ibmid, pmid, pmid2;   // the middle stage: my element, my rows, coupling with stage m-1 / m+1
};

// forward sweep: B holds the right-hand side on entry, W = T v on exit (both halves, local steps 0 .. NL)
template <bool UNROLL = false>
__device__ __forceinline__ void sweep_fwd(const Hot &h, const int NLr, const int half, uint32_t &gsel) {
  const int NL = UNROLL ? 4 : NLr;
  uint32_t tt = h.tT, tk = h.tKr, vb = h.v;
  TmRow a, b;
  tm_ld8(tt, a); tm_ld8(tk, b);
  double v = lds(vb);
#pragma unroll (UNROLL ? 4 : 1)
  for (int j = 0; j < NL; ++j) {
    sts(h.gpub ^ gsel, v);
    double bn = lds(vb + (uint32_t)h.vstr);
    if (j == NL - 1 && half) bn = 0.0;     // the middle stage's right-hand side enters once: through the left half
    __syncwarp();
    const uint32_t gg = h.ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    tm_wait_ld();
    double c0 = fma(-tm_get(b, 0), g0.x, bn), c1 = -(tm_get(b, 2) * g1.x), c2 = -(tm_get(b, 4) * g2.x), c3 = -(tm_get(b, 6) * g3.x);
    c0 = fma(-tm_get(b, 1), g0.y, c0); c1 = fma(-tm_get(b, 3), g1.y, c1); c2 = fma(-tm_get(b, 5), g2.y, c2); c3 = fma(-tm_get(b, 7), g3.y, c3);
    v = (c0 + c1) + (c2 + c3);
    const double w = dot8(a, g0, g1, g2, g3);
    tt += 16; tk += 16;
    tm_ld8(tt, a); tm_ld8(tk, b);   // next local step (the last multiplier request runs into the column area: harmless, never used)
    sts(vb, w);
    vb += (uint32_t)h.vstr; gsel ^= 256u;
  }
  v = v + xhalf(v);   // v_m = (b_m - K_m v_{m-1}) + (- J_m v_{m+1}): identical in both halves
  {  // the middle stage: x_m = T_m v_m (both halves, same value to the same slot)
    sts(h.gpub ^ gsel, v);
    __syncwarp();
    const uint32_t gg = h.ggat ^ gsel;
    const double2 g0 = lds2(gg), g1 = lds2<64>(gg), g2 = lds2<128>(gg), g3 = lds2<192>(gg);
    tm_wait_ld();
    sts(vb, dot8(a, g0, g1, g2, g3));
    gsel ^= 256u;
  }
}

// backward sweep only (polish): x written back into B
__device__ __forceinline__ void sweep_bwd_plain(const Hot &h, const int NL, uint32_t &gsel) {
  double gn[8];
  uint32_t vb = h.vmid, tc = h.tKc + (uint32_t)((NL - 1) * 16);
  TmRow e;
  tm_ld8(tc, e);
  __syncwarp();   // the middle slot was written by both halves
  {
    const double x = lds(vb);
    sts(h.gpub ^ gsel, x);
    gather_in(h.ggat, gsel, gn);
  }
#pragma unroll 1
  for (int j = NL - 1; j >= 0; --j) {
    vb -= (uint32_t)h.vstr;
    const double w = lds(vb);
    tm_wait_ld();
    const double xt = bwd_step(e, w, gn);
    tc -= 16;
    tm_ld8(j > 0 ? tc : h.tKc, e);   // multiplier column of local step j (j = 0: dummy request, keeps the code converged)
    sts(h.gpub ^ gsel, xt);
    sts(vb, xt);
    gather_in(h.ggat, gsel, gn);
  }
  tm_wait_ld();
  __syncwarp();
}

// backward sweep fused with the element-wise ADMM update (hot).  Walking local steps NL-1 .. 0, the update of the stage
// of local step j+1 runs behind the chain step that produces x~ of local step j.  The middle stage is updated by both
// halves with its neighbours in canonical order (x~_{m-1}, x~_{m+1}): bit-identical values to the same addresses.
template <bool UNROLL = false>
__device__ __forceinline__ void sweep_bwd_admm(const Hot &h, const Upd<KIND> &u, const int NLr, const int half, uint32_t &gsel) {
  const int NL = UNROLL ? 4 : NLr;
  double gn[8];
  uint32_t vb = h.vmid, ib = h.ibmid;
  uint32_t pb = h.pm + (uint32_t)(NL * h.pstr), pb2 = h.pm2 + (uint32_t)(NL * h.pstr);
  uint32_t tc = h.tKc + (uint32_t)((NL - 1) * 16);
  TmRow e;
  tm_ld8(tc, e);
  __syncwarp();   // the middle slot was written by both halves
  double x1 = lds(vb), x2 = 0.0;   // x~_m
  sts(h.gpub ^ gsel, x1);
  gather_in(h.ggat, gsel, gn);
  {  // local step NL-1 (stages m-1 / m+1) and the update of the middle stage
    UpdIn in;
    update_loads<KIND>(vb, ib, ib, h.pmid, h.pmid2, in);
    const double w = lds(vb - (uint32_t)h.vstr);
    tm_wait_ld();
    const double xt = bwd_step(e, w, gn);
    sts(h.gpub ^ gsel, xt);
    tc -= 16;
    tm_ld8(NL > 1 ? tc : h.tKc, e);
    const double xo = xhalf(xt);
    UpdMid q;
    update_part1<KIND>(u, 0, ib, in, x1, half ? xo : xt, half ? xt : xo, q);
    gather_in(h.ggat, gsel, gn);
    update_part2<KIND>(u, vb, x1, q);
    vb -= (uint32_t)h.vstr; ib -= (uint32_t)h.istr; pb -= (uint32_t)h.pstr; pb2 -= (uint32_t)h.pstr;
    x2 = x1; x1 = xt;
  }
#pragma unroll (UNROLL ? 3 : 1)
  for (int j = NL - 2; j >= 0; --j) {
    UpdIn in;
    update_loads<KIND>(vb, ib, ib, pb, pb2, in);   // local step j+1, independent of the chain below
    const double w = lds(vb - (uint32_t)h.vstr);
    tm_wait_ld();
    const double xt = bwd_step(e, w, gn);
    sts(h.gpub ^ gsel, xt);
    tc -= 16;
    tm_ld8(j > 0 ? tc : h.tKc, e);
    UpdMid q;
    update_part1<KIND>(u, 0, ib, in, x1, xt, x2, q);   // fills the publish -> gather latency
    gather_in(h.ggat, gsel, gn);
    update_part2<KIND>(u, vb, x1, q);
    vb -= (uint32_t)h.vstr; ib -= (uint32_t)h.istr; pb -= (uint32_t)h.pstr; pb2 -= (uint32_t)h.pstr;
    x2 = x1; x1 = xt;
  }
  tm_wait_ld();
  {  // local step 0 (stage 0 / stage N): no further neighbour
    UpdIn in;
    update_loads<KIND>(vb, ib, ib, pb, pb2, in);
    UpdMid q;
    update_part1<KIND>(u, 0, ib, in, x1, 0.0, x2, q);
    update_part2<KIND>(u, vb, x1, q);
  }
  __syncwarp();
}

// ---------------------------------------------------------------- cold routines (stage loops split over the halves)
// y_dyn <- y_dyn + rho_eq (alpha A_dyn XS - (alpha n + (1 - alpha) first) be);  XS <- 0
__device__ __noinline__ void sync_yd(const Ctx c, const bool live, const double rho_eq, const double alpha, const int n, const int first) {
  const int r = c.r;
  double *XS = c.V(V_XS);
  double *YD = c.cd(C_YD);
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED);
  const double cb = alpha * n + (first ? (1.0 - alpha) : 0.0);
  __syncwarp();
  if (n > 0) {
    H16_COLD_LOOP(i, k, kv) {
      if (c.xl && kv) {
        const int o = k * 8 + r;
        const double ax = rowA_dyn(c, ED, XS, VS, k);
        if (live) YD[o] = YD[o] + rho_eq * (alpha * ax - cb * BE[o]);
      }
    }
    __syncwarp();
    H16_COLD_LOOP(i, k, kv) { if (kv) XS[k * VS + r] = 0.0; }
  }
  __syncwarp();
}

// CR = rho_eq A_dyn' be (when `new_cr`);  R = zsel CR - A_dyn' y_dyn - q;  B = sigma x + R + A_in'(rho z - y)
__device__ __noinline__ void reproject(const Ctx c, const bool doit, const double rho, const double rho_eq, const double sigma,
                                      const double zsel, const bool new_cr) {
  const int N = c.N, r = c.r;
  double *BV = c.V(V_B), *R = c.V(V_R), *CR = c.V(V_CR);
  const double *X = c.V(V_X);
  const double *YD = c.cd(C_YD), *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
  (void)rho_eq;
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r, ov = k * VS + r;
    if (kv) {
      if (c.var_live(k)) {
        const double *gk = c.Gb(k);
        double crd = CR[ov];
        if (new_cr) {
          double acc = c.xl ? ED[o] * BE[o] : 0.0;
          if (k < N) {
            double g[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? BE[(k + 1) * 8 + rr] : 0.0;
            acc += h8t::coldot<NX>(gk, c.co, g);
          }
          crd = rho_eq * acc;
        }
        double aty = c.xl ? ED[o] * YD[o] : 0.0;
        if (k < N) {
          double g[8];
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? YD[(k + 1) * 8 + rr] : 0.0;
          aty += h8t::coldot<NX>(gk, c.co, g);
        }
        double sin = 0.0;
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) sin = fma(c.si(k, t), rho * c.zi(k, t) - c.yi(k, t), sin);
        }
        if (doit) {
          const double rr = (zsel * crd - aty) - QV[o];
          CR[ov] = crd; R[ov] = rr;
          BV[ov] = fma(sigma, X[ov], rr) + sin;
        }
      } else if (doit) { CR[ov] = 0.0; R[ov] = 0.0; BV[ov] = 0.0; }
    }
  }
  __syncwarp();
}

// residual norms at the current iterate (update_info); y_dyn must be in sync
__device__ __noinline__ void update_info(const Ctx c, Info *ip, const double zsel) {
  Info &I = *ip;
  const int r = c.r;
  const double *X = c.V(V_X);
  const double *YD = c.cd(C_YD), *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
  const double *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV), *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double a_rp = 0, a_z = 0, a_Ax = 0, b_rp = 0, b_z = 0, b_Ax = 0;
  double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, b_rd = 0, b_q = 0, b_Aty = 0, b_Px = 0;
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r;
    if (kv) {
      if (c.xl) {
        const double Ax = rowA_dyn(c, ED, X, VS, k), z = zsel * BE[o], rr = Ax - z, ei = EINV[o];
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
        const double Px = rowP(c, PD, PO, X, VS, k), Aty = colA(c, ED, YD, nullptr, true, k);
        const double rr = (QV[o] + Px) + Aty, di = DINV[o];
        a_rd = absmax(a_rd, rr); a_q = absmax(a_q, QV[o]); a_Aty = absmax(a_Aty, Aty); a_Px = absmax(a_Px, Px);
        b_rd = absmax(b_rd, di * rr); b_q = absmax(b_q, di * QV[o]); b_Aty = absmax(b_Aty, di * Aty); b_Px = absmax(b_Px, di * Px);
      }
    }
  }
  I.n_rp = qmax(a_rp); I.n_z = qmax(a_z); I.n_Ax = qmax(a_Ax); I.n_rd = qmax(a_rd); I.n_q = qmax(a_q); I.n_Aty = qmax(a_Aty); I.n_Px = qmax(a_Px);
  if (I.unscale) {
    I.pri_res = qmax(b_rp); I.u_z = qmax(b_z); I.u_Ax = qmax(b_Ax);
    I.dua_res = I.cinv * qmax(b_rd); I.u_q = qmax(b_q); I.u_Aty = qmax(b_Aty); I.u_Px = qmax(b_Px);
  } else {
    I.pri_res = I.n_rp; I.u_z = I.n_z; I.u_Ax = I.n_Ax; I.dua_res = I.n_rd; I.u_q = I.n_q; I.u_Aty = I.n_Aty; I.u_Px = I.n_Px;
  }
}

// ---------------------------------------------------------------- infeasibility certificates (rare): as in lpv_h8t.cuh
__device__ __noinline__ bool primal_infeasible(const Ctx c, const Info *ip, const double eps, const double rho_eq, const double alpha,
                                              const int last_was_first) {
  const bool unscale = ip->unscale;
  const int r = c.r;
  const double *X = c.V(V_X);
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED);
  const double *PVYI = c.cd(C_PVYI), *E = c.cd(C_E), *EI = c.cd(C_EI), *DINV = c.cd(C_DINV);
  double *DYD = c.cd(C_DYD), *DYI = c.cd(C_PYI), *XT = c.cd(C_ZT);
  const double ia = 1.0 / alpha, oma = 1.0 - alpha, cb = last_was_first ? 1.0 : alpha;
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    double pvx, dsc;
    tm_ld2(c.tQ(i), pvx, dsc);   // the saved iterate lives in tensor memory (collective load: outside any branch)
    if (kv) XT[k * 8 + r] = c.var_live(k) ? (X[k * VS + r] - oma * pvx) * ia : 0.0;
    (void)dsc;
  }
  __syncwarp();
  double nrm = 0.0, lhs = 0.0;
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r;
    if (kv) {
      double d = 0.0;
      if (c.xl) d = rho_eq * (alpha * rowA_dyn(c, ED, XT, 8, k) - cb * BE[o]);  // equality rows: no projection
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
          const double lo = -kInfty, up = c.ui(k, t);
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
  }
  nrm = qmax(nrm);
  lhs = qsum(lhs);
  __syncwarp();
  // the product with A' is only needed when the first two conditions of the certificate hold for some QP of the warp
  if (!__any_sync(kFull, (nrm > eps) && (lhs < -eps * nrm))) return false;
  double mx = 0.0;
  H16_COLD_LOOP(i, k, kv) {
    if (kv && c.var_live(k)) {
      const double at = colA(c, ED, DYD, DYI, false, k);
      mx = absmax(mx, unscale ? DINV[k * 8 + r] * at : at);
    }
  }
  mx = qmax(mx);
  return (nrm > eps) && (lhs < -eps * nrm) && (mx < eps * nrm);
}

__device__ __noinline__ bool dual_infeasible(const Ctx c, const Info *ip, const double eps) {
  const bool unscale = ip->unscale;
  const int r = c.r;
  const double *X = c.V(V_X);
  const double *QV = c.cd(C_Q), *ED = c.cd(C_ED);
  const double *DINV = c.cd(C_DINV), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV);
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO);
  double *DX = c.cd(C_PX);
  double nrm = 0.0, qdx = 0.0;
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r;
    double pvx, dsc;
    tm_ld2(c.tQ(i), pvx, dsc);   // saved iterate and scaling D from tensor memory
    if (kv) {
      const double dx = c.var_live(k) ? X[k * VS + r] - pvx : 0.0;
      DX[o] = dx;
      nrm = absmax(nrm, unscale ? dsc * dx : dx);
      qdx += QV[o] * dx;
    }
  }
  nrm = qmax(nrm); qdx = qsum(qdx);
  __syncwarp();
  const double cs = unscale ? ip->csc : 1.0;
  if (!__any_sync(kFull, (nrm > eps) && (qdx < -cs * eps * nrm))) return false;
  double mx = 0.0;
  int viol = 0;
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r;
    if (kv) {
      if (c.var_live(k)) {
        const double Pdx = rowP(c, PD, PO, DX, 8, k);
        mx = absmax(mx, unscale ? DINV[o] * Pdx : Pdx);
      }
      if (c.xl) {
        double v = rowA_dyn(c, ED, DX, 8, k);
        if (unscale) v = EINV[o] * v;
        if (v > eps * nrm || v < -eps * nrm) viol = 1;
      }
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          double v = c.si(k, t) * DX[o];
          if (unscale) v = EIINV[c.ci(k, t)] * v;
          if ((c.ui(k, t) < kInfty * kMinScaling) && (v > eps * nrm)) viol = 1;   // lower bounds are -OSQP_INFTY
        }
      }
    }
  }
  mx = qmax(mx);
  viol = qany(viol);
  return (nrm > eps) && (qdx < -cs * eps * nrm) && (mx < cs * eps * nrm) && !viol;
}

// returns 1 when a termination status was set for my QP (check_termination)
__device__ __noinline__ int check_termination(const Ctx c, const lpvmpc_settings &S, Info *ip, const bool live, const int approximate,
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
    if (__any_sync(kFull, live && !ncvx && !prim_ok)) prim_inf = primal_infeasible(c, ip, eps_pi, rho_eq, S.alpha, last_was_first) && !prim_ok;
    if (__any_sync(kFull, live && !ncvx && !dual_ok)) dual_inf = dual_infeasible(c, ip, eps_di) && !dual_ok;
  }
  if (!live) return 0;
  if (ncvx) { I.status = LPVMPC_NON_CVX; I.obj = nan(""); return 1; }
  if (prim_ok && dual_ok) { I.status = approximate ? LPVMPC_SOLVED_INACCURATE : LPVMPC_SOLVED; return 1; }
  if (prim_inf) { I.status = approximate ? LPVMPC_PRIMAL_INFEASIBLE_INACCURATE : LPVMPC_PRIMAL_INFEASIBLE; I.obj = kInfty; return 1; }
  if (dual_inf) { I.status = approximate ? LPVMPC_DUAL_INFEASIBLE_INACCURATE : LPVMPC_DUAL_INFEASIBLE; I.obj = -kInfty; return 1; }
  return 0;
}

__device__ __noinline__ double objective(const Ctx c, const double *xv, const int vs, const double scale) {
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *QV = c.cd(C_Q);
  double acc = 0.0;
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    if (kv && c.var_live(k)) acc += (0.5 * rowP(c, PD, PO, xv, vs, k) + QV[k * 8 + c.r]) * xv[k * vs + c.r];
  }
  return qsum(acc) * scale;
}

// ---------------------------------------------------------------- setup: schedule + build + Ruiz (cold, once per QP)
// As in lpv_h8t.cuh, the stage loops split over the halves.  Returns flags: bit 0 = Curvature() failed, bit 1 = bad data.
__device__ __noinline__ int setup(const Ctx c, const H8Params &p, const int b, const bool valid, double *csc_out) {
  const int N = c.N, r = c.r, h = c.h;
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &St = p.S;
  double *S = c.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N;
  const int ucomp = r - NX;
  int sched_err = 0, data_err = 0;
  double x0r = 0.0;
  double *sPD = c.cd(C_PD), *sPO = c.cd(C_PO), *sD = c.cd(C_Q), *sE = c.cd(C_BE), *sEI = c.cd(C_ED), *sDt = c.cd(C_YD);
  double *sEt = c.cd(C_DINV), *sEti = c.cd(C_EINV);
  double *Gs = S + kLG;
  __syncwarp();
  // ---- schedule: G_k = -[A_k B_k] (unscaled); every lane walks the serial roll-out, the owner of stage k keeps row r
  if (a.sched_mode == LPVMPC_SCHED_GIVEN) {
    H16_COLD_LOOP(i, k, kv) {
      if (c.xl && kv && k < N) {
        const double *Ar = a.A + ((size_t)b * N + k) * NX * NX + r * NX, *Br = a.Bm + ((size_t)b * N + k) * NX * 2 + r * 2;
        double row[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) row[cc] = (cc < NX) ? -Ar[cc < NX ? cc : 0] : ((cc < NB) ? -Br[cc - NX < 2 ? cc - NX : 0] : 0.0);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) if (notfinite(row[cc])) data_err = 1;
        double *gk = Gs + k * GS;
#pragma unroll
        for (int q = 0; q < 4; ++q) st2(gk + c.ro[q], row[2 * q], row[2 * q + 1]);
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
      if (c.xl && c.owns(k)) {
        double *gk = Gs + k * GS;
#pragma unroll
        for (int q = 0; q < 4; ++q) st2(gk + c.ro[q], -row[2 * q], -row[2 * q + 1]);
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
          if (valid && a.states_out && c.owns(k)) a.states_out[((size_t)b * N + k) * NX + r] = mine;
          if (k == 0 && a.x0_from_prediction) x0r = mine;
        }
      }
    }
  }
  __syncwarp();
  sched_err = qany(sched_err);

  // ---- build (PathFollowingLPVMPC.py:334-348, 397-464)
  double *X = c.V(V_X), *BV = c.V(V_B), *QV = c.V(V_R), *BE = c.V(V_CR), *ED = c.V(V_XS);   // q, be, ed: scratch homes [k*VS + r]
  {
    const double Qrr = c.xl ? M.Q[r * NX + r] : 0.0;
    const double Q0r = c.xl ? M.Q[r] : 0.0;
    const double Rcc = c.ul ? M.R[ucomp * 2 + ucomp] : 0.0;
    const double dRc = c.ul ? M.dR[ucomp] : 0.0;
    const double uold = (c.ul && a.u_old) ? a.u_old[(size_t)b * 2 + ucomp] : 0.0;
    H16_COLD_LOOP(i, k, kv) {
      const int o = k * 8 + r, ov = k * VS + r;
      if (kv) {
        double pd = 0.0, po = 0.0, q = 0.0, be = 0.0, ed = 0.0;
        if (c.xl) {
          pd = 2 * Qrr;
          q = -2 * (a.vel_ref[(size_t)b * (N + 1) + k] * Q0r);
          be = (k == 0) ? (x0r + 0.0) : (0.0 + (a.C ? a.C[((size_t)b * N + (k - 1)) * NX + r] : 0.0));
          ed = 1.0;
        } else if (c.ul) {
          double v = Rcc + 2 * dRc;
          if (k == N - 1) v = v - dRc;
          pd = (k < N) ? 2 * v : 0.0;
          q = (k == 0) ? -2 * (uold * dRc) : -2 * 0.0;
          po = (k < N - 1) ? 2 * (-dRc) : 0.0;
        }
        if (notfinite(q) || notfinite(be)) data_err = 1;   // NaN / inf in x0, C, vel_ref or u_old
        sPD[o] = pd; sPO[o] = po; QV[ov] = q; BE[ov] = be; ED[ov] = ed;
        sD[o] = 1.0; sE[o] = 1.0; sEI[o] = 1.0; sEti[o] = 1.0;
        X[ov] = 0.0; BV[ov] = 0.0;
      }
    }
    // every single-variable-row slot starts as a dummy row; block N+1 is all dummies (no tensor memory here: any split)
#pragma unroll 1
    for (int k = h; k <= N + 1; k += 2) {
      double *ibk = c.Ib(k);
      if (r < NSL) { ibk[r * 2] = 0.0; ibk[r * 2 + 1] = 0.0; ibk[NSL * 2 + r * 2] = 0.0; ibk[NSL * 2 + r * 2 + 1] = kInfty * kInfty; }
      if (r < 2) ibk[OPM + r] = 0.0;
    }
    __syncwarp();
    H16_COLD_LOOP(i, k, kv) {
      if (kv && c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          double si, up;
          if (r == 0) { si = t ? 1.0 : -1.0; up = t ? M.max_vel : -0.01; }
          else if (r == NX) { si = t ? -1.0 : 1.0; up = 0.249; }
          else { si = t ? -1.0 : 1.0; up = t ? 1.0 : 4.0; }
          up = (up < kInfty || up != up) ? up : kInfty;   // python wrapper: u = min(u, OSQP_INFTY); NaN stays NaN
          if (!(-kInfty <= up)) data_err = 1;             // l > u or a NaN bound
          c.si(k, t) = si; c.ui(k, t) = up; c.zi(k, t) = 0.0; c.yi(k, t) = 0.0;
        }
      }
    }
    __syncwarp();
  }
  data_err = qany(data_err);

  // ---- Ruiz equilibration (OSQP scale_data); see lpv_h8t.cuh for the fusions (pending cost scale `cp`, column norms taken
  // while the previous pass scales the columns)
  double csc = 1.0, cp = 1.0;
  double *sPC = c.V(V_DG), *sQA = c.V(V_B);
  const int oL = c.NL * 8 + r;   // my slot at the right half's first stage
#pragma unroll 1
  for (int it = 0; it < St.scaling; ++it) {
    H16_RUIZ_LOOP(i, k, kv) {
      const int o = k * 8 + r, ov = k * VS + r;
      if (kv) {
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
          for (int q = 0; q < 4; ++q) { const double2 e = ld2(gp + c.ro[q]); ea = absmax(ea, e.x); ea = absmax(ea, e.y); }
        }
        sEt[o] = frsqrt(limit_scaling(ea));
      }
    }
    __syncwarp();
    double qn = 0.0, ct = 0.0, npo_prev = 0.0;
    // the right half's first stage needs the scaled slew coupling of the stage before it (the left half's last): same
    // products as the left half forms, from the values before this pass touches them
    if (h && c.ul) npo_prev = ((sPO[oL - 8] * cp) * sDt[oL - 8]) * sDt[oL];
    __syncwarp();
    H16_RUIZ_LOOP(i, k, kv) {
      const int o = k * 8 + r, ov = k * VS + r;
      if (kv) {
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
        double pc = fabs(npd);
        if (c.ul) {
          if (k < N - 1) pc = absmax(pc, npo);
          if (k > 0 && k < N) pc = absmax(pc, npo_prev);
        }
        if (c.var_live(k)) { ct += pc; qn = absmax(qn, nq); }
        sPC[ov] = pc;
        npo_prev = npo;
      }
    }
    __syncwarp();
    qn = qmax(qn);
    ct = qsum(ct) / nz;
    qn = limit_scaling(qn);
    ct = ct > qn ? ct : qn;
    ct = limit_scaling(ct);
    ct = frcp(ct);
    cp = ct;
    csc *= ct;
  }
  if (St.scaling > 0) {
    H16_COLD_LOOP(i, k, kv) { if (kv) { const int o = k * 8 + r; sPD[o] *= cp; QV[k * VS + r] *= cp; sPO[o] *= cp; } }
    __syncwarp();
  }
  *csc_out = csc;
  // ---- scaled bounds; then the work vectors give way to the cold vectors they were parked in
  {
    H16_COLD_LOOP(i, k, kv) {
      if (kv && c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) { const int oc = c.ci(k, t); c.ui(k, t) = sEI[oc] * c.ui(k, t); }
      }
    }
    __syncwarp();
    double *cD = c.cd(C_D), *cE = c.cd(C_E), *cEi = c.cd(C_EI), *cEiI = c.cd(C_EIINV);
    double *cQ = c.cd(C_Q), *cBE = c.cd(C_BE), *cED = c.cd(C_ED), *cYD = c.cd(C_YD), *cDI = c.cd(C_DINV), *cEI = c.cd(C_EINV);
    H16_COLD_LOOP(i, k, kv) {
      const int o = k * 8 + r, ov = k * VS + r;
      // element o of every work vector is mine alone: read them all, then overwrite their slots
      const double d = sD[o], e = sE[o], ei = sEI[o], q = QV[ov], be = BE[ov], ed = ED[ov];
      const double po_prev = (c.ul && k > 0 && k < N) ? sPO[o - 8] : 0.0;
      tm_st4(c.tQ(i), 0.0, d, 0.0, 0.0);   // copy of D for the dual-infeasibility certificate (slot +2), saved iterate 0
      if (kv) {
        cD[o] = d; cE[o] = e; cEi[o] = ei; cEiI[o] = 1.0 / ei;
        cQ[o] = c.var_live(k) ? q : 0.0; cBE[o] = e * be; cED[o] = ed; cYD[o] = 0.0; cDI[o] = 1.0 / d; cEI[o] = 1.0 / e;
        if (c.ul) c.pm(k, ucomp) = po_prev;   // couples u_{k-1}, u_k
        c.V(V_XS)[ov] = 0.0;   // the scratch homes become hot vectors (R, CR, B are set by reproject, DG by factor)
        c.V(V_B)[ov] = 0.0;
      }
    }
    tm_wait_st();
    __syncwarp();
  }
  return (sched_err ? 1 : 0) | (data_err ? 2 : 0);
}

// ---------------------------------------------------------------- polish (cold, once per QP): the scheme of lpv_h8t.cuh
__device__ __noinline__ int polish(const Ctx c, const Hot h, const lpvmpc_settings &St, Info *ip, const bool do_pol, uint32_t gsel) {
  Info &I = *ip;
  const bool unscale = I.unscale;
  const int N = c.N, NL = c.NL, r = c.r;
  double *X = c.V(V_X), *BV = c.V(V_B);
  double *PX = c.V(V_R), *PYD = c.V(V_XS), *DX = c.V(V_DG), *TMP = c.V(V_CR);       // [k*VS + q]: tx, ty (dynamics rows), dx, row temporary
  double *YD = c.cd(C_YD);
  const double *BE = c.cd(C_BE), *ED = c.cd(C_ED), *QV = c.cd(C_Q);
  double *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
  const double *PD = c.cd(C_PD), *PO = c.cd(C_PO), *EINV = c.cd(C_EINV), *EIINV = c.cd(C_EIINV), *DINV = c.cd(C_DINV);
  const double delta = St.delta, idel = 1.0 / St.delta;
  uint32_t acti = 0;   // 2 bits per row (i, t) of mine (cold slot i): 1 = lower, 2 = upper, 3 = both
  uint32_t actd = 0;   // bit i: my dynamics row of cold slot i is in the polish system
  auto ai_of = [&](int i, int t) { return (int)((acti >> (2 * (i * NT + t))) & 3u); };
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    if (kv) {
      const int o = k * 8 + r;
      double ad = 0.0;
      if (c.xl && do_pol) ad = (0.0 < YD[o]) ? 2.0 : 1.0;   // every dynamics row stays in the system (see lpv_h8t.cuh)
      ACTD[o] = ad;
      if (ad != 0.0) actd |= 1u << i;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          int ai = 0;
          if (do_pol) { if (c.zi(k, t) - (-kInfty) < -c.yi(k, t)) ai += 1; if (c.ui(k, t) - c.zi(k, t) < c.yi(k, t)) ai += 2; }
          ACTI[c.ci(k, t)] = (double)ai;
          acti |= (uint32_t)ai << (2 * (i * NT + t));
        }
      }
    }
  }
  __syncwarp();
  FW fw; fw.polish = 1; fw.rho = 0.0; fw.rho_eq = 0.0; fw.idel = idel;
  factor(c, fw, delta);
  auto bred_i = [&](int i, int k, int t) { const int a = ai_of(i, t); return (a == 1 || a == 3) ? -kInfty : c.ui(k, t); };
  auto colAt = [&](const double *td, const double ti0, const double ti1, int k) {
    double acc = c.xl ? ED[k * 8 + r] * td[k * VS + r] : 0.0;
    if (k < N) {
      double g[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) g[rr] = (rr < NX) ? td[(k + 1) * VS + rr] : 0.0;
      acc += h8t::coldot<NX>(c.Gb(k), c.co, g);
    }
    if (c.has_in(k)) { acc = fma(c.si(k, 0), ti0, acc); acc = fma(c.si(k, 1), ti1, acc); }
    return acc;
  };
  auto solve = [&]() {
    __syncwarp();
    sweep_fwd(h, NL, c.h, gsel);
    sweep_bwd_plain(h, NL, gsel);
  };
  auto inner = [&]() {
#pragma unroll 1
    for (int ii = 0; ii < kPolishInner; ++ii) {
      H16_COLD_LOOP(i, k, kv) {
        const int o = k * 8 + r, ov = k * VS + r;
        TmQuad pq;
        tm_ld4(c.tP(i), pq);
        tm_wait_ld();
        if (kv) {
          double b = 0.0;
          if (c.var_live(k)) b = ((-QV[o] - rowP(c, PD, PO, PX, VS, k)) - colAt(PYD, tm_getq(pq, 0), tm_getq(pq, 1), k)) - delta * DX[ov];
          BV[ov] = b;
        }
      }
      solve();
      H16_COLD_LOOP(i, k, kv) {
        const int ov = k * VS + r;
        TmQuad pq;
        tm_ld4(c.tP(i), pq);
        tm_wait_ld();
        const double ddx = BV[ov];
        double py[2] = {tm_getq(pq, 0), tm_getq(pq, 1)};
        if (kv) {
          if (c.xl && ((actd >> i) & 1u)) PYD[ov] += idel * rowA_dyn(c, ED, BV, VS, k);
          if (c.has_in(k)) {
#pragma unroll
            for (int t = 0; t < NT; ++t) if (ai_of(i, t) != 0) py[t] += idel * (c.si(k, t) * ddx);
          }
        }
        tm_st4(c.tP(i), py[0], py[1], 0.0, 0.0);
        if (kv) { PX[ov] += ddx; DX[ov] += ddx; }
      }
      tm_wait_st();
      __syncwarp();
    }
  };
  // ---- s_0 = K_reg^-1 (-q, b_red): rhs = -q + A_red'(b_red / delta)
  H16_COLD_LOOP(i, k, kv) { if (kv) TMP[k * VS + r] = (c.xl && ((actd >> i) & 1u)) ? idel * BE[k * 8 + r] : 0.0; }
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    if (kv) {
      double t2[2] = {0.0, 0.0};
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) t2[t] = (ai_of(i, t) != 0) ? idel * bred_i(i, k, t) : 0.0;
      }
      BV[k * VS + r] = c.var_live(k) ? (-QV[k * 8 + r] + colAt(TMP, t2[0], t2[1], k)) : 0.0;
    }
  }
  solve();
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r, ov = k * VS + r;
    const double xk = BV[ov];
    double py[2] = {0.0, 0.0};
    if (kv) {
      PX[ov] = xk; DX[ov] = xk;
      PYD[ov] = (c.xl && ((actd >> i) & 1u)) ? idel * (rowA_dyn(c, ED, BV, VS, k) - BE[o]) : 0.0;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) py[t] = (ai_of(i, t) != 0) ? idel * (c.si(k, t) * xk - bred_i(i, k, t)) : 0.0;
      }
    }
    tm_st4(c.tP(i), py[0], py[1], 0.0, 0.0);
  }
  tm_wait_st();
  __syncwarp();
  inner();
#pragma unroll 1
  for (int it = 0; it < St.polish_refine_iter; ++it) {
    H16_COLD_LOOP(i, k, kv) {
      if (kv) {
        const int o = k * 8 + r, ov = k * VS + r;
        TMP[ov] = (c.xl && ((actd >> i) & 1u)) ? fma(-idel, BE[o] - rowA_dyn(c, ED, PX, VS, k), PYD[ov]) : 0.0;
      }
    }
    __syncwarp();
    H16_COLD_LOOP(i, k, kv) {
      const int o = k * 8 + r, ov = k * VS + r;
      TmQuad pq;
      tm_ld4(c.tP(i), pq);
      tm_wait_ld();
      if (kv) {
        double b = 0.0;
        if (c.var_live(k)) {
          double t2[2] = {0.0, 0.0};
          if (c.has_in(k)) {
#pragma unroll
            for (int t = 0; t < NT; ++t)
              t2[t] = (ai_of(i, t) != 0) ? fma(-idel, bred_i(i, k, t) - c.si(k, t) * PX[ov], tm_getq(pq, t)) : 0.0;
          }
          b = (-QV[o] - rowP(c, PD, PO, PX, VS, k)) - colAt(TMP, t2[0], t2[1], k);
        }
        BV[ov] = b;
      }
    }
    solve();
    H16_COLD_LOOP(i, k, kv) { if (kv) { const int ov = k * VS + r; const double ddx = BV[ov]; DX[ov] = ddx; PX[ov] += ddx; } }
    __syncwarp();
    H16_COLD_LOOP(i, k, kv) {
      const int o = k * 8 + r, ov = k * VS + r;
      TmQuad pq;
      tm_ld4(c.tP(i), pq);
      tm_wait_ld();
      double py[2] = {tm_getq(pq, 0), tm_getq(pq, 1)};
      if (kv) {
        if (c.xl && ((actd >> i) & 1u)) PYD[ov] += idel * (rowA_dyn(c, ED, PX, VS, k) - BE[o]);
        if (c.has_in(k)) {
#pragma unroll
          for (int t = 0; t < NT; ++t) if (ai_of(i, t) != 0) py[t] += idel * (c.si(k, t) * PX[ov] - bred_i(i, k, t));
        }
      }
      tm_st4(c.tP(i), py[0], py[1], 0.0, 0.0);
    }
    tm_wait_st();
    __syncwarp();
    inner();
  }
  // pol z = A x, normal-cone projection, residuals, acceptance.  Polished z of the single-variable rows -> columns 2, 3.
  double a_rp = 0, a_rd = 0;
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r, ov = k * VS + r;
    TmQuad pq;
    tm_ld4(c.tP(i), pq);
    tm_wait_ld();
    double py[2] = {tm_getq(pq, 0), tm_getq(pq, 1)}, zz[2] = {0.0, 0.0};
    if (kv) {
      if (c.xl) {
        const double Ax = rowA_dyn(c, ED, PX, VS, k), t = Ax + PYD[ov];
        TMP[ov] = t - BE[o];
        const double rr = Ax - BE[o];
        a_rp = absmax(a_rp, unscale ? EINV[o] * rr : rr);
      } else TMP[ov] = 0.0;
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int oc = c.ci(k, t);
          const double ax = c.si(k, t) * PX[ov], tt = ax + py[t];
          const double zc = clampd(tt, -kInfty, c.ui(k, t));
          zz[t] = zc; py[t] = tt - zc;
          const double rr = ax - zc;
          a_rp = absmax(a_rp, unscale ? EIINV[oc] * rr : rr);
        }
      }
    }
    tm_st4(c.tP(i), py[0], py[1], zz[0], zz[1]);
  }
  tm_wait_st();
  __syncwarp();
  H16_COLD_LOOP(i, k, kv) {
    const int o = k * 8 + r;
    TmQuad pq;
    tm_ld4(c.tP(i), pq);
    tm_wait_ld();
    if (kv && c.var_live(k)) {
      const double rr = (QV[o] + rowP(c, PD, PO, PX, VS, k)) + colAt(TMP, tm_getq(pq, 0), tm_getq(pq, 1), k);
      a_rd = absmax(a_rd, unscale ? DINV[o] * rr : rr);
    }
  }
  const double pol_pri = qmax(a_rp), pol_dua = (unscale ? I.cinv : 1.0) * qmax(a_rd);
  const double pol_obj = objective(c, PX, VS, St.scaling ? I.cinv : 1.0);
  const bool ok = (pol_pri < I.pri_res && pol_dua < I.dua_res) || (pol_pri < I.pri_res && I.dua_res < 1e-10) ||
                  (pol_dua < I.dua_res && I.pri_res < 1e-10);
  const bool take = do_pol && ok;
  if (take) { I.obj = pol_obj; I.pri_res = pol_pri; I.dua_res = pol_dua; }
  H16_COLD_LOOP(i, k, kv) {   // the tensor-memory loads are collective: every lane walks the loop, only `take` QPs store
    const int o = k * 8 + r, ov = k * VS + r;
    TmQuad pq;
    tm_ld4(c.tP(i), pq);
    tm_wait_ld();
    if (take && kv) {
      X[ov] = PX[ov]; YD[o] = TMP[ov];
      if (c.has_in(k)) {
#pragma unroll
        for (int t = 0; t < NT; ++t) { c.zi(k, t) = tm_getq(pq, 2 + t); c.yi(k, t) = tm_getq(pq, t); }
      }
    }
  }
  __syncwarp();
  if (!do_pol) return 0;
  return ok ? 1 : -1;
}

// ---------------------------------------------------------------- persistent warps, 2 QPs at a time each
__global__ void __launch_bounds__(256, 1) lpv_solve_h16t_kernel(const __grid_constant__ H8Params p) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int gq0 = lane >> 4, half = (lane >> 3) & 1, r = lane & 7, g8 = lane >> 3;
  const Lay &L = p.L;
  constexpr int N = Ctx::N, NL = Ctx::NL;
  Ctx c;
  c.S = smem; c.cold = p.cold;
  c.r = r; c.h = half;
  c.kc0 = half ? NL : 0; c.nck = half ? NL + 1 : NL;
  c.xl = r < NX; c.ul = (r >= NX) && (r < NB);
  c.islot = (r == 0) ? 0 : ((r >= NX) ? (r - NX + 1) * 2 : 0);
#pragma unroll
  for (int j = 0; j < 4; ++j) { c.ro[j] = h8t::chunk(r, j); c.co[j] = (((r >> 1) ^ j) << 1) | (r & 1); }
  const lpvmpc_args &a = p.a;
  const lpvmpc_settings &S = p.S;
  const int nx = NX * (N + 1), nz = nx + 2 * N;
  const int m = 6 * N + nx;
  const int ucomp = r - NX;
  double *wsm = smem + (size_t)warp * 2 * L.total;
  const size_t wslot = (size_t)(blockIdx.x * wpc + warp) * 2;
  // all-gather buffers: two 256-byte buffers per warp (4 halves x 8 doubles), 512-byte aligned, after the QP regions
  const uint32_t smem_a = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t gbuf = ((smem_a + (uint32_t)(wpc * 2 * L.total * 8) + 511u) & ~511u) + (uint32_t)warp * 512u;
  uint32_t gsel = 0;
  // tensor memory: the whole SM's 512 columns (one CTA per SM); warp w owns lanes 32 (w % 4) .. +31, columns 256 (w / 4) .. +255
  __shared__ uint32_t tmem_base_s;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  c.tm = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));

  __shared__ unsigned round_base_s;
  for (;;) {
    unsigned base = 0;
    if (p.cta_rounds) {
      // One round: every warp of the CTA takes its two QPs at the same time.  The kernel is 280 KB of SASS against a 32 KB
      // instruction cache per SM; warps that drift into different phases (setup / ADMM loop / polish) evict each other's
      // code (profiles/r3o_*: 8.2 stall cycles per issue waiting for instructions on long batches, instruction-cache hit
      // rate 56 %), warps that start together stay in the same code for most of a round.
      __syncthreads();
      if (threadIdx.x == 0) round_base_s = atomicAdd(p.queue, 2u * (unsigned)wpc);
      __syncthreads();
      const unsigned rb = round_base_s, rend = (unsigned)p.B;
      if ((int)rb >= p.B) break;
      base = rb + 2u * (unsigned)warp;
      if (base >= rend) {
        if (p.cta_rounds > 1) __syncthreads();   // the phase barrier in front of the polish (below)
        continue;
      }
    } else {
      if (lane == 0) base = atomicAdd(p.queue, 2u);
      base = __shfl_sync(kFull, base, 0);
      if ((int)base >= p.B) break;
    }
    // the second QP slot of a warp at the batch tail mirrors the first: same problem, same shared / slab region, same values
    // written by the same instruction; only user-visible outputs are guarded
    const bool valid = (int)(base + gq0) < p.B;
    const int gq = valid ? gq0 : 0;
    const int b = p.perm ? p.perm[(int)base + gq] : (int)base + gq;
    c.S = wsm + gq * L.total;
    c.cold = p.cold + (wslot + gq) * L.cold_total;

    Hot h;
    {
      const uint32_t sq = smem_a + (uint32_t)((warp * 2 + gq) * L.total) * 8u;
      constexpr int ISB = IS * 8;
      const int k0 = half ? N : 0, sgn = half ? -1 : 1;
      h.tT = c.tT(0); h.tKr = c.tKr(1); h.tKc = c.tKc(1);
      h.v = sq + (uint32_t)(kLV + k0 * VS + r) * 8u;
      h.vstr = sgn * VB;
      const bool rows = (r == 0 || c.ul);
      const uint32_t dm = sq + (uint32_t)(kLI + (N + 1) * IS) * 8u;   // block N+1: dummy rows, zero coupling
      // single-variable rows exist at stages 0 .. N-1: local step 0 of the right half (stage N) has the dummy block's
      // layout right behind it, so the signed stride walks from block N (all dummies after setup) down to block m
      h.ib = rows ? sq + (uint32_t)(kLI + k0 * IS + c.islot * 2) * 8u : dm;
      h.istr = rows ? sgn * ISB : 0;
      // slew coupling: pm(k) couples u_{k-1}, u_k.  The update of local step j sees x~ of local step j-1 (the one just
      // computed: argument xm) and of local step j+1 (xp): left half xm = stage k-1 -> pm(k), xp -> pm(k+1);
      // right half xm = stage k+1 -> pm(k+1), xp = stage k-1 -> pm(k)
      const uint32_t pk0 = sq + (uint32_t)(kLI + k0 * IS + OPM + (c.ul ? ucomp : 0)) * 8u;
      h.pm = c.ul ? (half ? pk0 + (uint32_t)ISB : pk0) : dm + (uint32_t)OPM * 8u;
      h.pm2 = c.ul ? (half ? pk0 : pk0 + (uint32_t)ISB) : dm + (uint32_t)OPM * 8u;
      h.pstr = c.ul ? sgn * ISB : 0;
      h.gpub = gbuf + (uint32_t)(64 * (r >> 1) + 16 * g8 + 8 * (r & 1));
      h.ggat = gbuf + (uint32_t)(16 * g8);
      h.vmid = sq + (uint32_t)(kLV + NL * VS + r) * 8u;
      h.ibmid = rows ? sq + (uint32_t)(kLI + NL * IS + c.islot * 2) * 8u : dm;
      const uint32_t pkm = sq + (uint32_t)(kLI + NL * IS + OPM + (c.ul ? ucomp : 0)) * 8u;
      h.pmid = c.ul ? pkm : dm + (uint32_t)OPM * 8u;
      h.pmid2 = c.ul ? pkm + (uint32_t)ISB : dm + (uint32_t)OPM * 8u;
    }

    Info I;
    double csc = 1.0;
    const int flags = setup(c, p, b, valid, &csc);
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
      factor(c, fw, sigma);
    }
    reproject(c, true, rho, rho_eq, sigma, 0.0, true);
    bool live = (flags == 0);
    const bool failed = flags != 0;
    int iter_done = 0, rho_updates = 0;
    int adapt_interval = S.adaptive_rho_interval;
    if (S.adaptive_rho && !adapt_interval) adapt_interval = S.check_termination ? 4 * S.check_termination : 100;
    const int ct = S.check_termination, ai = S.adaptive_rho ? adapt_interval : 0;

    Upd<KIND> u;
    u.sigma = sigma; u.alpha = alpha; u.oma = 1.0 - alpha; u.eqm = 0; u.loosem = 0; u.N = N;
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
          H16_COLD_LOOP(i, k, kv) {
            tm_st1(c.tQ(i), X[k * VS + r]);   // collective: every lane stores (a frozen QP's copy is never read again)
            if (kv && live && c.has_in(k)) {
#pragma unroll
              for (int t = 0; t < NT; ++t) PVYI[c.ci(k, t)] = c.yi(k, t);
            }
          }
          tm_wait_st();
        }
        u.cc = (iter == 0) ? 2.0 : alpha;
#ifndef H16T_UNROLL_SWEEPS
#define H16T_UNROLL_SWEEPS 1
#endif
        sweep_fwd<H16T_UNROLL_SWEEPS != 0>(h, NL, half, gsel);          // N = 8: four local steps, unrolled (the host admits no other horizon)
        sweep_bwd_admm<H16T_UNROLL_SWEEPS != 0>(h, u, NL, half, gsel);
        if (iter == 0) first_in = 1;
        ++nsync;
        zsel = 1.0;
      }
      __syncwarp();
      const int last_was_first = (iter == 1);
      rho_eq_last = rho_eq;
      sync_yd(c, live, rho_eq, alpha, nsync, first_in);
      nsync = 0; first_in = 0;
      const bool can_check = ct && (iter % ct == 0);
      const bool can_adapt = ai && (iter % ai == 0);
      bool new_cr = false;
      checked_last = can_check;
      if (can_check || can_adapt) {
        Info J = I;
        update_info(c, &J, zsel);
        if (live) { I = J; iter_done = iter; }
        if (can_check) {
          if (check_termination(c, S, &I, live, 0, rho_eq_last, last_was_first, true)) live = false;  // frozen: stores are predicated on `live`
        }
        if (can_adapt) {
          const double pr = I.n_rp / ((I.n_z > I.n_Ax ? I.n_z : I.n_Ax) + 1e-10);
          double dn = I.n_q; dn = (I.n_Aty > dn) ? I.n_Aty : dn; dn = (I.n_Px > dn) ? I.n_Px : dn;
          const double dr = I.n_rd / (dn + 1e-10);
          double rho_new = rho * sqrt(pr / (dr + 1e-10));
          rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
          const bool upd = live && ((rho_new > rho * S.adaptive_rho_tolerance) || (rho_new < rho / S.adaptive_rho_tolerance));
          if (__any_sync(kFull, upd)) {
            // QPs that do not update must keep their factor: re-factorising with unchanged rho reproduces it
            if (upd) { rho = rho_new; rho_eq = kRhoEqOverIneq * rho; ++rho_updates; }
            FW fw; fw.polish = 0; fw.rho = rho; fw.rho_eq = rho_eq; fw.idel = 0.0;
            __syncwarp();
            factor(c, fw, sigma);
            new_cr = true;
          }
        }
      }
      // re-project r (and the pending right-hand side) from the explicit iterate: removes the drift of the recursion
      if (iter < S.max_iter && __any_sync(kFull, live)) reproject(c, live, rho, rho_eq, sigma, zsel, new_cr);
    }
    if (!checked_last && __any_sync(kFull, live)) {
      Info J = I;
      update_info(c, &J, zsel);
      if (live) { I = J; iter_done = iter; }
      if (check_termination(c, S, &I, live, 0, rho_eq_last, iter == 1, iter > 0)) live = false;
    }
    {
      const bool unsolved = (I.status == LPVMPC_UNSOLVED);
      if (__any_sync(kFull, unsolved)) {
        if (!check_termination(c, S, &I, unsolved, 1, rho_eq_last, iter == 1, iter > 0) && unsolved) I.status = LPVMPC_MAX_ITER_REACHED;
      }
    }
    const int status = I.status;
    const bool has_sol = !(status == LPVMPC_PRIMAL_INFEASIBLE || status == LPVMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_DUAL_INFEASIBLE || status == LPVMPC_DUAL_INFEASIBLE_INACCURATE ||
                           status == LPVMPC_NON_CVX || status == LPVMPC_SCHEDULE_ERROR || status == LPVMPC_DATA_ERROR);
    {
      const double o = objective(c, c.V(V_X), VS, S.scaling ? I.cinv : 1.0);
      if (has_sol) I.obj = o;
    }
    // row / variable indices in the reference order
    auto ref_dyn = [&](int k) { return 6 * N + k * NX + r; };
    auto ref_in = [&](int k, int t) { return (r == 0) ? (2 * k + t) : (2 * N + 4 * k + 2 * ucomp + t); };
    auto ref_var = [&](int k) { return c.xl ? (k * NX + r) : (nx + k * 2 + ucomp); };
    if (valid && (a.xs || a.zs || a.ys)) {
      const double *X = c.V(V_X);
      const double *YD = c.cd(C_YD), *BE = c.cd(C_BE);
      H16_COLD_LOOP(i, k, kv) {
        if (kv) {
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
    }
    int polish_status = 0;
    // second phase barrier (cta_rounds = 2): nobody enters the polish code while another warp of the CTA still runs the ADMM loop
    if (p.cta_rounds > 1) __syncthreads();
    const bool do_pol = S.polish && status == LPVMPC_SOLVED;
    bool polished_sets = false;
    if (__any_sync(kFull, do_pol)) {
      polish_status = polish(c, h, S, &I, do_pol, gsel);
      polished_sets = do_pol;
    }
    // ---- outputs
    if (valid) {
      const double *X = c.V(V_X);
      const double *YD = c.cd(C_YD), *D = c.cd(C_D), *E = c.cd(C_E), *EI = c.cd(C_EI), *ACTD = c.cd(C_ACTD), *ACTI = c.cd(C_ACTI);
      H16_COLD_LOOP(i, k, kv) {
        if (kv) {
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
      }
      if (r == 0 && half == 0) {
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
}

#undef H16_COLD_LOOP
#undef H16_RUIZ_LOOP

}  // namespace h16t
}  // namespace lpv
