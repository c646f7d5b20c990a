// Kernels and C-ABI (include/lpvmpc.h) of the B200-native batched LPV-MPC QP solver.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda.h>           // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "lpvmpc.h"
#include "lpv_model.cuh"
#include "lpv_qp.cuh"
#ifdef LPVMPC_LEGACY   // earlier kernel generations (variants 2, 3, 4): A/B measurements only, not in the product build
#include "legacy/lpv_t8.cuh"
#include "legacy/lpv_g8.cuh"
#endif
#include "lpv_h8.cuh"
#include "lpv_h8t.cuh"
#include "lpv_h16t.cuh"
#include "lpv_loop.cuh"
#include "lpv_aux.cuh"

namespace lpv {

// ------------------------------------------------------------------------------------------------
// Scheduling inside the solve kernel: fills G = -[A_k B_k] of the warp's workspace.
// Returns (warp-uniform) nonzero when Curvature() failed for any stage.
template <int KIND>
__device__ int schedule_into(QP<KIND> &qp, const Params &p, int b, double *x0_eff) {
  constexpr int NX = QP<KIND>::NX, NB = QP<KIND>::NB;
  const Layout &L = qp.L;
  const Model &M = qp.M;
  const int N = L.N, lane = qp.lane;
  double *G = qp.w + L.G;
  const lpvmpc_args &a = p.a;
  int err = 0;
  if (a.sched_mode == LPVMPC_SCHED_GIVEN) {
    for (int e = lane; e < N * NX * NB; e += 32) {
      const int k = e / (NX * NB), rc = e - k * NX * NB, r = rc / NB, c = rc - r * NB;
      const double v = (c < NX) ? a.A[((size_t)b * N + k) * NX * NX + r * NX + c]
                                : a.Bm[((size_t)b * N + k) * NX * NU + r * NU + (c - NX)];
      G[e] = -v;
    }
    for (int r = lane; r < NX; r += 32) x0_eff[r] = a.x0[(size_t)b * NX + r];
  } else if (a.sched_mode == LPVMPC_SCHED_PREDICT) {
    if (lane == 0) {
      double st[NX], Ai[NX * NX], Bi[NX * NU];
      const double *xs = a.x_sched ? a.x_sched + (size_t)b * NX : a.x0 + (size_t)b * NX;
#pragma unroll
      for (int r = 0; r < NX; ++r) st[r] = xs[r];
      const double *up = a.u_prev + (size_t)b * N * NU;
      const int lap = a.lap ? a.lap[b] : a.lap_all;
      for (int i = 0; i < N; ++i) {
        if (KIND == LPVMPC_CONTROLLER) {
          const double cur = (lap == 0) ? curvature(M.track, M.nseg, st[4], err) : a.curv_ref[(size_t)b * N + i];
          ctrl_stage(M, a.Cf_new, a.Cf_new, a.vel_ref[(size_t)b * (N + 1) + i], st[1], st[3], st[5], cur, up[i * NU], Ai, Bi);
        } else {
          const double cur = curvature(M.track, M.nseg, a.SS[(size_t)b * (N + 1) + i], err);
          plan_stage(M, st[0], st[1], st[3], st[4], cur, up[i * NU], Ai, Bi);
        }
#pragma unroll
        for (int r = 0; r < NX; ++r) {
#pragma unroll
          for (int c = 0; c < NX; ++c) G[(size_t)(i * NX + r) * NB + c] = -Ai[r * NX + c];
#pragma unroll
          for (int c = 0; c < NU; ++c) G[(size_t)(i * NX + r) * NB + NX + c] = -Bi[r * NU + c];
        }
        if (a.A_out) for (int e = 0; e < NX * NX; ++e) a.A_out[((size_t)b * N + i) * NX * NX + e] = Ai[e];
        if (a.B_out) for (int e = 0; e < NX * NU; ++e) a.B_out[((size_t)b * N + i) * NX * NU + e] = Bi[e];
        propagate<NX>(Ai, Bi, up + i * NU, st);
        if (a.states_out) for (int r = 0; r < NX; ++r) a.states_out[((size_t)b * N + i) * NX + r] = st[r];
        if (i == 0) {
#pragma unroll
          for (int r = 0; r < NX; ++r) x0_eff[r] = a.x0_from_prediction ? st[r] : a.x0[(size_t)b * NX + r];
        }
      }
    }
  } else {  // ESTIMATE: stages are independent
    for (int i = lane; i < N; i += 32) {
      double Ai[NX * NX], Bi[NX * NU];
      const double *t = a.traj + ((size_t)b * N + i) * 6;
      const double delta = a.u_prev[((size_t)b * N + i) * NU];
      if (KIND == LPVMPC_CONTROLLER) {
        const double cur = curvature(M.track, M.nseg, t[4], err);
        ctrl_stage(M, M.Cf, M.Cr, t[0], t[1], t[3], t[5], cur, delta, Ai, Bi);
      } else {
        const double cur = curvature(M.track, M.nseg, t[5], err);
        plan_stage(M, t[0], t[1], t[3], t[4], cur, delta, Ai, Bi);
      }
      for (int r = 0; r < NX; ++r) {
        for (int c = 0; c < NX; ++c) G[(size_t)(i * NX + r) * NB + c] = -Ai[r * NX + c];
        for (int c = 0; c < NU; ++c) G[(size_t)(i * NX + r) * NB + NX + c] = -Bi[r * NU + c];
      }
      if (a.A_out) for (int e = 0; e < NX * NX; ++e) a.A_out[((size_t)b * N + i) * NX * NX + e] = Ai[e];
      if (a.B_out) for (int e = 0; e < NX * NU; ++e) a.B_out[((size_t)b * N + i) * NX * NU + e] = Bi[e];
    }
    for (int r = lane; r < NX; r += 32) x0_eff[r] = a.x0[(size_t)b * NX + r];
  }
  __syncwarp();
  return __any_sync(kFull, err);
}

// One warp per QP, persistent over the batch.
template <int KIND, bool SMEM>
__global__ void __launch_bounds__(32) lpv_solve_kernel(const __grid_constant__ Params p) {
  extern __shared__ double smem[];
  constexpr int NX = QP<KIND>::NX;
  const int lane = threadIdx.x;
  const Layout &L = p.L;
  double *w = SMEM ? smem : p.gws + (size_t)blockIdx.x * L.total;
  QP<KIND> qp(L, p.M, w, lane);
  const lpvmpc_args &a = p.a;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    double *x0_eff = w + L.xt;  // free until the first ADMM step
    const int sched_err = schedule_into<KIND>(qp, p, b, x0_eff);
    Outcome out;
    out.status = LPVMPC_UNSOLVED; out.iter = 0; out.rho_updates = 0; out.polish_status = 0;
    out.obj = nan(""); out.pri_res = nan(""); out.dua_res = nan("");
    bool solved_path = false;
    double c = 1.0;
    if (sched_err) out.status = LPVMPC_SCHEDULE_ERROR;
    else {
      qp.build(p, b, x0_eff);
      if (qp.bounds_invalid()) out.status = LPVMPC_DATA_ERROR;
      else {
        if (p.S.scaling) c = qp.scale(p.S.scaling);
        else {
          for (int j = lane; j < L.nz; j += 32) { w[L.D + j] = 1.0; w[L.Dinv + j] = 1.0; }
          for (int i = lane; i < L.m; i += 32) { w[L.E + i] = 1.0; w[L.Einv + i] = 1.0; }
          __syncwarp();
        }
        qp.classify();
        admm_run<KIND>(qp, p.S, c, out, p, b);
        solved_path = true;
      }
    }
    const bool has_sol = solved_path &&
                         !(out.status == LPVMPC_PRIMAL_INFEASIBLE || out.status == LPVMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                           out.status == LPVMPC_DUAL_INFEASIBLE || out.status == LPVMPC_DUAL_INFEASIBLE_INACCURATE ||
                           out.status == LPVMPC_NON_CVX);
    // unpack: x = D x_bar (PathFollowingLPVMPC.py:157-158 split)
    const double *x = w + L.x, *D = w + L.D;
    for (int j = lane; j < L.nz; j += 32) {
      const double v = has_sol ? D[j] * x[j] : nan("");
      if (j < L.nx) a.x_pred[(size_t)b * L.nx + j] = v;
      else a.u_pred[(size_t)b * (L.nz - L.nx) + (j - L.nx)] = v;
    }
    if (a.y || a.active_lo || a.active_up) {
      const double *y = w + L.y, *E = w + L.E;
      const int8_t *ty = qp.types();
      const double cinv = 1.0 / c;
      for (int i = lane; i < L.m; i += 32) {
        const size_t o = (size_t)b * L.m + Spec<KIND>::ref_row(L, i);
        if (a.y) a.y[o] = has_sol ? cinv * (E[i] * y[i]) : nan("");
        const int t = solved_path ? ty[i] : 0;
        if (a.active_lo) a.active_lo[o] = (t & 4) ? 1 : 0;
        if (a.active_up) a.active_up[o] = (t & 8) ? 1 : 0;
      }
    }
    if (lane == 0) {
      a.status[b] = out.status;
      if (a.iters) a.iters[b] = out.iter;
      if (a.rho_updates) a.rho_updates[b] = out.rho_updates;
      if (a.polish_status) a.polish_status[b] = out.polish_status;
      if (a.obj) a.obj[b] = out.obj;
      if (a.pri_res) a.pri_res[b] = out.pri_res;
      if (a.dua_res) a.dua_res[b] = out.dua_res;
    }
    __syncwarp();
    (void)NX;
  }
}

// Stand-alone LPVPrediction / _EstimateABC: one thread per QP (the roll-out is serial in k), HBM-bound (3.8 KB per
// controller QP at N = 8, 91 % of it the matrices written).  A thread's own records are 288 / 96 / 48 bytes per stage, 2.3 KB
// apart from its neighbour's, so the stage results of a warp's 32 QPs go through a shared-memory tile and leave as runs
// of consecutive 16-byte pieces, 512 bytes per store instruction.  Tile row = one QP's record [A | B | state]; its stride
// in 16-byte units is odd (27 controller, 21 planner), which makes the 128-bit row writes of a quarter-warp conflict-free.
constexpr int kSchedThreads = 64;
template <int KIND> struct SchedTile {
  static constexpr int NX = Spec<KIND>::NX, NA = NX * NX, NBm = NX * NU, REC = NA + NBm + NX;
  static constexpr int STRIDE = (((REC + 1) / 2) | 1) * 2;   // doubles: 54 (controller), 42 (planner)
  static constexpr bool PAIRS = (NA % 2 == 0) && (NX % 2 == 0);   // every record a whole number of 16-byte pieces
};
// copy records of `len` doubles (QP q of the warp at tile[q * STRIDE + off]) to dst[(b0 + q) * qstride + e]
template <int STRIDE, int LEN, bool PAIRS>
__device__ __forceinline__ void sched_copy_out(const double *tile, int off, double *dst, size_t qstride, int nq, int lane) {
  if (PAIRS && nq == 32) {
    // full warp: 32 records = LEN / 2 pieces per lane exactly; up to 9 tile reads in flight before their stores
    constexpr int L2 = LEN / 2, STEP_Q = 32 / L2, STEP_E = 32 % L2, CH = (L2 % 9 == 0) ? 9 : ((L2 % 6 == 0) ? 6 : ((L2 % 3 == 0) ? 3 : 1));
    int q = lane / L2, e = lane - q * L2;
#pragma unroll 1
    for (int c0 = 0; c0 < L2; c0 += CH) {
      double2 v[CH];
      int go[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        v[j] = *reinterpret_cast<const double2 *>(tile + q * STRIDE + off + 2 * e);
        go[j] = q * (int)qstride + 2 * e;
        q += STEP_Q; e += STEP_E;
        if (e >= L2) { e -= L2; ++q; }
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) *reinterpret_cast<double2 *>(dst + go[j]) = v[j];
    }
  } else if (PAIRS) {
    constexpr int L2 = LEN / 2, STEP_Q = 32 / L2, STEP_E = 32 % L2;
    int q = lane / L2, e = lane - q * L2;
    for (; q < nq; ) {
      const double2 v = *reinterpret_cast<const double2 *>(tile + q * STRIDE + off + 2 * e);
      *reinterpret_cast<double2 *>(dst + (size_t)q * qstride + 2 * e) = v;
      q += STEP_Q; e += STEP_E;
      if (e >= L2) { e -= L2; ++q; }
    }
  } else {
    constexpr int STEP_Q = 32 / LEN, STEP_E = 32 % LEN;
    int q = lane / LEN, e = lane - q * LEN;
    for (; q < nq; ) {
      dst[(size_t)q * qstride + e] = tile[q * STRIDE + off + e];
      q += STEP_Q; e += STEP_E;
      if (e >= LEN) { e -= LEN; ++q; }
    }
  }
}
// 256-bit global store (sm_100: STG.E.256): one full 32-byte sector per lane and instruction
__device__ __forceinline__ void stg256(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// DIRECT (controller, 32-byte aligned outputs): every thread stores its own records straight from registers with 256-bit
// stores (A_k: 9, B_k: 3 per stage; the 48-byte state with three 128-bit stores) -- whole sectors, no shared-memory
// staging, 15 memory instructions per stage instead of 27 tile writes + 27 tile reads + 27 stores.
template <int KIND, bool DIRECT = false>
__global__ void __launch_bounds__(kSchedThreads, 8) lpv_schedule_kernel(const __grid_constant__ Params p, int *sched_err) {
  using T = SchedTile<KIND>;
  constexpr int NX = T::NX, NA = T::NA, NBm = T::NBm, STRIDE = T::STRIDE;
  static_assert(!DIRECT || (NA % 4 == 0 && NBm % 4 == 0 && NX % 2 == 0), "direct stores need records of whole 32-byte pieces");
  __shared__ __align__(16) double tile_s[DIRECT ? 1 : kSchedThreads / 32][DIRECT ? 2 : 32 * STRIDE];   // 27 KB: 8 CTAs = 16 warps per SM
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b0 = blockIdx.x * blockDim.x + warp * 32;      // first QP of my warp
  if (b0 >= p.B) return;
  const int nq = (p.B - b0 < 32) ? p.B - b0 : 32;          // QPs of my warp
  const bool live = lane < nq;
  const int b = live ? b0 + lane : b0;                      // idle lanes mirror the warp's first QP (no output of their own)
  double *tile = tile_s[DIRECT ? 0 : warp], *mine = tile + (DIRECT ? 0 : lane * STRIDE);
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const int N = p.L.N;
  int err = 0;
  double st[NX], Ai[NA], Bi[NBm];
  const bool estimate = a.sched_mode == LPVMPC_SCHED_ESTIMATE;
  const double *up = a.u_prev + (size_t)b * N * NU;
  int lap = 0;
  if (!estimate) {
    const double *xs = a.x_sched ? a.x_sched + (size_t)b * NX : a.x0 + (size_t)b * NX;
#pragma unroll
    for (int r = 0; r < NX; ++r) st[r] = xs[r];
    lap = a.lap ? a.lap[b] : a.lap_all;
  }
  // per-stage inputs one stage ahead of their use (the first touch of every record is a DRAM round trip)
  auto in0 = [&](int i) { return estimate ? 0.0 : (KIND == LPVMPC_CONTROLLER ? a.vel_ref[(size_t)b * (N + 1) + i] : a.SS[(size_t)b * (N + 1) + i]); };
  auto in1 = [&](int i) { return (!estimate && KIND == LPVMPC_CONTROLLER && lap != 0) ? a.curv_ref[(size_t)b * N + i] : 0.0; };
  double n_u0 = up[0], n_u1 = up[1], n_a = in0(0), n_c = in1(0);
  for (int i = 0; i < N; ++i) {
    const double u01[2] = {n_u0, n_u1}, va = n_a, vc = n_c;
    if (i + 1 < N) { n_u0 = up[(i + 1) * NU]; n_u1 = up[(i + 1) * NU + 1]; n_a = in0(i + 1); n_c = in1(i + 1); }
    if (estimate) {
      const double *t = a.traj + ((size_t)b * N + i) * 6;
      if (KIND == LPVMPC_CONTROLLER) ctrl_stage(M, M.Cf, M.Cr, t[0], t[1], t[3], t[5], curvature(M.track, M.nseg, t[4], err), u01[0], Ai, Bi);
      else plan_stage(M, t[0], t[1], t[3], t[4], curvature(M.track, M.nseg, t[5], err), u01[0], Ai, Bi);
    } else if (KIND == LPVMPC_CONTROLLER) {
      const double cur = (lap == 0) ? curvature(M.track, M.nseg, st[4], err) : vc;
      ctrl_stage(M, a.Cf_new, a.Cf_new, va, st[1], st[3], st[5], cur, u01[0], Ai, Bi);
    } else {
      plan_stage(M, st[0], st[1], st[3], st[4], curvature(M.track, M.nseg, va, err), u01[0], Ai, Bi);
    }
    if (!estimate) propagate<NX>(Ai, Bi, u01, st);
    if (DIRECT) {
      if (live) {
        if (a.A_out) {
          double *d = a.A_out + ((size_t)b * N + i) * NA;
#pragma unroll
          for (int e = 0; e < NA; e += 4) stg256(d + e, Ai[e], Ai[e + 1], Ai[e + 2], Ai[e + 3]);
        }
        if (a.B_out) {
          double *d = a.B_out + ((size_t)b * N + i) * NBm;
#pragma unroll
          for (int e = 0; e < NBm; e += 4) stg256(d + e, Bi[e], Bi[e + 1], Bi[e + 2], Bi[e + 3]);
        }
        if (a.states_out && !estimate) {
          double *d = a.states_out + ((size_t)b * N + i) * NX;
#pragma unroll
          for (int r = 0; r < NX; r += 2) *reinterpret_cast<double2 *>(d + r) = make_double2(st[r], st[r + 1 < NX ? r + 1 : r]);
        }
      }
      continue;
    }
    if (T::PAIRS) {
#pragma unroll
      for (int e = 0; e < NA; e += 2) *reinterpret_cast<double2 *>(mine + e) = make_double2(Ai[e], Ai[e + 1]);
#pragma unroll
      for (int e = 0; e < NBm; e += 2) *reinterpret_cast<double2 *>(mine + NA + e) = make_double2(Bi[e], Bi[e + 1]);
      if (!estimate) {
#pragma unroll
        for (int r = 0; r < NX; r += 2) *reinterpret_cast<double2 *>(mine + NA + NBm + r) = make_double2(st[r], st[r + 1 < NX ? r + 1 : r]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < NA; ++e) mine[e] = Ai[e];
#pragma unroll
      for (int e = 0; e < NBm; ++e) mine[NA + e] = Bi[e];
      if (!estimate) {
#pragma unroll
        for (int r = 0; r < NX; ++r) mine[NA + NBm + r] = st[r];
      }
    }
    __syncwarp();
    if (a.A_out) sched_copy_out<STRIDE, NA, T::PAIRS>(tile, 0, a.A_out + ((size_t)b0 * N + i) * NA, (size_t)N * NA, nq, lane);
    if (a.B_out) sched_copy_out<STRIDE, NBm, T::PAIRS>(tile, NA, a.B_out + ((size_t)b0 * N + i) * NBm, (size_t)N * NBm, nq, lane);
    if (a.states_out && !estimate) sched_copy_out<STRIDE, NX, T::PAIRS>(tile, NA + NBm, a.states_out + ((size_t)b0 * N + i) * NX, (size_t)N * NX, nq, lane);
    __syncwarp();
  }
  if (sched_err && live) sched_err[b] = err;
}

// TMA variant (controller; A_out / states_out 16-byte, B_out 32-byte aligned).  What bounds this kernel is the WRITE PATTERN
// (tools/write_pattern3.cu, profiles/r3_write_pattern.jsonl): with one thread per QP the three arrays are written stage by
// stage, 288 / 96 / 48 bytes per QP at strides of 2,304 / 768 / 384 bytes, and the 48-byte state records in particular
// (one and a half sectors) halve what the memory system sustains (3.45 TB/s for the pattern against 5.0 TB/s when B_k and
// the states of FOUR consecutive stages leave together; A_k is indifferent to it).  So:
//   A_k     one dense shared-memory box [32][36] per stage -> cp.async.bulk.tensor store, tensor {36, N, B}, box {36, 1, 32}
//   states  shared-memory box [32][4][6], stored once per four stages, tensor {6, N, B}, box {6, 4, 32} (192 bytes per QP)
//   B_k     its three varying entries stay in registers for four stages, then 12 256-bit stores per lane (384 bytes per QP;
//           a third TMA box would not fit next to the other two at 14 resident warps per SM, and writing it over the A box
//           needs an exposed wait for the A store every second stage: measured 64.5 against 62.5 us)
// The TMA unit walks the strides between neighbouring QPs and clips partial warps and partial chunks; one lane issues.
struct SchedMaps { CUtensorMap A, S; };
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}
constexpr int kTmaWarps = 2;   // warps per CTA
constexpr int kTmaChunk = 4;   // stages of B_k / states written together
__global__ void __launch_bounds__(32 * kTmaWarps, 7) lpv_schedule_tma_kernel(const __grid_constant__ Params p, int *sched_err,
                                                                           const __grid_constant__ SchedMaps maps) {
  constexpr int NX = 6, NA = 36, NBm = 12, CH = kTmaChunk;
  __shared__ __align__(128) double tile_s[kTmaWarps][32 * (NA + CH * NX)];   // per warp: A box 9,216 B | state box 6,144 B
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b0 = blockIdx.x * blockDim.x + warp * 32;      // first QP of my warp
  if (b0 >= p.B) return;
  const int nq = (p.B - b0 < 32) ? p.B - b0 : 32;
  const bool live = lane < nq;
  const int b = live ? b0 + lane : b0;                      // idle lanes mirror the warp's first QP (their rows are clipped)
  double *tA = tile_s[warp], *tS = tA + 32 * NA;
  double *mA = tA + lane * NA, *mS = tS + lane * (CH * NX);
  const uint32_t sA = (uint32_t)__cvta_generic_to_shared(tA), sS = (uint32_t)__cvta_generic_to_shared(tS);
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const int N = p.L.N;
  int err = 0;
  double st[NX], Ai[NA], Bi[NBm];
  const bool estimate = a.sched_mode == LPVMPC_SCHED_ESTIMATE;
  const double *up = a.u_prev + (size_t)b * N * NU;
  int lap = 0;
  if (!estimate) {
    const double *xs = a.x_sched ? a.x_sched + (size_t)b * NX : a.x0 + (size_t)b * NX;
#pragma unroll
    for (int r = 0; r < NX; ++r) st[r] = xs[r];
    lap = a.lap ? a.lap[b] : a.lap_all;
  }
  // curv_ref is read whenever it is given (it is only USED when lap != 0): the load must not wait for `lap` to arrive
  const bool have_cr = !estimate && a.curv_ref != nullptr;
  auto in0 = [&](int i) { return estimate ? 0.0 : a.vel_ref[(size_t)b * (N + 1) + i]; };
  auto in1 = [&](int i) { return have_cr ? a.curv_ref[(size_t)b * N + i] : 0.0; };
  // per-stage inputs two stages ahead of their use
  double n_u0 = up[0], n_u1 = up[1], n_a = in0(0), n_c = in1(0);
  double m_u0 = 0.0, m_u1 = 0.0, m_a = 0.0, m_c = 0.0;
  if (N > 1) { m_u0 = up[NU]; m_u1 = up[NU + 1]; m_a = in0(1); m_c = in1(1); }
  const bool wantS = a.states_out && !estimate;
#pragma unroll 1
  for (int i0 = 0; i0 < N; i0 += CH) {
    double bq[CH][3];
#pragma unroll
    for (int s = 0; s < CH; ++s) {
      const int i = i0 + s;
      bq[s][0] = bq[s][1] = bq[s][2] = 0.0;
      if (i < N) {
        const double u01[2] = {n_u0, n_u1}, va = n_a, vc = n_c;
        n_u0 = m_u0; n_u1 = m_u1; n_a = m_a; n_c = m_c;
        if (i + 2 < N) { m_u0 = up[(i + 2) * NU]; m_u1 = up[(i + 2) * NU + 1]; m_a = in0(i + 2); m_c = in1(i + 2); }
        if (estimate) {
          const double *t = a.traj + ((size_t)b * N + i) * 6;
          ctrl_stage(M, M.Cf, M.Cr, t[0], t[1], t[3], t[5], curvature(M.track, M.nseg, t[4], err), u01[0], Ai, Bi);
        } else {
          const double cur = (lap == 0) ? curvature(M.track, M.nseg, st[4], err) : vc;
          ctrl_stage(M, a.Cf_new, a.Cf_new, va, st[1], st[3], st[5], cur, u01[0], Ai, Bi);
          propagate<NX>(Ai, Bi, u01, st);
        }
        bq[s][0] = Bi[0]; bq[s][1] = Bi[2]; bq[s][2] = Bi[4];
        // the stores issued so far have read their boxes (the last one a whole stage of arithmetic ago)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int e = 0; e < NA; e += 2) *reinterpret_cast<double2 *>(mA + e) = make_double2(Ai[e], Ai[e + 1]);
        if (!estimate) {
#pragma unroll
          for (int r = 0; r < NX; r += 2) *reinterpret_cast<double2 *>(mS + s * NX + r) = make_double2(st[r], st[r + 1]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic-proxy writes, before the async-proxy reads
        __syncwarp();
        if (lane == 0) {
          if (a.A_out) tma_store_3d(&maps.A, sA, 0, i, b0);
          if (wantS && (s == CH - 1 || i == N - 1)) tma_store_3d(&maps.S, sS, 0, i0, b0);   // rows of stages >= N are clipped
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (a.B_out && live) {   // B_k = dt [[B11, 1], [B21, 0], [B31, 0], 0 ...]: the chunk's records, 96 bytes each
      double *d = a.B_out + ((size_t)b * N + i0) * NBm;
      const double one = Bi[1], zero = Bi[3];   // dt * 1.0, dt * 0.0 as ctrl_stage forms them
#pragma unroll
      for (int s = 0; s < CH; ++s) {
        if (i0 + s < N) {
          stg256(d + s * NBm, bq[s][0], one, bq[s][1], zero);
          stg256(d + s * NBm + 4, bq[s][2], zero, zero, zero);
          stg256(d + s * NBm + 8, zero, zero, zero, zero);
        }
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the reads
  __syncwarp();
  if (sched_err && live) sched_err[b] = err;
}

// The first version of the stand-alone kernel (every thread stores its own records): kept for the A/B line of
// bench.py --workload sched65536 (LPVMPC_SCHED_NAIVE=1), not used otherwise.
template <int KIND>
__global__ void __launch_bounds__(128) lpv_schedule_naive_kernel(const __grid_constant__ Params p, int *sched_err) {
  constexpr int NX = Spec<KIND>::NX;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const Model &M = p.M;
  const lpvmpc_args &a = p.a;
  const int N = p.L.N;
  int err = 0;
  double st[NX], Ai[NX * NX], Bi[NX * NU];
  if (a.sched_mode == LPVMPC_SCHED_ESTIMATE) {
    for (int i = 0; i < N; ++i) {
      const double *t = a.traj + ((size_t)b * N + i) * 6;
      const double delta = a.u_prev[((size_t)b * N + i) * NU];
      if (KIND == LPVMPC_CONTROLLER) ctrl_stage(M, M.Cf, M.Cr, t[0], t[1], t[3], t[5], curvature(M.track, M.nseg, t[4], err), delta, Ai, Bi);
      else plan_stage(M, t[0], t[1], t[3], t[4], curvature(M.track, M.nseg, t[5], err), delta, Ai, Bi);
      if (a.A_out) for (int e = 0; e < NX * NX; ++e) a.A_out[((size_t)b * N + i) * NX * NX + e] = Ai[e];
      if (a.B_out) for (int e = 0; e < NX * NU; ++e) a.B_out[((size_t)b * N + i) * NX * NU + e] = Bi[e];
    }
  } else {
    const double *xs = a.x_sched ? a.x_sched + (size_t)b * NX : a.x0 + (size_t)b * NX;
    for (int r = 0; r < NX; ++r) st[r] = xs[r];
    const double *up = a.u_prev + (size_t)b * N * NU;
    const int lap = a.lap ? a.lap[b] : a.lap_all;
    for (int i = 0; i < N; ++i) {
      if (KIND == LPVMPC_CONTROLLER) {
        const double cur = (lap == 0) ? curvature(M.track, M.nseg, st[4], err) : a.curv_ref[(size_t)b * N + i];
        ctrl_stage(M, a.Cf_new, a.Cf_new, a.vel_ref[(size_t)b * (N + 1) + i], st[1], st[3], st[5], cur, up[i * NU], Ai, Bi);
      } else {
        plan_stage(M, st[0], st[1], st[3], st[4], curvature(M.track, M.nseg, a.SS[(size_t)b * (N + 1) + i], err), up[i * NU], Ai, Bi);
      }
      if (a.A_out) for (int e = 0; e < NX * NX; ++e) a.A_out[((size_t)b * N + i) * NX * NX + e] = Ai[e];
      if (a.B_out) for (int e = 0; e < NX * NU; ++e) a.B_out[((size_t)b * N + i) * NX * NU + e] = Bi[e];
      propagate<NX>(Ai, Bi, up + i * NU, st);
      if (a.states_out) for (int r = 0; r < NX; ++r) a.states_out[((size_t)b * N + i) * NX + r] = st[r];
    }
  }
  if (sched_err) sched_err[b] = err;
}

// Visiting order of a controller batch: counting sort of the problems by k = vel_ref[0] - vx0 (descending, 256 buckets
// over [-3, 3] m/s).  The ADMM iteration count of the controller QP is a steep function of that velocity error (it
// decides which input bounds are active): batches in their natural order lose 38 % of the hot-loop work to warps
// waiting for their slowest QP (4 QPs advance in lockstep), grouped by k they lose 9 % (profiles/r2g_*).  One CTA.
// With `hint` (expected iterations per problem, e.g. the counts of the previous tick of a closed loop: they predict the
// next tick's almost perfectly) the buckets are hint / 25, longest first.
__global__ void __launch_bounds__(1024) lpv_order_kernel(const double *__restrict__ x0, const double *__restrict__ vel_ref, int nx,
                                                         int nv, int B, const int *__restrict__ hint, int *__restrict__ perm) {
  __shared__ int hist[256];
  __shared__ int offs[256];
  const int t = threadIdx.x;
  if (t < 256) hist[t] = 0;
  __syncthreads();
  auto bucket = [&](int b) {
    if (hint) { int q = hint[b] / 25; q = q < 0 ? 0 : (q > 255 ? 255 : q); return 255 - q; }
    const double k = vel_ref[(size_t)b * nv] - x0[(size_t)b * nx];
    double f = (3.0 - k) * (256.0 / 6.0);
    if (!(f == f)) f = 255.0;
    f = f < 0.0 ? 0.0 : (f > 255.0 ? 255.0 : f);
    return (int)f;
  };
  // the first few elements of a thread keep their bucket in registers (a 4,096-QP batch: all of them)
  constexpr int KEEP = 4;
  int mine[KEEP];
#pragma unroll
  for (int j = 0; j < KEEP; ++j) {
    const int b = t + j * (int)blockDim.x;
    mine[j] = (b < B) ? bucket(b) : -1;
    if (mine[j] >= 0) atomicAdd(&hist[mine[j]], 1);
  }
  for (int b = t + KEEP * (int)blockDim.x; b < B; b += blockDim.x) atomicAdd(&hist[bucket(b)], 1);
  __syncthreads();
  // exclusive prefix sum of the 256 bucket counts: warp scans + the 8 warp totals
  __shared__ int wtot[8];
  int incl = 0, cnt = 0;
  if (t < 256) {
    cnt = hist[t]; incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += v; }
    if ((t & 31) == 31) wtot[t >> 5] = incl;
  }
  __syncthreads();
  if (t < 256) {
    int base = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) base += (w < (t >> 5)) ? wtot[w] : 0;
    offs[t] = base + incl - cnt;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < KEEP; ++j) if (mine[j] >= 0) perm[atomicAdd(&offs[mine[j]], 1)] = t + j * (int)blockDim.x;
  for (int b = t + KEEP * (int)blockDim.x; b < B; b += blockDim.x) perm[atomicAdd(&offs[bucket(b)], 1)] = b;
}

}  // namespace lpv

// ================================================================================================
// Host side: handle, layout, staging, C-ABI
// ================================================================================================
using lpv::Layout;
using lpv::Params;

namespace {

thread_local std::string g_create_error;

struct Field {  // one staged array of lpvmpc_args
  size_t off_args;   // offsetof the pointer inside lpvmpc_args
  size_t elem;       // bytes per problem
  bool output;
};

inline int align2(int v) { return (v + 1) & ~1; }

Layout make_layout(int kind, int N, int delay) {
  const int NX = kind == LPVMPC_CONTROLLER ? 6 : 5, NUc = 2, NB = NX + NUc;
  Layout L;
  std::memset(&L, 0, sizeof(L));
  L.N = N; L.nx = NX * (N + 1); L.nz = L.nx + NUc * N; L.md = L.nx;
  L.ms = kind == LPVMPC_CONTROLLER ? 6 * N + delay : L.nz;
  L.m = L.md + L.ms;
  int o = 0;
  auto take = [&](int n) { const int r = o; o += align2(n); return r; };
  L.G = take(N * NX * NB); L.gI = take(L.md); L.sc = take(L.ms);
  L.Pxx = take((N + 1) * NX * NX); L.Puu = take(N * NUc * NUc); L.Pud = take((N > 1 ? N - 1 : 0) * NUc);
  L.q = take(L.nz); L.D = take(L.nz); L.Dinv = take(L.nz);
  L.E = take(L.m); L.Einv = take(L.m); L.l = take(L.m); L.u = take(L.m); L.z = take(L.m); L.y = take(L.m);
  L.x = take(L.nz); L.xp = take(L.nz); L.xt = take(L.nz); L.tn = take(L.nz); L.pdx = take(L.nz);
  L.dy = take(L.m); L.tm = take(L.m);
  L.K = take((N + 1) * NB * NB); L.T = take((N + 1) * NB * NB); L.So = take(NB * NB);
  L.type = take((L.m + 7) / 8);
  L.total = o;
  return L;
}

#ifdef LPVMPC_LEGACY
lpv::g8::Lay make_g8_layout(int kind, int N) {
  const int NX = kind == LPVMPC_CONTROLLER ? 6 : 5;
  lpv::g8::Lay L;
  std::memset(&L, 0, sizeof(L));
  L.N = N; L.gs = NX * 8;
  int o = 0;
  auto take = [&](int n) { const int r = o; o += n; return r; };
  L.T = take((N + 1) * 64); L.K = take(N * 64); L.G = take(N * L.gs);
  const int v = (N + 1) * 8;
  L.X = take(v); L.Q = take(v); L.B = take(v); L.YD = take(v); L.BE = take(v); L.ED = take(v);
  L.ZI = take(v); L.YI = take(v); L.SI = take(v); L.UI = take(v); L.LI = take(v);
  L.total = o;
  L.cold_total = lpv::g8::C_COUNT * v;
  return L;
}
#endif

// ring = 0: block factor resident in shared memory; ring = 4 (lpv::h8::kRing): factor in the slab, staged through a ring
// of stage blocks by TMA bulk copies issued lpv::h8::kAhead stage steps ahead
lpv::h8::Lay make_h8_layout(int kind, int N, int ring) {
  const int NX = kind == LPVMPC_CONTROLLER ? 6 : 5;
  lpv::h8::Lay L;
  std::memset(&L, 0, sizeof(L));
  L.N = N; L.nsl = kind == LPVMPC_CONTROLLER ? 6 : 7;
  L.is = kind == LPVMPC_CONTROLLER ? 26 : 38;
  L.ring = ring; L.pfd = ring ? ring - 2 : 0;
  int o = 0;
  auto take = [&](int n) { const int r = o; o += (n + 1) & ~1; return r; };
  L.TK = take(ring ? ring * lpv::h8::TKS : (N + 1) * lpv::h8::TKS);   // slot N's multiplier half is used by the twisted kernels
  L.V = take((N + 1) * lpv::h8::VS);
  L.I = take((N + 2) * L.is);
  while (o % 16 != 8) o += 2;  // neighbouring groups of a warp 64 B apart mod 128
  L.total = o;
  const int v = (N + 1) * 8;
  L.cG = lpv::h8::C_COUNT * v;
  L.cold_total = L.cG + N * NX * 8;
  L.cTK = L.cold_total;
  if (ring) L.cold_total += (N + 1) * lpv::h8::TKS;   // also setup's scratch (8 (N+1) 8 + N NX 8 + nz doubles fit)
  return L;
}

lpv::h8t::Lay make_h8t_layout(int kind, int N) {
  const int NX = kind == LPVMPC_CONTROLLER ? 6 : 5;
  lpv::h8t::Lay L;
  std::memset(&L, 0, sizeof(L));
  L.N = N; L.nsl = kind == LPVMPC_CONTROLLER ? 6 : 7;
  L.is = kind == LPVMPC_CONTROLLER ? 26 : 38;
  int o = 0;
  auto take = [&](int n) { const int r = o; o += (n + 1) & ~1; return r; };
  L.V = take((N + 1) * lpv::h8t::VS);
  L.I = take((N + 2) * L.is);
  L.G = take(N * NX * 8);
  L.CS = take(lpv::h8t::C_NSMEM * (N + 1) * 8);
  L.FS = take(128);
  while (o % 16 != 8) o += 2;  // neighbouring groups of a warp 64 B apart mod 128
  L.total = o;
  L.cold_total = (lpv::h8t::C_COUNT - lpv::h8t::C_NSMEM) * (N + 1) * 8;
  return L;
}

// H16T: the H8T layout without the factorisation scratch (it lives in dead stage-vector slots)
lpv::h8t::Lay make_h16t_layout(int kind, int N) {
  lpv::h8t::Lay L = make_h8t_layout(kind, N);
  int o = L.FS;              // FS was the last block
  while (o % 16 != 8) o += 2;
  L.FS = 0; L.total = o;
  return L;
}

}  // namespace

struct lpvmpc_handle {
  lpvmpc_cfg cfg;
  Layout L;
  lpv::Model M;
  int n, d;
  int device;
  int sm_count;
  int smem_optin;
  bool smem_mode;
  int grid_cap;           // resident CTAs (persistent grid)
  size_t ws_bytes;
  double *d_track = nullptr;
  double *d_gws = nullptr;
  int variant = 1;           // 1: generic warp-per-QP kernel, 2: T8 register/shared-resident kernel, 3: G8 compact kernel
#ifdef LPVMPC_LEGACY
  lpv::g8::Lay GL;           // G8 shared-memory layout
#endif
  lpv::h8::Lay HL;           // H8 layout (shared memory + slab)
  lpv::h8t::Lay TL;          // H8T layout (tensor memory + shared memory + slab)
  lpv::h8t::Lay TL16;        // H16T layout
  int wpc = 1;               // H8: warps per CTA
  bool twisted = false;      // H8, one QP per warp: twisted factorisation kernel
  int helpers = 0;           // ... with this many helper warps per QP for the element-wise updates
  int qpw = 4;               // T8 / G8 / H8: QPs per warp
  unsigned *d_queue = nullptr;  // work-queue counter of the persistent T8 kernel
  int *d_perm = nullptr;        // visiting order of the batch (lpv_order_kernel)
  double *d_cold = nullptr;     // T8 scratch slab (scalings, P) per resident lane
  // staging for the host API
  char *d_stage = nullptr, *h_stage = nullptr;
  size_t stage_bytes = 0;
  char *h_ring = nullptr;      // two pinned result arenas of lpvmpc_solve_host_view (allocated on its first call)
  size_t ring_bytes = 0;
  int ring_slot = 0;
  bool zc_in = false;          // host API, solve calls: the kernel reads its inputs from the pinned arena instead of one H2D in front of it (default for the
                               // H16T kernel: ctrl4096 end to end 1.117 -> 1.081 ms; LPVMPC_ZERO_COPY_IN=0 / 1 overrides.  The one-QP-per-warp kernels walk a serial
                               // roll-out over their inputs -- a PCIe round trip per stage -- and keep the copy)
  bool zc_out = true;          // host API: the kernel writes its results straight into the pinned arena (LPVMPC_ZERO_COPY_OUT=0: D2H copy)
  cudaStream_t stream = nullptr;
  // The _dev entry points launch on the caller's stream but share per-handle device scratch (work queue, visiting order,
  // slab, loop buffers): a call arriving on another stream than the previous one first waits for that one's work.
  cudaEvent_t last_ev = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_valid = false;
  long long launches = 0;
  std::string err;
  // closed-loop fleet (lpvmpc_loop_*)
  char *d_loop = nullptr;        // one allocation holding every array below
  lpv::loop::LoopParams LP;
  lpvmpc_args loop_args;         // device pointers of the per-tick solve
  double *d_loop_xpred = nullptr, *d_loop_velref = nullptr;
  int32_t *d_loop_status = nullptr, *d_loop_iters = nullptr;
  int loop_B = 0;
  long long loop_tick = 0;       // ticks since the last init (selects the warm-up path)
  cudaEvent_t loop_ev = nullptr; // last lpvmpc_loop_run_dev on the caller's stream (lpvmpc_loop_read_host waits for it)
  // planner loop (lpvmpc_plan_loop_*); shares d_loop_status / d_loop_iters / loop_B / loop_tick / loop_ev with the above
  char *d_ploop = nullptr;
  lpv::loop::PlanLoopParams PP;
  double *d_ploop_maxey = nullptr;
  // planner -> controller references (lpvmpc_plan_refs_*)
  double *d_refs_W = nullptr;    // W then Wc, each [n_out, N]
  int refs_n_out = 0;
};

namespace {

int fail(lpvmpc_handle *h, int code, const std::string &msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
#define CUDA_TRY(h, expr)                                                                     \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(h, LPVMPC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));     \
  } while (0)

// Restores the calling thread's current device on every return path (a process that drives several GPUs keeps its own
// current device; torch's allocator and default stream follow it).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define ON_DEVICE(h)                                                                           \
  DeviceGuard dev_guard__((h)->device);                                                        \
  if (dev_guard__.err != cudaSuccess)                                                          \
    return fail(h, LPVMPC_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(dev_guard__.err))

// Stream hand-over of the per-handle scratch (see lpvmpc_handle::last_ev).
int stream_enter(lpvmpc_handle *h, cudaStream_t s) {
  if (h->last_valid && h->last_stream != s) CUDA_TRY(h, cudaStreamWaitEvent(s, h->last_ev, 0));
  return LPVMPC_OK;
}
int stream_leave(lpvmpc_handle *h, cudaStream_t s) {
  CUDA_TRY(h, cudaEventRecord(h->last_ev, s));
  h->last_stream = s; h->last_valid = true;
  return LPVMPC_OK;
}

const char *settings_error(const lpvmpc_settings *s) {
  if (s->max_iter < 1) return "settings: max_iter must be >= 1";
  if (!(s->rho > 0) || !(s->sigma > 0) || !(s->delta > 0)) return "settings: rho, sigma and delta must be > 0";
  if (!(s->alpha > 0) || !(s->alpha < 2)) return "settings: alpha must be in (0, 2)";
  if (!(s->eps_abs >= 0) || !(s->eps_rel >= 0) || !(s->eps_prim_inf > 0) || !(s->eps_dual_inf > 0)) return "settings: negative / NaN tolerance";
  if (s->check_termination < 0 || s->scaling < 0 || s->polish_refine_iter < 0 || s->adaptive_rho_interval < 0)
    return "settings: check_termination, scaling, polish_refine_iter and adaptive_rho_interval must be >= 0";
  if (s->adaptive_rho && !(s->adaptive_rho_tolerance >= 1)) return "settings: adaptive_rho_tolerance must be >= 1";
  return nullptr;
}

Params make_params(const lpvmpc_handle *h, int B, const lpvmpc_args *a) {
  Params p;
  p.L = h->L; p.M = h->M; p.S = h->cfg.settings; p.a = *a; p.B = B; p.gws = h->d_gws;
  return p;
}

int validate_args(lpvmpc_handle *h, int B, const lpvmpc_args *a, bool solve) {
  if (!h || !a) return fail(h, LPVMPC_E_ARG, "null handle/args");
  if (B < 0 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "batch exceeds max_batch");
  const bool ctrl = h->cfg.kind == LPVMPC_CONTROLLER;
  if (a->sched_mode < 0 || a->sched_mode > 2) return fail(h, LPVMPC_E_ARG, "bad sched_mode");
  if (solve) {
    if (!a->x0 || !a->x_pred || !a->u_pred || !a->status) return fail(h, LPVMPC_E_ARG, "x0, x_pred, u_pred, status are required");
    if (ctrl && !a->vel_ref) return fail(h, LPVMPC_E_ARG, "controller needs vel_ref");
    if (!ctrl && !a->max_ey) return fail(h, LPVMPC_E_ARG, "planner needs max_ey");
    if (ctrl && h->cfg.steering_delay > 0 && !a->old_steering) return fail(h, LPVMPC_E_ARG, "steering_delay needs old_steering");
    if (a->sched_mode == LPVMPC_SCHED_GIVEN && (!a->A || !a->Bm)) return fail(h, LPVMPC_E_ARG, "SCHED_GIVEN needs A and Bm");
  }
  if (a->sched_mode == LPVMPC_SCHED_PREDICT) {
    if (!a->u_prev || (!a->x0 && !a->x_sched)) return fail(h, LPVMPC_E_ARG, "PREDICT needs x0/x_sched and u_prev");
    if (ctrl && !a->vel_ref) return fail(h, LPVMPC_E_ARG, "controller PREDICT needs vel_ref");
    if (ctrl && !a->curv_ref && (a->lap || a->lap_all != 0)) return fail(h, LPVMPC_E_ARG, "controller PREDICT with lap != 0 needs curv_ref");
    if (!ctrl && !a->SS) return fail(h, LPVMPC_E_ARG, "planner PREDICT needs SS");
  }
  if (a->sched_mode == LPVMPC_SCHED_ESTIMATE && (!a->traj || !a->u_prev)) return fail(h, LPVMPC_E_ARG, "ESTIMATE needs traj and u_prev");
  return LPVMPC_OK;
}

#ifdef LPVMPC_LEGACY
constexpr int kT8N = 8;

int launch_t8(lpvmpc_handle *h, const Params &p, cudaStream_t s) {
  if (p.B == 0) return LPVMPC_OK;
  const int warps = (p.B + h->qpw - 1) / h->qpw;
  const int grid = warps < h->grid_cap ? warps : h->grid_cap;
  CUDA_TRY(h, cudaMemsetAsync(h->d_queue, 0, sizeof(unsigned), s));
  if (h->qpw == 2) lpv::t8::lpv_solve_t8_kernel<kT8N, 2><<<grid, 32, h->ws_bytes, s>>>(p, h->d_queue, h->d_cold);
  else lpv::t8::lpv_solve_t8_kernel<kT8N, 4><<<grid, 32, h->ws_bytes, s>>>(p, h->d_queue, h->d_cold);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}

template <int KIND, int QPW>
void g8_launch(int grid, size_t smem, cudaStream_t s, const lpv::g8::G8Params &gp) {
  lpv::g8::lpv_solve_g8_kernel<KIND, QPW><<<grid, 32, smem, s>>>(gp);
}
template <int KIND, int QPW>
cudaError_t g8_attr(size_t smem) {
  return cudaFuncSetAttribute(lpv::g8::lpv_solve_g8_kernel<KIND, QPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int KIND>
int launch_g8(lpvmpc_handle *h, const Params &p, cudaStream_t s) {
  if (p.B == 0) return LPVMPC_OK;
  lpv::g8::G8Params gp;
  gp.L = h->GL; gp.M = p.M; gp.S = p.S; gp.a = p.a; gp.B = p.B; gp.queue = h->d_queue; gp.cold = h->d_cold;
  const int warps = (p.B + h->qpw - 1) / h->qpw;
  const int grid = warps < h->grid_cap ? warps : h->grid_cap;
  CUDA_TRY(h, cudaMemsetAsync(h->d_queue, 0, sizeof(unsigned), s));
  if (h->qpw == 4) g8_launch<KIND, 4>(grid, h->ws_bytes, s, gp);
  else if (h->qpw == 2) g8_launch<KIND, 2>(grid, h->ws_bytes, s, gp);
  else g8_launch<KIND, 1>(grid, h->ws_bytes, s, gp);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}
#endif  // LPVMPC_LEGACY

// controller batches: the grouped visiting order (NULL when the caller wants the batch order or it does not apply)
const int *batch_order(lpvmpc_handle *h, const Params &p, cudaStream_t s) {
  // one QP per warp: nothing advances in lockstep; longest-first dispatch only on the caller's `order_hint` (the velocity-error
  // key does not predict the iteration count of the long-horizon batch: ctrl1024N100 49.1 ms ordered by it, 44.6 ms unordered)
  if (p.a.natural_order || p.B < 8 || (h->qpw < 2 && !p.a.order_hint)) return nullptr;
  if (!p.a.order_hint && (h->cfg.kind != LPVMPC_CONTROLLER || !p.a.vel_ref || !p.a.x0)) return nullptr;
  lpv::lpv_order_kernel<<<1, 1024, 0, s>>>(p.a.x0, p.a.vel_ref, h->n, p.L.N + 1, p.B, p.a.order_hint, h->d_perm);
  ++h->launches;
  return h->d_perm;
}

template <int KIND, int QPW>
void h8_launch(int grid, int threads, size_t smem, cudaStream_t s, const lpv::h8::H8Params &hp) {
  lpv::h8::lpv_solve_h8_kernel<KIND, QPW, false><<<grid, threads, smem, s>>>(hp);
}
template <int KIND, int QPW>
cudaError_t h8_attr(size_t smem) {
  return cudaFuncSetAttribute(lpv::h8::lpv_solve_h8_kernel<KIND, QPW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
// twisted factorisation: one QP per warp, resident factor, even N
template <int KIND>
void h8w_launch(int grid, int threads, size_t smem, cudaStream_t s, const lpv::h8::H8Params &hp) {
  lpv::h8::lpv_solve_h8_kernel<KIND, 1, false, true><<<grid, threads, smem, s>>>(hp);
}
template <int KIND>
cudaError_t h8w_attr(size_t smem) {
  return cudaFuncSetAttribute(lpv::h8::lpv_solve_h8_kernel<KIND, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
// twisted factorisation + NH helper warps: one QP per CTA of NH + 1 warps
template <int KIND, int NH>
void h8h_launch(int grid, size_t smem, cudaStream_t s, const lpv::h8::H8Params &hp) {
  lpv::h8::lpv_solve_h8_kernel<KIND, 1, false, true, NH><<<grid, 32 * (NH + 1), smem, s>>>(hp);
}
template <int KIND, int NH>
cudaError_t h8h_attr(size_t smem) {
  return cudaFuncSetAttribute(lpv::h8::lpv_solve_h8_kernel<KIND, 1, false, true, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
// streamed factor: one QP per warp, one warp per CTA
template <int KIND>
void h8s_launch(int grid, size_t smem, cudaStream_t s, const lpv::h8::H8Params &hp) {
  lpv::h8::lpv_solve_h8_kernel<KIND, 1, true><<<grid, 32, smem, s>>>(hp);
}
template <int KIND>
cudaError_t h8s_attr(size_t smem) {
  return cudaFuncSetAttribute(lpv::h8::lpv_solve_h8_kernel<KIND, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int KIND>
int launch_h8(lpvmpc_handle *h, const Params &p, cudaStream_t s) {
  if (p.B == 0) return LPVMPC_OK;
  lpv::h8::H8Params hp;
  hp.L = h->HL; hp.M = p.M; hp.S = p.S; hp.a = p.a; hp.B = p.B; hp.queue = h->d_queue; hp.cold = h->d_cold;
  hp.perm = batch_order(h, p, s);
  const int per_cta = h->qpw * h->wpc;
  const int ctas = (p.B + per_cta - 1) / per_cta;
  const int grid = ctas < h->grid_cap ? ctas : h->grid_cap;
  CUDA_TRY(h, cudaMemsetAsync(h->d_queue, 0, sizeof(unsigned), s));
  if (h->HL.ring) h8s_launch<KIND>(grid, h->ws_bytes, s, hp);
  else if (h->qpw == 4) h8_launch<KIND, 4>(grid, 32 * h->wpc, h->ws_bytes, s, hp);
  else if (h->qpw == 2) h8_launch<KIND, 2>(grid, 32 * h->wpc, h->ws_bytes, s, hp);
  else if (h->twisted && h->helpers) {
    if (KIND == LPVMPC_PLANNER) h8h_launch<LPVMPC_PLANNER, 1>(grid, h->ws_bytes, s, hp);
    else if (h->helpers == 7) h8h_launch<LPVMPC_CONTROLLER, 7>(grid, h->ws_bytes, s, hp);
    else h8h_launch<LPVMPC_CONTROLLER, 3>(grid, h->ws_bytes, s, hp);
  }
  else if (h->twisted) h8w_launch<KIND>(grid, 32 * h->wpc, h->ws_bytes, s, hp);
  else h8_launch<KIND, 1>(grid, 32 * h->wpc, h->ws_bytes, s, hp);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}

int launch_h8t(lpvmpc_handle *h, const Params &p, cudaStream_t s) {
  if (p.B == 0) return LPVMPC_OK;
  lpv::h8t::H8Params hp;
  hp.L = h->TL; hp.M = p.M; hp.S = p.S; hp.a = p.a; hp.B = p.B; hp.queue = h->d_queue; hp.cold = h->d_cold;
  hp.perm = batch_order(h, p, s);
  hp.cta_rounds = 0;
  const int ctas = (p.B + 15) / 16;
  const int grid = ctas < h->grid_cap ? ctas : h->grid_cap;
  CUDA_TRY(h, cudaMemsetAsync(h->d_queue, 0, sizeof(unsigned), s));
  lpv::h8t::lpv_solve_h8t_kernel<LPVMPC_CONTROLLER, 4><<<grid, 128, h->ws_bytes, s>>>(hp);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}

int launch_h16t(lpvmpc_handle *h, const Params &p, cudaStream_t s) {
  if (p.B == 0) return LPVMPC_OK;
  lpv::h8t::H8Params hp;
  hp.L = h->TL16; hp.M = p.M; hp.S = p.S; hp.a = p.a; hp.B = p.B; hp.queue = h->d_queue; hp.cold = h->d_cold;
  hp.perm = batch_order(h, p, s);
  static const int rounds = [] { const char *e = std::getenv("LPVMPC_H16T_ROUNDS"); return e ? std::atoi(e) : 2; }();   // 0: every warp pulls on its own, 1: rounds, 2: + phase barrier before the polish (sharing the last round out evenly over the CTAs was measured: no gain)
  hp.cta_rounds = rounds;
  const int ctas = (p.B + 15) / 16;
  const int grid = ctas < h->grid_cap ? ctas : h->grid_cap;
  CUDA_TRY(h, cudaMemsetAsync(h->d_queue, 0, sizeof(unsigned), s));
  lpv::h16t::lpv_solve_h16t_kernel<<<grid, 256, h->ws_bytes, s>>>(hp);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}

template <int KIND>
int launch_solve(lpvmpc_handle *h, const Params &p, cudaStream_t s) {
  if (h->variant == 8) return launch_h16t(h, p, s);
  if (h->variant == 6) return launch_h8t(h, p, s);
#ifdef LPVMPC_LEGACY
  if (h->variant == 2) return launch_t8(h, p, s);
  if (h->variant == 3) return launch_g8<KIND>(h, p, s);
#endif
  if (h->variant == 5) return launch_h8<KIND>(h, p, s);
  const int grid = p.B < h->grid_cap ? p.B : h->grid_cap;
  if (grid == 0) return LPVMPC_OK;
  if (h->smem_mode) lpv::lpv_solve_kernel<KIND, true><<<grid, 32, h->ws_bytes, s>>>(p);
  else lpv::lpv_solve_kernel<KIND, false><<<grid, 32, 0, s>>>(p);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}

// The staged fields of lpvmpc_args for the host API: {offset, bytes per problem, is_output}
std::vector<Field> staged_fields(const lpvmpc_handle *h) {
  const int n = h->n, d = h->d, N = h->L.N, nz = h->L.nz, m = h->L.m, delay = h->cfg.steering_delay;
  const size_t D = sizeof(double);
  std::vector<Field> f;
#define IN(name, cnt) f.push_back({offsetof(lpvmpc_args, name), (size_t)(cnt), false})
#define OUT(name, cnt) f.push_back({offsetof(lpvmpc_args, name), (size_t)(cnt), true})
  IN(x0, n * D); IN(x_sched, n * D); IN(A, N * n * n * D); IN(Bm, N * n * d * D); IN(C, N * n * D);
  IN(u_prev, N * d * D); IN(vel_ref, (N + 1) * D); IN(curv_ref, N * D); IN(SS, (N + 1) * D); IN(lap, sizeof(int32_t));
  IN(traj, N * 6 * D); IN(u_old, d * D); IN(old_steering, (delay > 0 ? delay : 1) * D); IN(max_ey, D);
  IN(ey_lo, (N + 1) * D); IN(ey_hi, (N + 1) * D); IN(order_hint, sizeof(int32_t));
  OUT(x_pred, (N + 1) * n * D); OUT(u_pred, N * d * D); OUT(status, sizeof(int32_t)); OUT(iters, sizeof(int32_t));
  OUT(rho_updates, sizeof(int32_t)); OUT(polish_status, sizeof(int32_t)); OUT(obj, D); OUT(pri_res, D); OUT(dua_res, D);
  OUT(active_lo, m); OUT(active_up, m); OUT(y, m * D); OUT(A_out, N * n * n * D); OUT(B_out, N * n * d * D);
  OUT(states_out, N * n * D); OUT(xs, nz * D); OUT(zs, m * D); OUT(ys, m * D);
#undef IN
#undef OUT
  return f;
}

inline void *&ptr_at(lpvmpc_args *a, size_t off) { return *reinterpret_cast<void **>(reinterpret_cast<char *>(a) + off); }
inline const void *cptr_at(const lpvmpc_args *a, size_t off) {
  return *reinterpret_cast<void *const *>(reinterpret_cast<const char *>(a) + off);
}
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Host-side staging copies (user arrays <-> the pinned arena) are most of the gap between the kernel and the end-to-end
// time of the _host entry points (3.8 MB per 4,096-QP controller call): spread them over a few OpenMP threads.
struct CopyJob { void *dst; const void *src; size_t bytes; };
#ifdef _OPENMP
// Threads of one process: at most 8, and the host's cores shared out over the processes of the node (one per GPU under
// torchrun: LOCAL_WORLD_SIZE) — 8 ranks x 8 spinning OpenMP threads on a 16-core host cost 22 % of the end-to-end rate
// (profiles/r2l_bench_ctrl4096_g8.json).  LPVMPC_COPY_THREADS overrides.
int copy_threads() {
  static const int n = [] {
    if (const char *e = std::getenv("LPVMPC_COPY_THREADS")) { const int v = std::atoi(e); if (v >= 1) return v > 64 ? 64 : v; }
    int procs = 1;
    if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) { const int v = std::atoi(e); if (v >= 1) procs = v; }
    // not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every rank, which made these copies serial
    // (2.4 M instead of 3.0 M QP/s per rank end to end, profiles/r2q_*); the num_threads clause overrides it
    const int t = omp_get_num_procs() / procs;
    return t < 1 ? 1 : (t > 8 ? 8 : t);
  }();
  return n;
}
#endif
void run_copies(const std::vector<CopyJob> &jobs) {
  constexpr size_t kChunk = 128 * 1024;
  std::vector<CopyJob> parts;
  size_t total = 0;
  for (const CopyJob &j : jobs) {
    total += j.bytes;
    for (size_t o = 0; o < j.bytes; o += kChunk)
      parts.push_back({static_cast<char *>(j.dst) + o, static_cast<const char *>(j.src) + o, (j.bytes - o < kChunk) ? j.bytes - o : kChunk});
  }
  if (total < 4 * kChunk) {
    for (const CopyJob &q : parts) std::memcpy(q.dst, q.src, q.bytes);
    return;
  }
  const int n = (int)parts.size();
#ifdef _OPENMP
  const int threads = copy_threads();
  if (threads > 1) {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < n; ++i) std::memcpy(parts[i].dst, parts[i].src, parts[i].bytes);
    return;
  }
#endif
  for (int i = 0; i < n; ++i) std::memcpy(parts[i].dst, parts[i].src, parts[i].bytes);
}

}  // namespace

extern "C" {

int lpvmpc_abi_version(void) { return LPVMPC_ABI_VERSION; }
#ifdef LPV_H8_PHASE_TIMING
// development builds only: cycles per phase added up by the helper-warp H8 kernels (lpv_h8.cuh), optionally cleared
int lpvmpc_debug_phase_cycles(unsigned long long *out16, int reset) {
  unsigned long long z[16] = {0};
  if (out16 && cudaMemcpyFromSymbol(out16, lpv::h8::g_phase_cycles, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(lpv::h8::g_phase_cycles, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
#endif

void lpvmpc_default_settings(lpvmpc_settings *s) {
  if (!s) return;
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->delta = 1e-6; s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000; s->check_termination = 25; s->scaling = 10;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 0;
  s->polish = 1; s->polish_refine_iter = 3; s->scaled_termination = 0;
}

int lpvmpc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char *lpvmpc_last_error(const lpvmpc_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lpvmpc_create(const lpvmpc_cfg *cfg, lpvmpc_handle **out) {
  if (!cfg || !out) return fail(nullptr, LPVMPC_E_ARG, "null cfg/out");
  *out = nullptr;
  if (cfg->abi_version != LPVMPC_ABI_VERSION) return fail(nullptr, LPVMPC_E_ARG, "abi_version mismatch");
  if (cfg->kind != LPVMPC_CONTROLLER && cfg->kind != LPVMPC_PLANNER) return fail(nullptr, LPVMPC_E_ARG, "bad kind");
  if (cfg->N < 2 || cfg->N > 4096) return fail(nullptr, LPVMPC_E_ARG, "horizon N must be in [2, 4096]");
  if (cfg->max_batch < 1) return fail(nullptr, LPVMPC_E_ARG, "max_batch must be >= 1");
  if (!cfg->track || cfg->n_track_seg < 1) return fail(nullptr, LPVMPC_E_ARG, "track table required");
  if (cfg->steering_delay < 0 || cfg->steering_delay > cfg->N || (cfg->kind == LPVMPC_PLANNER && cfg->steering_delay))
    return fail(nullptr, LPVMPC_E_ARG, "bad steering_delay");
  if (const char *why = settings_error(&cfg->settings)) return fail(nullptr, LPVMPC_E_ARG, why);
  {
    // Curvature() reduces s modulo TrackLength = start + length of the last segment: it must be a positive finite number
    const double *last = cfg->track + 6 * (size_t)(cfg->n_track_seg - 1);
    const double track_len = last[3] + last[4];
    if (!(track_len > 0.0) || !(track_len <= 1e300)) return fail(nullptr, LPVMPC_E_ARG, "track table: TrackLength must be positive and finite");
    for (int i = 0; i < 6 * cfg->n_track_seg; ++i)
      if (!(cfg->track[i] == cfg->track[i])) return fail(nullptr, LPVMPC_E_ARG, "track table holds NaN");
  }
  if (!(cfg->dt > 0.0)) return fail(nullptr, LPVMPC_E_ARG, "dt must be > 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, LPVMPC_E_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, LPVMPC_E_ARG, "bad device ordinal");
  lpvmpc_handle *h = new (std::nothrow) lpvmpc_handle();
  if (!h) return fail(nullptr, LPVMPC_E_ARG, "out of host memory");
  h->cfg = *cfg;
  h->cfg.track = nullptr;
  h->device = cfg->device;
  h->n = cfg->kind == LPVMPC_CONTROLLER ? 6 : 5;
  h->d = 2;
  h->L = make_layout(cfg->kind, cfg->N, cfg->steering_delay);
  auto bail = [&](int code) { std::string m = h->err; lpvmpc_destroy(h); g_create_error = m; return code; };
#define CTRY(expr)                                                                                   \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) { h->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return bail(LPVMPC_E_CUDA); } \
  } while (0)
  DeviceGuard dev_guard(h->device);
  CTRY(dev_guard.err);
  cudaDeviceProp prop;
  CTRY(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major < 10) { h->err = "built for sm_100a (B200); device is older"; return bail(LPVMPC_E_UNSUPPORTED); }
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  {
    // the specialised kernels need diagonal Q and R and no steering delay; planner rows use 64-bit stage masks (N <= 63)
    bool pdiag = true;
    const int nxk = cfg->kind == LPVMPC_CONTROLLER ? 6 : 5;
    for (int i = 0; i < nxk; ++i) for (int j = 0; j < nxk; ++j) if (i != j && cfg->Q[i * nxk + j] != 0.0) pdiag = false;
    if (cfg->R[1] != 0.0 || cfg->R[2] != 0.0) pdiag = false;
    h->variant = 1;
#ifdef LPVMPC_LEGACY
    {
      // T8 kernel: controller, N = 8, no steering delay, diagonal Q and R
      const bool eligible = cfg->kind == LPVMPC_CONTROLLER && cfg->N == kT8N && cfg->steering_delay == 0 && pdiag;
      if ((cfg->variant == 2 || cfg->variant == 4) && !eligible) { h->err = "variants 2/4 (T8) need controller, N=8, steering_delay=0, diagonal Q and R"; return bail(LPVMPC_E_UNSUPPORTED); }
      if (cfg->variant == 2 || cfg->variant == 4) h->variant = 2;
      h->qpw = (cfg->variant == 4) ? 2 : 4;
      // G8 kernel: per-QP state fits shared memory
      h->GL = make_g8_layout(cfg->kind, cfg->N);
      const size_t per_qp = (size_t)h->GL.total * sizeof(double);
      const bool g8_ok = pdiag && cfg->steering_delay == 0 && per_qp <= (size_t)h->smem_optin &&
                         (cfg->kind == LPVMPC_CONTROLLER || cfg->N <= 63);
      if (cfg->variant == 3 && !g8_ok) { h->err = "variant 3 (G8) needs diagonal Q and R, steering_delay=0, planner N<=63 and a per-QP state that fits shared memory"; return bail(LPVMPC_E_UNSUPPORTED); }
      if (cfg->variant == 3) h->variant = 3;
    }
#else
    if (cfg->variant >= 2 && cfg->variant <= 4) { h->err = "variants 2-4 (T8 / G8, earlier kernel generations) are only in builds with -DLPVMPC_LEGACY"; return bail(LPVMPC_E_UNSUPPORTED); }
#endif
    if (cfg->variant < 0 || cfg->variant > 8) { h->err = "unknown kernel variant"; return bail(LPVMPC_E_ARG); }
    // H8 kernel: same restrictions as G8, smaller shared-memory footprint (cold data in an L2 slab)
    h->HL = make_h8_layout(cfg->kind, cfg->N, 0);
    const bool h8_ok = lpv::h8::layout_matches(h->HL, cfg->kind == LPVMPC_CONTROLLER ? 6 : 5) && pdiag && cfg->steering_delay == 0 && (size_t)h->HL.total * sizeof(double) + 1024 <= (size_t)h->smem_optin &&
                       (cfg->kind == LPVMPC_CONTROLLER || cfg->N <= 63);
    if (cfg->variant == 5 && !h8_ok) { h->err = "variant 5 (H8) needs diagonal Q and R, steering_delay=0, planner N<=63 and a per-QP state that fits shared memory"; return bail(LPVMPC_E_UNSUPPORTED); }
    if (cfg->variant == 5 || (cfg->variant == 0 && h8_ok)) h->variant = 5;
    // H8S = H8 with the factor streamed from the slab through a ring of 4 stage blocks (variant 7): N <= 254, stage
    // vectors and single-variable rows must still fit shared memory
    const lpv::h8::Lay SL4 = make_h8_layout(cfg->kind, cfg->N, 4);
    const bool h8s_ok = lpv::h8::layout_matches(SL4, cfg->kind == LPVMPC_CONTROLLER ? 6 : 5) && pdiag && cfg->steering_delay == 0 && cfg->N <= 254 && (size_t)SL4.total * sizeof(double) + 2048 <= (size_t)h->smem_optin &&
                        (cfg->kind == LPVMPC_CONTROLLER || cfg->N <= 63);
    if (cfg->variant == 7 && !h8s_ok) { h->err = "variant 7 (H8S) needs diagonal Q and R, steering_delay=0, planner N<=63, N<=254"; return bail(LPVMPC_E_UNSUPPORTED); }
    if (cfg->variant == 7) { h->variant = 5; h->HL = SL4; }
    // auto: the resident factor wins whenever it fits (planner N=40: 352 vs 394 ms, controller N=100: 73 vs 85 ms per batch,
    // profiles/r1v_*: the batches are bound by the longest chains, and staging adds latency to every stage step); the
    // streamed factor takes over when the factor no longer fits shared memory (controller N > ~135)
    else if (cfg->variant == 0 && h8s_ok && !h8_ok) { h->variant = 5; h->HL = SL4; }
    // H8T kernel: controller, block factor in tensor memory (48 N + 16 columns of the SM's 512), 16 QPs per CTA, one CTA per SM
    h->TL = make_h8t_layout(cfg->kind, cfg->N);
    const size_t h8t_bytes = (size_t)h->TL.total * sizeof(double) * 16 + 4 * 512 + 512;
    const bool h8t_ok = pdiag && cfg->steering_delay == 0 && cfg->kind == LPVMPC_CONTROLLER && 60 * cfg->N + 28 <= 512 &&
                        h8t_bytes + 64 <= (size_t)h->smem_optin;
    if (cfg->variant == 6 && !h8t_ok) { h->err = "variant 6 (H8T) needs controller, diagonal Q and R, steering_delay=0, N<=8 (60 N + 28 tensor-memory columns <= 512)"; return bail(LPVMPC_E_UNSUPPORTED); }
    if (cfg->variant == 6 || (cfg->variant == 0 && h8t_ok)) h->variant = 6;
    // H16T kernel: 16 lanes per QP, twisted factorisation (both ends towards the middle stage), 8 warps per CTA; the
    // reference's controller horizon only (the factorisation scratch borrows stage-vector rows 0..7; N = 2 NL)
    h->TL16 = make_h16t_layout(cfg->kind, cfg->N);
    const size_t h16t_bytes = (size_t)h->TL16.total * sizeof(double) * 16 + 8 * 512 + 512;
    const bool h16t_ok = pdiag && cfg->steering_delay == 0 && cfg->kind == LPVMPC_CONTROLLER && cfg->N == 8 &&
                         h16t_bytes + 64 <= (size_t)h->smem_optin && lpv::h16t::layout_matches(h->TL16);   // (the kernel holds the layout as constants)
    if (cfg->variant == 8 && !h16t_ok) { h->err = "variant 8 (H16T) needs controller, N=8, diagonal Q and R, steering_delay=0"; return bail(LPVMPC_E_UNSUPPORTED); }
    static const bool auto16 = [] { const char *e = std::getenv("LPVMPC_AUTO_H16T"); return !e || std::atoi(e) != 0; }();   // LPVMPC_AUTO_H16T=0: H8T
    if (cfg->variant == 8 || (cfg->variant == 0 && h16t_ok && auto16)) h->variant = 8;
  }
  h->ws_bytes = (size_t)h->L.total * sizeof(double);
  h->smem_mode = h->ws_bytes <= (size_t)h->smem_optin;
  if (h->variant == 8) {
    h->qpw = 2; h->wpc = 8;
    h->ws_bytes = (size_t)h->TL16.total * sizeof(double) * 16 + 8 * 512 + 512;
    h->smem_mode = true;
    h->grid_cap = h->sm_count;
    CTRY((cudaFuncSetAttribute(lpv::h16t::lpv_solve_h16t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ws_bytes)));
    CTRY(cudaMalloc(&h->d_queue, sizeof(unsigned)));
    CTRY(cudaMalloc(&h->d_cold, sizeof(double) * (size_t)h->grid_cap * 16 * h->TL16.cold_total));
  } else if (h->variant == 6) {
    h->qpw = 4; h->wpc = 4;
    h->ws_bytes = (size_t)h->TL.total * sizeof(double) * 16 + 4 * 512 + 512;
    h->smem_mode = true;
    h->grid_cap = h->sm_count;
    CTRY((cudaFuncSetAttribute(lpv::h8t::lpv_solve_h8t_kernel<LPVMPC_CONTROLLER, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ws_bytes)));
    CTRY(cudaMalloc(&h->d_queue, sizeof(unsigned)));
    CTRY(cudaMalloc(&h->d_cold, sizeof(double) * (size_t)h->grid_cap * 16 * h->TL.cold_total));
  } else if (h->variant == 5) {
    // pick (QPs per warp, warps per CTA) that keeps the most QPs resident per SM
    const size_t per_qp = (size_t)h->HL.total * sizeof(double);
    const size_t sm_bytes = prop.sharedMemPerMultiprocessor;
    int best_q = 0, best_qpw = 1, best_wpc = 1, best_ctas = 1;
    // resident factor: prefer many QPs per warp; streamed factor: prefer many warps (ties go to the first candidate)
    const int qs_res[3] = {4, 2, 1}, qs_str[3] = {1, 1, 1};   // the streamed kernel: one QP per warp, one warp per CTA
    const int *qs = h->HL.ring ? qs_str : qs_res;
    for (int qi = 0; qi < 3; ++qi) for (int w = (h->HL.ring ? 1 : 2); w >= 1; --w) {
      const size_t cta = per_qp * qs[qi] * w + 512 * (size_t)w + 512 + 64 * (size_t)(qs[qi] * w);
      if (cta > (size_t)h->smem_optin) continue;
      int ctas = (int)(sm_bytes / (cta + 1024));
      if (ctas > 32) ctas = 32;
      const int q = ctas * qs[qi] * w;
      if (q > best_q) { best_q = q; best_qpw = qs[qi]; best_wpc = w; best_ctas = ctas; }
    }
    // tuning knobs for experiments (not part of the ABI): force QPs per warp / cap resident CTAs per SM
    if (const char *e = std::getenv("LPVMPC_H8_QPW")) {
      const int q = std::atoi(e);
      const size_t cta = per_qp * q + 512 + 512 + 64 * (size_t)q;
      if (!h->HL.ring && (q == 1 || q == 2 || q == 4) && cta <= (size_t)h->smem_optin) {
        best_qpw = q; best_wpc = 1; best_ctas = (int)(sm_bytes / (cta + 1024));
        if (best_ctas > 32) best_ctas = 32;
      }
    }
    if (const char *e = std::getenv("LPVMPC_H8_CTAS")) { const int v = std::atoi(e); if (v >= 1 && v < best_ctas) best_ctas = v; }
    h->qpw = best_qpw; h->wpc = best_wpc;
    h->ws_bytes = per_qp * h->qpw * h->wpc + 512 * (size_t)h->wpc + 512 + 64 * (size_t)(h->qpw * h->wpc);
    h->smem_mode = true;
    h->grid_cap = h->sm_count * best_ctas;
    const bool ctrl = cfg->kind == LPVMPC_CONTROLLER;
    if (h->HL.ring) CTRY(ctrl ? h8s_attr<LPVMPC_CONTROLLER>(h->ws_bytes) : h8s_attr<LPVMPC_PLANNER>(h->ws_bytes));
    else if (h->qpw == 4) CTRY(ctrl ? (h8_attr<LPVMPC_CONTROLLER, 4>(h->ws_bytes)) : (h8_attr<LPVMPC_PLANNER, 4>(h->ws_bytes)));
    else if (h->qpw == 2) CTRY(ctrl ? (h8_attr<LPVMPC_CONTROLLER, 2>(h->ws_bytes)) : (h8_attr<LPVMPC_PLANNER, 2>(h->ws_bytes)));
    else {
      // one QP per warp, even horizon: twisted factorisation (both ends towards the middle stage); LPVMPC_H8_TWISTED=0: plain
      static const bool tw_on = [] { const char *e = std::getenv("LPVMPC_H8_TWISTED"); return !e || std::atoi(e) != 0; }();
      h->twisted = tw_on && (cfg->N % 2 == 0) && cfg->N >= 4;
      // helper warps for the element-wise updates (LPVMPC_H8_HELPERS=0: none): one QP per CTA, so the choice only stands when
      // it was one warp per CTA anyway.  Controller (long horizons, one CTA per SM): 7 helpers (the cold phases split 32 ways).  Planner (3 CTAs per SM): ONE --
      // a warp's registers live in its SM sub-partition (16 K registers), so at 224-255 registers per thread an SM holds 8
      // warps: 3 CTAs of 2.  3 CTAs of 3 or 4 warps need <= 168 registers: with that cap forced (__launch_bounds__(96, 4) /
      // (128, 3): 300 bytes of spills) plan16384 takes 319 / 315 ms against 297 ms with one helper at 255 registers, same box
      static const bool hw_on = [] { const char *e = std::getenv("LPVMPC_H8_HELPERS"); return !e || std::atoi(e) != 0; }();
      static const int ctrl_nh = [] { const char *e = std::getenv("LPVMPC_H8_CTRL_HELPERS"); return (e && std::atoi(e) == 3) ? 3 : 7; }();   // LPVMPC_H8_CTRL_HELPERS=3: three helpers (ctrl1024N100 23.23 ms against 22.44 ms with seven, same binary)
      h->helpers = (h->twisted && hw_on && h->wpc == 1) ? (ctrl ? ctrl_nh : 1) : 0;
      if (h->helpers) {
        const size_t ws2 = h->ws_bytes + 512 + 1536;   // the helpers' mailbox (lpv::h8::HwShared) behind the gather buffers, then the CTA's reduction scratch (4 warps x 16 doubles)
        int ctas = (int)(sm_bytes / (ws2 + 1024));
        if (ctas > 32) ctas = 32;
        const int by_regs = ctrl ? 1 : 3;      // __launch_bounds__ of the two instantiations
        if (ctas > by_regs) ctas = by_regs;
        if (ctas >= best_ctas) {   // never at the price of fewer QPs per SM
          h->ws_bytes = ws2; h->grid_cap = h->sm_count * (ctas < 1 ? 1 : ctas);
          CTRY(ctrl ? (h->helpers == 7 ? h8h_attr<LPVMPC_CONTROLLER, 7>(h->ws_bytes) : h8h_attr<LPVMPC_CONTROLLER, 3>(h->ws_bytes)) : (h8h_attr<LPVMPC_PLANNER, 1>(h->ws_bytes)));
        } else h->helpers = 0;
      }
      if (h->helpers) {}
      else if (h->twisted) CTRY(ctrl ? h8w_attr<LPVMPC_CONTROLLER>(h->ws_bytes) : h8w_attr<LPVMPC_PLANNER>(h->ws_bytes));
      else CTRY(ctrl ? (h8_attr<LPVMPC_CONTROLLER, 1>(h->ws_bytes)) : (h8_attr<LPVMPC_PLANNER, 1>(h->ws_bytes)));
    }
    CTRY(cudaMalloc(&h->d_queue, sizeof(unsigned)));
    CTRY(cudaMalloc(&h->d_cold, sizeof(double) * (size_t)h->grid_cap * h->wpc * h->qpw * h->HL.cold_total));
#ifdef LPVMPC_LEGACY
  } else if (h->variant == 3) {
    const size_t per_qp = (size_t)h->GL.total * sizeof(double);
    h->qpw = (4 * per_qp <= (size_t)h->smem_optin) ? 4 : ((2 * per_qp <= (size_t)h->smem_optin) ? 2 : 1);
    h->ws_bytes = per_qp * h->qpw;
    h->smem_mode = true;
    int per_sm = (int)(prop.sharedMemPerMultiprocessor / (h->ws_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    h->grid_cap = h->sm_count * per_sm;
    const bool ctrl = cfg->kind == LPVMPC_CONTROLLER;
    if (h->qpw == 4) CTRY(ctrl ? (g8_attr<LPVMPC_CONTROLLER, 4>(h->ws_bytes)) : (g8_attr<LPVMPC_PLANNER, 4>(h->ws_bytes)));
    else if (h->qpw == 2) CTRY(ctrl ? (g8_attr<LPVMPC_CONTROLLER, 2>(h->ws_bytes)) : (g8_attr<LPVMPC_PLANNER, 2>(h->ws_bytes)));
    else CTRY(ctrl ? (g8_attr<LPVMPC_CONTROLLER, 1>(h->ws_bytes)) : (g8_attr<LPVMPC_PLANNER, 1>(h->ws_bytes)));
    CTRY(cudaMalloc(&h->d_queue, sizeof(unsigned)));
    CTRY(cudaMalloc(&h->d_cold, sizeof(double) * (size_t)h->grid_cap * h->qpw * h->GL.cold_total));
  } else if (h->variant == 2) {
    // variant 4 = experiment: 2 QPs per warp, 8 warps / SM (slower: the fetch limit is per SM sub-partition)
    h->ws_bytes = (size_t)h->qpw * lpv::t8::Reg<kT8N>::TOTAL * sizeof(double);
    h->smem_mode = true;
    int per_sm = (int)(prop.sharedMemPerMultiprocessor / (h->ws_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16 / h->qpw) per_sm = 16 / h->qpw;
    h->grid_cap = h->sm_count * per_sm;
    if (h->qpw == 2) CTRY((cudaFuncSetAttribute(lpv::t8::lpv_solve_t8_kernel<kT8N, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ws_bytes)));
    else CTRY((cudaFuncSetAttribute(lpv::t8::lpv_solve_t8_kernel<kT8N, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ws_bytes)));
    CTRY(cudaMalloc(&h->d_queue, sizeof(unsigned)));
    CTRY(cudaMalloc(&h->d_cold, sizeof(double) * (size_t)h->grid_cap * 32 * lpv::t8::Cold<kT8N>::TOTAL));
#endif
  } else if (h->smem_mode) {
    int per_sm = (int)(prop.sharedMemPerMultiprocessor / (h->ws_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 32) per_sm = 32;
    h->grid_cap = h->sm_count * per_sm;
    if (cfg->kind == LPVMPC_CONTROLLER)
      CTRY(cudaFuncSetAttribute(lpv::lpv_solve_kernel<LPVMPC_CONTROLLER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ws_bytes));
    else
      CTRY(cudaFuncSetAttribute(lpv::lpv_solve_kernel<LPVMPC_PLANNER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ws_bytes));
  } else {
    h->grid_cap = h->sm_count * 16;
    if (h->grid_cap > cfg->max_batch) h->grid_cap = cfg->max_batch;
    CTRY(cudaMalloc(&h->d_gws, h->ws_bytes * (size_t)h->grid_cap));
  }
  CTRY(cudaMalloc(&h->d_perm, sizeof(int) * (size_t)cfg->max_batch));
  CTRY(cudaMalloc(&h->d_track, sizeof(double) * 6 * (size_t)cfg->n_track_seg));
  CTRY(cudaMemcpy(h->d_track, cfg->track, sizeof(double) * 6 * (size_t)cfg->n_track_seg, cudaMemcpyHostToDevice));
  CTRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CTRY(cudaEventCreateWithFlags(&h->last_ev, cudaEventDisableTiming));
  lpv::Model &M = h->M;
  M.dt = cfg->dt; M.lf = cfg->lf; M.lr = cfg->lr; M.m = cfg->m; M.Iz = cfg->Iz; M.Cf = cfg->Cf; M.Cr = cfg->Cr; M.mu = cfg->mu;
  M.max_vel = cfg->max_vel; M.min_vel = cfg->min_vel;
  std::memcpy(M.Q, cfg->Q, sizeof(M.Q)); std::memcpy(M.R, cfg->R, sizeof(M.R));
  std::memcpy(M.dR, cfg->dR, sizeof(M.dR)); std::memcpy(M.L_cf, cfg->L_cf, sizeof(M.L_cf));
  M.track = h->d_track; M.nseg = cfg->n_track_seg; M.delay = cfg->steering_delay;
  // staging arena for the host API
  size_t per = 0;
  for (const Field &f : staged_fields(h)) per += align256(f.elem * (size_t)cfg->max_batch);
  per += align256(sizeof(int32_t) * (size_t)cfg->max_batch);  // sched_err of lpvmpc_schedule_host
  h->stage_bytes = per;
  CTRY(cudaMalloc(&h->d_stage, per));
  CTRY(cudaMallocHost(&h->h_stage, per));
  if (const char *e = std::getenv("LPVMPC_ZERO_COPY_OUT")) h->zc_out = std::atoi(e) != 0;
  h->zc_in = (h->variant == 8);
  if (const char *e = std::getenv("LPVMPC_ZERO_COPY_IN")) h->zc_in = std::atoi(e) != 0;
#undef CTRY
  *out = h;
  return LPVMPC_OK;
}

void lpvmpc_destroy(lpvmpc_handle *h) {
  if (!h) return;
  DeviceGuard dev_guard(h->device);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->last_ev) cudaEventDestroy(h->last_ev);
  cudaFree(h->d_perm);
  cudaFree(h->d_track); cudaFree(h->d_gws); cudaFree(h->d_stage); cudaFree(h->d_queue); cudaFree(h->d_cold);
  cudaFree(h->d_loop); cudaFree(h->d_ploop); cudaFree(h->d_refs_W);
  if (h->loop_ev) cudaEventDestroy(h->loop_ev);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->h_ring) cudaFreeHost(h->h_ring);
  delete h;
}

int lpvmpc_get_info(const lpvmpc_handle *h, lpvmpc_info *info) {
  if (!h || !info) return LPVMPC_E_ARG;
  info->n = h->n; info->d = h->d; info->N = h->L.N; info->nz = h->L.nz; info->m = h->L.m;
  info->variant = (h->variant == 5 && h->HL.ring) ? 7 : h->variant;   // 7: H8 with the factor streamed
  info->workspace_in_smem = h->smem_mode ? 1 : 0;
  info->smem_bytes_per_qp = h->smem_mode ? (int)(h->variant == 8 ? (size_t)h->TL16.total * sizeof(double) : h->variant == 6 ? (size_t)h->TL.total * sizeof(double) : h->variant == 5 ? (size_t)h->HL.total * sizeof(double) : ((h->variant == 2 || h->variant == 3) ? h->ws_bytes / h->qpw : h->ws_bytes)) : 0;
  info->workspace_bytes = (long long)(h->stage_bytes + (h->smem_mode ? 0 : h->ws_bytes * (size_t)h->grid_cap));
  info->kernel_launches = h->launches;
  return LPVMPC_OK;
}

int lpvmpc_update_settings(lpvmpc_handle *h, const lpvmpc_settings *s) {
  if (!h || !s) return LPVMPC_E_ARG;
  if (const char *why = settings_error(s)) return fail(h, LPVMPC_E_ARG, why);
  h->cfg.settings = *s;
  return LPVMPC_OK;
}

int lpvmpc_solve_dev(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, void *stream) {
  int rc = validate_args(h, B, a, true);
  if (rc) return rc;
  ON_DEVICE(h);
  if ((rc = stream_enter(h, (cudaStream_t)stream))) return rc;
  const Params p = make_params(h, B, a);
  rc = h->cfg.kind == LPVMPC_CONTROLLER ? launch_solve<LPVMPC_CONTROLLER>(h, p, (cudaStream_t)stream)
                                        : launch_solve<LPVMPC_PLANNER>(h, p, (cudaStream_t)stream);
  return rc ? rc : stream_leave(h, (cudaStream_t)stream);
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn sched_maps_fn() {
  static TensorMapEncodeFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (TensorMapEncodeFn)f;
  }();
  return fn;
}

int lpvmpc_schedule_dev(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, int32_t *sched_err, void *stream) {
  int rc = validate_args(h, B, a, false);
  if (rc) return rc;
  if (a->sched_mode == LPVMPC_SCHED_GIVEN) return fail(h, LPVMPC_E_ARG, "schedule needs PREDICT or ESTIMATE");
  if (B == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  if ((rc = stream_enter(h, (cudaStream_t)stream))) return rc;
  const Params p = make_params(h, B, a);
  static const bool naive = [] { const char *e = std::getenv("LPVMPC_SCHED_NAIVE"); return e && std::atoi(e) != 0; }();
  // the tiled kernel stores 16-byte pieces: output arrays that are only 8-byte aligned take the plain kernel
  const bool misaligned = (((uintptr_t)a->A_out | (uintptr_t)a->B_out | (uintptr_t)a->states_out) & 15u) != 0;
  if (naive || misaligned) {
    const int g = (B + 127) / 128;
    if (h->cfg.kind == LPVMPC_CONTROLLER) lpv::lpv_schedule_naive_kernel<LPVMPC_CONTROLLER><<<g, 128, 0, (cudaStream_t)stream>>>(p, sched_err);
    else lpv::lpv_schedule_naive_kernel<LPVMPC_PLANNER><<<g, 128, 0, (cudaStream_t)stream>>>(p, sched_err);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
    return stream_leave(h, (cudaStream_t)stream);
  }
  const int grid = (B + lpv::kSchedThreads - 1) / lpv::kSchedThreads;
  // controller, 16-byte aligned outputs: dense shared-memory boxes + TMA tensor stores (LPVMPC_SCHED_MODE: 2 = TMA (default),
  // 1 = 256-bit stores straight from registers (needs 32-byte alignment), 0 = tile-staged 16-byte stores)
  static const int sched_mode = [] { const char *e = std::getenv("LPVMPC_SCHED_MODE"); return e ? std::atoi(e) : 2; }();
  if (h->cfg.kind == LPVMPC_CONTROLLER && sched_mode == 2 && sched_maps_fn() && ((uintptr_t)a->B_out & 31u) == 0) {
    lpv::SchedMaps maps;
    std::memset(&maps, 0, sizeof(maps));
    const int N = h->L.N;
    bool ok = true;
    auto enc = [&](CUtensorMap *m, double *base, int len, int stages) {
      if (!base) return;
      const cuuint64_t gdim[3] = {(cuuint64_t)len, (cuuint64_t)N, (cuuint64_t)B};
      const cuuint64_t gstr[2] = {(cuuint64_t)len * 8, (cuuint64_t)len * 8 * (cuuint64_t)N};
      const cuuint32_t box[3] = {(cuuint32_t)len, (cuuint32_t)stages, 32}, es[3] = {1, 1, 1};
      ok = ok && sched_maps_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    enc(&maps.A, a->A_out, 36, 1);
    if (a->sched_mode != LPVMPC_SCHED_ESTIMATE) enc(&maps.S, a->states_out, 6, lpv::kTmaChunk);
    if (ok) {
      const int g2 = (B + 32 * lpv::kTmaWarps - 1) / (32 * lpv::kTmaWarps);
      lpv::lpv_schedule_tma_kernel<<<g2, 32 * lpv::kTmaWarps, 0, (cudaStream_t)stream>>>(p, sched_err, maps);
      ++h->launches;
      CUDA_TRY(h, cudaGetLastError());
      return stream_leave(h, (cudaStream_t)stream);
    }
  }
  const bool direct_on = sched_mode == 1;
  const bool al32 = (((uintptr_t)a->A_out | (uintptr_t)a->B_out) & 31u) == 0;
  if (h->cfg.kind == LPVMPC_CONTROLLER && direct_on && al32)
    lpv::lpv_schedule_kernel<LPVMPC_CONTROLLER, true><<<grid, lpv::kSchedThreads, 0, (cudaStream_t)stream>>>(p, sched_err);
  else if (h->cfg.kind == LPVMPC_CONTROLLER) lpv::lpv_schedule_kernel<LPVMPC_CONTROLLER><<<grid, lpv::kSchedThreads, 0, (cudaStream_t)stream>>>(p, sched_err);
  else lpv::lpv_schedule_kernel<LPVMPC_PLANNER><<<grid, lpv::kSchedThreads, 0, (cudaStream_t)stream>>>(p, sched_err);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return stream_leave(h, (cudaStream_t)stream);
}

// Host-pointer variants: pack every non-NULL input into one pinned arena, one H2D, kernel, one D2H.
// `views` (lpvmpc_solve_host_view): the results stay in pinned memory -- one of two result arenas used in turn -- and the
// caller gets pointers into it instead of copies into its own arrays (valid until the call after the next one)
static int run_host(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, int32_t *sched_err, bool solve, lpvmpc_args *views = nullptr) {
  int rc = validate_args(h, B, a, solve);
  if (rc) return rc;
  if (views) std::memset(views, 0, sizeof(*views));
  if (B == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  lpvmpc_args dev = *a;
  const bool zin = solve && h->zc_in;
  const std::vector<Field> fields = staged_fields(h);
  char *ring = nullptr;   // this call's result arena (view mode)
  if (views) {
    if (!h->h_ring) {
      size_t ob = 0;
      for (const Field &f : fields) if (f.output) ob += align256(f.elem * (size_t)h->cfg.max_batch);
      h->ring_bytes = ob;
      CUDA_TRY(h, cudaMallocHost(&h->h_ring, 2 * ob));
    }
    h->ring_slot ^= 1;
    ring = h->h_ring + (size_t)h->ring_slot * h->ring_bytes;
  }
  size_t off = 0, in_end = 0, out_begin = 0;
  bool first_out = true;
  std::vector<size_t> offs(fields.size());
  std::vector<CopyJob> jobs;
  for (size_t i = 0; i < fields.size(); ++i) {
    const Field &f = fields[i];
    if (f.output && first_out) { out_begin = off; first_out = false; }
    offs[i] = off;
    if (cptr_at(a, f.off_args)) {
      const size_t bytes = f.elem * (size_t)B;
      if (!f.output) jobs.push_back({h->h_stage + off, cptr_at(a, f.off_args), bytes});
      // results: written by the kernel through the pinned arena's device mapping (posted PCIe writes behind the compute:
      // no D2H copy after the kernel); the kernels only ever write their outputs
      ptr_at(&dev, f.off_args) = ((f.output ? h->zc_out : zin) ? ((f.output && ring) ? ring - out_begin : h->h_stage) : h->d_stage) + off;
      off += align256(bytes);
      if (!f.output) in_end = off;
    }
  }
  run_copies(jobs);
  size_t se_off = off;
  int32_t *d_se = nullptr;
  if (sched_err) { d_se = reinterpret_cast<int32_t *>(h->d_stage + off); off += align256(sizeof(int32_t) * (size_t)B); }
  if (off > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow");
  if (in_end && !zin) CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, in_end, cudaMemcpyHostToDevice, h->stream));
  rc = solve ? lpvmpc_solve_dev(h, B, &dev, h->stream) : lpvmpc_schedule_dev(h, B, &dev, d_se, h->stream);
  if (rc) return rc;
  const size_t d2h_begin = (h->zc_out && !first_out) ? se_off : out_begin;   // zero-copy results: only sched_err comes back by copy
  if (off > d2h_begin) {
    if (ring && !h->zc_out && se_off > out_begin) {   // view mode without zero-copy results: the D2H copy lands in the result arena
      CUDA_TRY(h, cudaMemcpyAsync(ring, h->d_stage + out_begin, se_off - out_begin, cudaMemcpyDeviceToHost, h->stream));
      if (off > se_off) CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + se_off, h->d_stage + se_off, off - se_off, cudaMemcpyDeviceToHost, h->stream));
    } else {
      CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + d2h_begin, h->d_stage + d2h_begin, off - d2h_begin, cudaMemcpyDeviceToHost, h->stream));
    }
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (views) {
    for (size_t i = 0; i < fields.size(); ++i) {
      const Field &f = fields[i];
      if (f.output && cptr_at(a, f.off_args)) ptr_at(views, f.off_args) = ring + (offs[i] - out_begin);
    }
    if (sched_err) std::memcpy(sched_err, h->h_stage + se_off, sizeof(int32_t) * (size_t)B);
    return LPVMPC_OK;
  }
  jobs.clear();
  for (size_t i = 0; i < fields.size(); ++i) {
    const Field &f = fields[i];
    if (f.output && cptr_at(a, f.off_args))
      jobs.push_back({const_cast<void *>(cptr_at(a, f.off_args)), h->h_stage + offs[i], f.elem * (size_t)B});
  }
  run_copies(jobs);
  if (sched_err) std::memcpy(sched_err, h->h_stage + se_off, sizeof(int32_t) * (size_t)B);
  return LPVMPC_OK;
}

int lpvmpc_solve_host(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a) { return run_host(h, B, a, nullptr, true); }
int lpvmpc_solve_host_view(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, lpvmpc_args *views) {
  if (!views) return fail(h, LPVMPC_E_ARG, "null views");
  return run_host(h, B, a, nullptr, true, views);
}

// ------------------------------------------------------------------------------------------------
// closed-loop fleet
void lpvmpc_loop_default_cfg(lpvmpc_loop_cfg *c) {
  if (!c) return;
  c->sim_dt = 0.005; c->substeps = 7; c->warmup_ticks = 9; c->swap_ey_epsi = 1; c->reserved = 0;
  c->vel_ref = 1.0; c->Cf_new = 60.0; c->half_width = 0.3; c->slack = 0.45; c->sim_mu = 0.05;
}

int lpvmpc_loop_init_dev(lpvmpc_handle *h, int32_t B, const lpvmpc_loop_cfg *c, const double *sim0, void *stream) {
  if (!h || !c || !sim0) return fail(h, LPVMPC_E_ARG, "null handle/cfg/sim0");
  if (h->cfg.kind != LPVMPC_CONTROLLER || h->cfg.steering_delay != 0 || h->L.N > 20)
    return fail(h, LPVMPC_E_UNSUPPORTED, "the closed loop needs a controller handle with N <= 20 and steering_delay = 0");
  if (B < 1 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "fleet size must be in [1, max_batch]");
  if (c->substeps < 0 || c->warmup_ticks < 0 || !(c->sim_dt > 0)) return fail(h, LPVMPC_E_ARG, "bad loop cfg");
  ON_DEVICE(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = stream_enter(h, s)) return rc;
  const int N = h->L.N;
  const size_t D = sizeof(double), mb = (size_t)h->cfg.max_batch;
  if (!h->d_loop) {
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t r = off; off += align256(bytes * mb); return r; };
    const size_t o_sim = take(8 * D), o_cmd = take(2 * D), o_up = take(2 * N * D), o_xp = take(6 * (N + 1) * D),
                 o_loc = take(6 * D), o_stat = take(4 * D), o_ctr = take(8 * sizeof(int32_t)), o_x0 = take(6 * D),
                 o_uprev = take(2 * N * D), o_traj = take(6 * N * D), o_uold = take(2 * D), o_vel = take((N + 1) * D),
                 o_status = take(sizeof(int32_t)), o_iters = take(sizeof(int32_t));
    CUDA_TRY(h, cudaMalloc(&h->d_loop, off));
    char *base = h->d_loop;
    lpv::loop::LoopParams &P = h->LP;
    P.sim = (double *)(base + o_sim); P.cmd = (double *)(base + o_cmd); P.u_pred = (double *)(base + o_up);
    h->d_loop_xpred = (double *)(base + o_xp); P.local = (double *)(base + o_loc); P.stat = (double *)(base + o_stat);
    P.ctr = (int *)(base + o_ctr); P.x0 = (double *)(base + o_x0); P.u_prev = (double *)(base + o_uprev);
    P.traj = (double *)(base + o_traj); P.u_old = (double *)(base + o_uold); h->d_loop_velref = (double *)(base + o_vel);
    h->d_loop_status = (int32_t *)(base + o_status); h->d_loop_iters = (int32_t *)(base + o_iters);
    P.status = h->d_loop_status; P.iters = h->d_loop_iters;
  }
  lpv::loop::LoopParams &P = h->LP;
  P.lc = *c; P.lf = h->M.lf; P.lr = h->M.lr; P.m = h->M.m; P.Iz = h->M.Iz; P.track = h->d_track; P.nseg = h->M.nseg;
  P.N = N; P.B = B;
  h->loop_B = B; h->loop_tick = 0;
  // state: sim <- sim0, everything else zero, first_it = 1, lap tick = -1, vel_ref = const
  CUDA_TRY(h, cudaMemcpyAsync(P.sim, sim0, 8 * D * (size_t)B, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(h, cudaMemsetAsync(P.cmd, 0, 2 * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(P.u_pred, 0, 2 * N * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(h->d_loop_xpred, 0, 6 * (N + 1) * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(P.local, 0, 6 * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(h->d_loop_status, 0, sizeof(int32_t) * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(h->d_loop_iters, 0, sizeof(int32_t) * (size_t)B, s));
  {
    std::vector<int32_t> ctr((size_t)B * 8, 0);
    std::vector<double> stat((size_t)B * 4, 0.0), vel((size_t)B * (N + 1), c->vel_ref);
    for (int b = 0; b < B; ++b) { ctr[(size_t)b * 8] = 1; stat[(size_t)b * 4 + 3] = -1.0; }
    // pageable sources: the copies are staged before the calls return
    CUDA_TRY(h, cudaMemcpyAsync(P.ctr, ctr.data(), ctr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(P.stat, stat.data(), stat.size() * D, cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_loop_velref, vel.data(), vel.size() * D, cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
  }
  lpvmpc_args &a = h->loop_args;
  std::memset(&a, 0, sizeof(a));
  a.lap_all = 0; a.Cf_new = c->Cf_new;
  a.x0 = P.x0; a.u_prev = P.u_prev; a.vel_ref = h->d_loop_velref; a.traj = P.traj; a.u_old = P.u_old;
  a.x_pred = h->d_loop_xpred; a.u_pred = P.u_pred; a.status = h->d_loop_status; a.iters = h->d_loop_iters;
  return stream_leave(h, s);
}

int lpvmpc_loop_run_dev(lpvmpc_handle *h, int32_t n_ticks, void *stream) {
  if (!h) return LPVMPC_E_ARG;
  if (!h->d_loop || h->loop_B < 1) return fail(h, LPVMPC_E_ARG, "lpvmpc_loop_init_* first");
  if (n_ticks < 0) return fail(h, LPVMPC_E_ARG, "n_ticks < 0");
  if (n_ticks == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = stream_enter(h, s)) return rc;
  const int B = h->loop_B, grid = (B + 127) / 128;
  for (int t = 0; t < n_ticks; ++t) {
    const int warm = h->loop_tick < (long long)h->LP.lc.warmup_ticks ? 1 : 0;
    lpv::loop::lpv_loop_kernel<<<grid, 128, 0, s>>>(h->LP, t > 0 ? 1 : 0, 1, warm);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
    lpvmpc_args &a = h->loop_args;
    a.sched_mode = warm ? LPVMPC_SCHED_ESTIMATE : LPVMPC_SCHED_PREDICT;
    a.x0_from_prediction = warm ? 0 : 1;
    a.order_hint = h->loop_tick > 0 ? h->d_loop_iters : nullptr;   // last tick's iteration counts group the fleet
    const int rc = lpvmpc_solve_dev(h, B, &a, stream);
    if (rc) return rc;
    ++h->loop_tick;
  }
  lpv::loop::lpv_loop_kernel<<<grid, 128, 0, s>>>(h->LP, 1, 0, 0);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  if (!h->loop_ev) CUDA_TRY(h, cudaEventCreateWithFlags(&h->loop_ev, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventRecord(h->loop_ev, s));
  return stream_leave(h, s);
}

int lpvmpc_loop_init_host(lpvmpc_handle *h, int32_t B, const lpvmpc_loop_cfg *c, const double *sim0) {
  if (!h || !sim0) return fail(h, LPVMPC_E_ARG, "null handle/sim0");
  if (B < 1 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "fleet size must be in [1, max_batch]");
  ON_DEVICE(h);
  const size_t bytes = 8 * sizeof(double) * (size_t)B;
  if (bytes > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow");
  std::memcpy(h->h_stage, sim0, bytes);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, bytes, cudaMemcpyHostToDevice, h->stream));
  const int rc = lpvmpc_loop_init_dev(h, B, c, reinterpret_cast<const double *>(h->d_stage), h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return LPVMPC_OK;
}

int lpvmpc_loop_run_host(lpvmpc_handle *h, int32_t n_ticks) {
  if (!h) return LPVMPC_E_ARG;
  const int rc = lpvmpc_loop_run_dev(h, n_ticks, h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return LPVMPC_OK;
}

// ------------------------------------------------------------------------------------------------
// planner loop
int lpvmpc_plan_loop_init_dev(lpvmpc_handle *h, int32_t B, const double *xstart, const double *s0, double max_ey, double accel_rate,
                              void *stream) {
  if (!h || !xstart) return fail(h, LPVMPC_E_ARG, "null handle/xstart");
  if (h->cfg.kind != LPVMPC_PLANNER) return fail(h, LPVMPC_E_UNSUPPORTED, "the planner loop needs a planner handle");
  if (B < 1 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "fleet size must be in [1, max_batch]");
  ON_DEVICE(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = stream_enter(h, s)) return rc;
  const int N = h->L.N;
  const size_t D = sizeof(double), mb = (size_t)h->cfg.max_batch;
  lpv::loop::PlanLoopParams &P = h->PP;
  if (!h->d_ploop) {
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t r = off; off += align256(bytes * mb); return r; };
    const size_t o_xs = take(5 * D), o_xp = take(5 * (N + 1) * D), o_up = take(2 * N * D), o_ss = take((N + 1) * D),
                 o_ctr = take(8 * sizeof(int32_t)), o_stat = take(4 * D), o_x0 = take(5 * D), o_uprev = take(2 * N * D),
                 o_traj = take(6 * N * D), o_mey = take(D), o_status = take(sizeof(int32_t)), o_iters = take(sizeof(int32_t));
    CUDA_TRY(h, cudaMalloc(&h->d_ploop, off));
    char *base = h->d_ploop;
    P.xstart = (double *)(base + o_xs); P.x_pred = (double *)(base + o_xp); P.u_pred = (double *)(base + o_up);
    P.SS = (double *)(base + o_ss); P.ctr = (int *)(base + o_ctr); P.stat = (double *)(base + o_stat);
    P.x0 = (double *)(base + o_x0); P.u_prev = (double *)(base + o_uprev); P.traj = (double *)(base + o_traj);
    h->d_ploop_maxey = (double *)(base + o_mey);
    P.status = (int *)(base + o_status); P.iters = (int *)(base + o_iters);
  }
  P.dt = h->M.dt; P.accel_rate = accel_rate; P.track = h->d_track; P.nseg = h->M.nseg; P.N = N; P.B = B;
  h->loop_B = B; h->loop_tick = 0;
  CUDA_TRY(h, cudaMemcpyAsync(const_cast<double *>(P.xstart), xstart, 5 * D * (size_t)B, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(h, cudaMemsetAsync(P.x_pred, 0, 5 * (N + 1) * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(P.u_pred, 0, 2 * N * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(P.SS, 0, (N + 1) * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(P.ctr, 0, 8 * sizeof(int32_t) * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(P.stat, 0, 4 * D * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(const_cast<int *>(P.status), 0, sizeof(int32_t) * (size_t)B, s));
  CUDA_TRY(h, cudaMemsetAsync(const_cast<int *>(P.iters), 0, sizeof(int32_t) * (size_t)B, s));
  if (s0) CUDA_TRY(h, cudaMemcpy2DAsync(P.SS, (N + 1) * D, s0, D, D, (size_t)B, cudaMemcpyDeviceToDevice, s));
  {
    std::vector<double> mey((size_t)B, max_ey);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_ploop_maxey, mey.data(), mey.size() * D, cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
  }
  lpvmpc_args &a = h->loop_args;
  std::memset(&a, 0, sizeof(a));
  a.x0 = P.x0; a.x_sched = P.x0; a.u_prev = P.u_prev; a.SS = P.SS; a.traj = P.traj; a.max_ey = h->d_ploop_maxey;
  a.x_pred = P.x_pred; a.u_pred = P.u_pred; a.status = const_cast<int32_t *>(P.status); a.iters = const_cast<int32_t *>(P.iters);
  return stream_leave(h, s);
}

int lpvmpc_plan_loop_run_dev(lpvmpc_handle *h, int32_t n_ticks, void *stream) {
  if (!h) return LPVMPC_E_ARG;
  if (!h->d_ploop || h->loop_B < 1 || h->cfg.kind != LPVMPC_PLANNER) return fail(h, LPVMPC_E_ARG, "lpvmpc_plan_loop_init_* first");
  if (n_ticks < 0) return fail(h, LPVMPC_E_ARG, "n_ticks < 0");
  if (n_ticks == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = stream_enter(h, s)) return rc;
  const int B = h->loop_B, grid = (B + 127) / 128;
  for (int t = 0; t < n_ticks; ++t) {
    const int first = h->loop_tick == 0 ? 1 : 0;
    lpv::loop::lpv_plan_loop_kernel<<<grid, 128, 0, s>>>(h->PP, t > 0 ? 1 : 0, 1, first);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
    lpvmpc_args &a = h->loop_args;
    a.sched_mode = first ? LPVMPC_SCHED_ESTIMATE : LPVMPC_SCHED_PREDICT;
    const int rc = lpvmpc_solve_dev(h, B, &a, stream);
    if (rc) return rc;
    ++h->loop_tick;
  }
  lpv::loop::lpv_plan_loop_kernel<<<grid, 128, 0, s>>>(h->PP, 1, 0, 0);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  if (!h->loop_ev) CUDA_TRY(h, cudaEventCreateWithFlags(&h->loop_ev, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventRecord(h->loop_ev, s));
  return stream_leave(h, s);
}

int lpvmpc_plan_loop_init_host(lpvmpc_handle *h, int32_t B, const double *xstart, const double *s0, double max_ey, double accel_rate) {
  if (!h || !xstart) return fail(h, LPVMPC_E_ARG, "null handle/xstart");
  if (B < 1 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "fleet size must be in [1, max_batch]");
  ON_DEVICE(h);
  const size_t bx = 5 * sizeof(double) * (size_t)B, bs = sizeof(double) * (size_t)B, o2 = align256(bx);
  if (o2 + bs > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow");
  std::memcpy(h->h_stage, xstart, bx);
  if (s0) std::memcpy(h->h_stage + o2, s0, bs);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, o2 + bs, cudaMemcpyHostToDevice, h->stream));
  const int rc = lpvmpc_plan_loop_init_dev(h, B, reinterpret_cast<const double *>(h->d_stage),
                                           s0 ? reinterpret_cast<const double *>(h->d_stage + o2) : nullptr, max_ey, accel_rate, h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return LPVMPC_OK;
}

int lpvmpc_plan_loop_run_host(lpvmpc_handle *h, int32_t n_ticks) {
  if (!h) return LPVMPC_E_ARG;
  const int rc = lpvmpc_plan_loop_run_dev(h, n_ticks, h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return LPVMPC_OK;
}

int lpvmpc_plan_loop_view_dev(lpvmpc_handle *h, lpvmpc_plan_loop_state *view, int32_t *B) {
  if (!h || !view) return LPVMPC_E_ARG;
  if (!h->d_ploop || h->loop_B < 1) return fail(h, LPVMPC_E_ARG, "lpvmpc_plan_loop_init_* first");
  view->x_pred = h->PP.x_pred; view->u_pred = h->PP.u_pred; view->SS = h->PP.SS; view->stat = h->PP.stat; view->ctr = h->PP.ctr;
  if (B) *B = h->loop_B;
  return LPVMPC_OK;
}

int lpvmpc_plan_loop_read_host(lpvmpc_handle *h, const lpvmpc_plan_loop_state *dst) {
  if (!h || !dst) return LPVMPC_E_ARG;
  if (!h->d_ploop || h->loop_B < 1) return fail(h, LPVMPC_E_ARG, "lpvmpc_plan_loop_init_* first");
  ON_DEVICE(h);
  const size_t B = (size_t)h->loop_B, D = sizeof(double), N = (size_t)h->L.N;
  struct Item { void *dst; const void *src; size_t bytes; };
  const Item items[] = {{dst->x_pred, h->PP.x_pred, 5 * (N + 1) * D * B}, {dst->u_pred, h->PP.u_pred, 2 * N * D * B},
                        {dst->SS, h->PP.SS, (N + 1) * D * B}, {dst->stat, h->PP.stat, 4 * D * B}, {dst->ctr, h->PP.ctr, 8 * sizeof(int32_t) * B}};
  if (h->loop_ev) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->loop_ev, 0));
  size_t off = 0;
  for (const Item &it : items) {
    if (!it.dst) continue;
    if (off + it.bytes > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow");
    CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + off, it.src, it.bytes, cudaMemcpyDeviceToHost, h->stream));
    off += align256(it.bytes);
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  off = 0;
  for (const Item &it : items) {
    if (!it.dst) continue;
    std::memcpy(it.dst, h->h_stage + off, it.bytes);
    off += align256(it.bytes);
  }
  return LPVMPC_OK;
}

// ------------------------------------------------------------------------------------------------
// planner -> controller references
int lpvmpc_plan_refs_setup(lpvmpc_handle *h, int32_t n_out, const double *W, const double *Wc) {
  if (!h || !W || !Wc) return fail(h, LPVMPC_E_ARG, "null handle/W/Wc");
  if (h->cfg.kind != LPVMPC_PLANNER) return fail(h, LPVMPC_E_UNSUPPORTED, "references need a planner handle");
  if (n_out < 1 || n_out > 4096) return fail(h, LPVMPC_E_ARG, "n_out must be in [1, 4096]");
  ON_DEVICE(h);
  const size_t bytes = sizeof(double) * (size_t)n_out * (size_t)h->L.N;
  cudaFree(h->d_refs_W); h->d_refs_W = nullptr;
  CUDA_TRY(h, cudaMalloc(&h->d_refs_W, 2 * bytes));
  CUDA_TRY(h, cudaMemcpy(h->d_refs_W, W, bytes, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(reinterpret_cast<char *>(h->d_refs_W) + bytes, Wc, bytes, cudaMemcpyHostToDevice));
  h->refs_n_out = n_out;
  return LPVMPC_OK;
}

int lpvmpc_plan_refs_dev(lpvmpc_handle *h, int32_t B, const double *x_pred, const double *SS, const double *xyth0, double *refs,
                         int32_t *err, void *stream) {
  if (!h || !x_pred || !SS || !xyth0 || !refs) return fail(h, LPVMPC_E_ARG, "null handle/x_pred/SS/xyth0/refs");
  if (!h->d_refs_W) return fail(h, LPVMPC_E_ARG, "lpvmpc_plan_refs_setup first");
  if (B < 0 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "batch exceeds max_batch");
  if (B == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  lpv::loop::PlanRefsParams p;
  p.track = h->d_track; p.nseg = h->M.nseg; p.N = h->L.N; p.n_out = h->refs_n_out; p.B = B;
  p.W = h->d_refs_W; p.Wc = h->d_refs_W + (size_t)h->refs_n_out * (size_t)h->L.N;
  p.x_pred = x_pred; p.SS = SS; p.xyth0 = xyth0; p.refs = refs; p.err = err;
  lpv::loop::lpv_plan_refs_kernel<<<B, 128, sizeof(double) * 5 * (size_t)h->L.N, (cudaStream_t)stream>>>(p);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}

int lpvmpc_plan_refs_host(lpvmpc_handle *h, int32_t B, const double *x_pred, const double *SS, const double *xyth0, double *refs,
                          int32_t *err) {
  if (!h || !x_pred || !SS || !xyth0 || !refs) return fail(h, LPVMPC_E_ARG, "null handle/x_pred/SS/xyth0/refs");
  if (!h->d_refs_W) return fail(h, LPVMPC_E_ARG, "lpvmpc_plan_refs_setup first");
  if (B < 0 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "batch exceeds max_batch");
  if (B == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  const size_t D = sizeof(double), N = (size_t)h->L.N, nb = (size_t)B;
  const size_t b_x = 5 * (N + 1) * D * nb, b_s = (N + 1) * D * nb, b_0 = 3 * D * nb, b_r = 5 * (size_t)h->refs_n_out * D * nb, b_e = sizeof(int32_t) * nb;
  const size_t o_x = 0, o_s = o_x + align256(b_x), o_0 = o_s + align256(b_s), o_r = o_0 + align256(b_0), o_e = o_r + align256(b_r);
  if (o_e + align256(b_e) > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow");
  run_copies({{h->h_stage + o_x, x_pred, b_x}, {h->h_stage + o_s, SS, b_s}, {h->h_stage + o_0, xyth0, b_0}});
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, o_r, cudaMemcpyHostToDevice, h->stream));
  const int rc = lpvmpc_plan_refs_dev(h, B, reinterpret_cast<const double *>(h->d_stage + o_x), reinterpret_cast<const double *>(h->d_stage + o_s),
                                      reinterpret_cast<const double *>(h->d_stage + o_0), reinterpret_cast<double *>(h->d_stage + o_r),
                                      reinterpret_cast<int32_t *>(h->d_stage + o_e), h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + o_r, h->d_stage + o_r, (o_e - o_r) + align256(b_e), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  run_copies({{refs, h->h_stage + o_r, b_r}});
  if (err) std::memcpy(err, h->h_stage + o_e, b_e);
  return LPVMPC_OK;
}

// ------------------------------------------------------------------------------------------------
// controller <- planner hand-off (trajectory-tracking branch of the controller node)
int lpvmpc_track_inputs_dev(lpvmpc_handle *h, int32_t B, const double *gstate, const int32_t *lap, const double *s_prev, const double *refs,
                            int32_t n_ref, const int32_t *index, int32_t index_max, double *x0, double *vel_ref, double *curv_ref, double *ex,
                            void *stream) {
  if (!h || !gstate || !s_prev || !refs || !x0 || !vel_ref || !curv_ref) return fail(h, LPVMPC_E_ARG, "null handle/gstate/s_prev/refs/x0/vel_ref/curv_ref");
  if (h->cfg.kind != LPVMPC_CONTROLLER) return fail(h, LPVMPC_E_UNSUPPORTED, "the hand-off needs a controller handle");
  if (index_max < 0 || n_ref < h->L.N + index_max) return fail(h, LPVMPC_E_ARG, "n_ref must hold N + index_max samples");
  if (B < 0 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "batch exceeds max_batch");
  if (B == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  lpv::loop::TrackInputsParams p;
  p.N = h->L.N; p.n_ref = n_ref; p.B = B; p.dt = h->cfg.dt;
  p.gstate = gstate; p.lap = lap; p.s_prev = s_prev; p.refs = refs; p.index = index;
  p.x0 = x0; p.vel_ref = vel_ref; p.curv_ref = curv_ref; p.ex = ex;
  lpv::loop::lpv_track_inputs_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}

int lpvmpc_track_inputs_host(lpvmpc_handle *h, int32_t B, const double *gstate, const int32_t *lap, const double *s_prev, const double *refs,
                             int32_t n_ref, const int32_t *index, double *x0, double *vel_ref, double *curv_ref, double *ex) {
  if (!h || !gstate || !s_prev || !refs || !x0 || !vel_ref || !curv_ref) return fail(h, LPVMPC_E_ARG, "null handle/gstate/s_prev/refs/x0/vel_ref/curv_ref");
  if (B < 0 || B > h->cfg.max_batch) return fail(h, LPVMPC_E_ARG, "batch exceeds max_batch");
  if (B == 0) return LPVMPC_OK;
  int32_t index_max = 0;
  if (index)
    for (int32_t b = 0; b < B; ++b) {
      if (index[b] < 0) return fail(h, LPVMPC_E_ARG, "negative window index");
      if (index[b] > index_max) index_max = index[b];
    }
  ON_DEVICE(h);
  const size_t D = sizeof(double), I = sizeof(int32_t), N = (size_t)h->L.N, nb = (size_t)B;
  const size_t b_g = 6 * D * nb, b_l = I * nb, b_s = D * nb, b_r = 5 * (size_t)(n_ref > 0 ? n_ref : 0) * D * nb, b_i = I * nb;
  const size_t b_x = 6 * D * nb, b_v = (N + 1) * D * nb, b_c = N * D * nb, b_e = D * nb;
  const size_t o_g = 0, o_l = o_g + align256(b_g), o_s = o_l + align256(b_l), o_r = o_s + align256(b_s), o_i = o_r + align256(b_r);
  const size_t o_x = o_i + align256(b_i), o_v = o_x + align256(b_x), o_c = o_v + align256(b_v), o_e = o_c + align256(b_c);
  if (o_e + align256(b_e) > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow");
  run_copies({{h->h_stage + o_g, gstate, b_g}, {h->h_stage + o_s, s_prev, b_s}, {h->h_stage + o_r, refs, b_r}});
  if (lap) std::memcpy(h->h_stage + o_l, lap, b_l);
  if (index) std::memcpy(h->h_stage + o_i, index, b_i);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, o_x, cudaMemcpyHostToDevice, h->stream));
  const int rc = lpvmpc_track_inputs_dev(h, B, reinterpret_cast<const double *>(h->d_stage + o_g),
                                         lap ? reinterpret_cast<const int32_t *>(h->d_stage + o_l) : nullptr,
                                         reinterpret_cast<const double *>(h->d_stage + o_s), reinterpret_cast<const double *>(h->d_stage + o_r), n_ref,
                                         index ? reinterpret_cast<const int32_t *>(h->d_stage + o_i) : nullptr, index_max,
                                         reinterpret_cast<double *>(h->d_stage + o_x), reinterpret_cast<double *>(h->d_stage + o_v),
                                         reinterpret_cast<double *>(h->d_stage + o_c), reinterpret_cast<double *>(h->d_stage + o_e), h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + o_x, h->d_stage + o_x, (o_e - o_x) + align256(b_e), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  run_copies({{x0, h->h_stage + o_x, b_x}, {vel_ref, h->h_stage + o_v, b_v}, {curv_ref, h->h_stage + o_c, b_c}});
  if (ex) std::memcpy(ex, h->h_stage + o_e, b_e);
  return LPVMPC_OK;
}

// ------------------------------------------------------------------------------------------------
// TS-fuzzy (ANFIS) scheduling blend and the polytopic LPV observer (csrc/lpv_aux.cuh)
int lpvmpc_anfis_abc_dev(lpvmpc_handle *h, int32_t n, const double *sched, const double *A_tab, const double *B_tab, const double *C_tab,
                         const double *bell, double *A, double *B, double *C, void *stream) {
  if (!h || !sched || !A_tab || !B_tab || !C_tab || !bell || !A || !B || !C) return fail(h, LPVMPC_E_ARG, "null handle/sched/tables/outputs");
  if (n < 0) return fail(h, LPVMPC_E_ARG, "negative element count");
  if (n == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  lpv::aux::AnfisParams p;
  p.sched = sched; p.A_tab = A_tab; p.B_tab = B_tab; p.C_tab = C_tab; p.bell = bell; p.A = A; p.B = B; p.C = C; p.n = n;
  lpv::aux::lpv_anfis_abc_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}
int lpvmpc_anfis_abc_host(lpvmpc_handle *h, int32_t n, const double *sched, const double *A_tab, const double *B_tab, const double *C_tab,
                          const double *bell, double *A, double *B, double *C) {
  if (!h || !sched || !A_tab || !B_tab || !C_tab || !bell || !A || !B || !C) return fail(h, LPVMPC_E_ARG, "null handle/sched/tables/outputs");
  if (n < 0) return fail(h, LPVMPC_E_ARG, "negative element count");
  if (n == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  const size_t D = sizeof(double), nb = (size_t)n;
  const size_t b_s = 5 * D * nb, b_t = (96 + 64 + 32 + 30) * D, b_a = 3 * D * nb, b_b = 2 * D * nb, b_c = D * nb;
  const size_t o_s = 0, o_t = o_s + align256(b_s), o_a = o_t + align256(b_t), o_b = o_a + align256(b_a), o_c = o_b + align256(b_b);
  if (o_c + align256(b_c) > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow (more elements than the handle's arena holds)");
  double *t = reinterpret_cast<double *>(h->h_stage + o_t);
  run_copies({{h->h_stage + o_s, sched, b_s}});
  std::memcpy(t, A_tab, 96 * D); std::memcpy(t + 96, B_tab, 64 * D); std::memcpy(t + 160, C_tab, 32 * D); std::memcpy(t + 192, bell, 30 * D);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, o_a, cudaMemcpyHostToDevice, h->stream));
  const double *dt_ = reinterpret_cast<const double *>(h->d_stage + o_t);
  const int rc = lpvmpc_anfis_abc_dev(h, n, reinterpret_cast<const double *>(h->d_stage + o_s), dt_, dt_ + 96, dt_ + 160, dt_ + 192,
                                      reinterpret_cast<double *>(h->d_stage + o_a), reinterpret_cast<double *>(h->d_stage + o_b),
                                      reinterpret_cast<double *>(h->d_stage + o_c), h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + o_a, h->d_stage + o_a, (o_c - o_a) + align256(b_c), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  run_copies({{A, h->h_stage + o_a, b_a}, {B, h->h_stage + o_b, b_b}, {C, h->h_stage + o_c, b_c}});
  return LPVMPC_OK;
}

int lpvmpc_observer_step_dev(lpvmpc_handle *h, int32_t n, double *est, const double *y, const double *u, const double *lim_ls,
                             const double *gains_ls, const double *lim_hs, const double *gains_hs, const double *C_obs, double dt,
                             const int32_t *use_est, int32_t use_all, void *stream) {
  if (!h || !est || !y || !u || !lim_ls || !gains_ls || !lim_hs || !gains_hs || !C_obs) return fail(h, LPVMPC_E_ARG, "null handle/state/measurement/tables");
  if (n < 0 || !(dt > 0)) return fail(h, LPVMPC_E_ARG, "negative element count or dt <= 0");
  if (n == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  lpv::aux::ObserverParams p;
  p.est = est; p.y = y; p.u = u; p.lim_ls = lim_ls; p.gains_ls = gains_ls; p.lim_hs = lim_hs; p.gains_hs = gains_hs; p.C_obs = C_obs;
  p.use_est = use_est; p.use_all = use_all; p.dt = dt; p.n = n;
  lpv::aux::lpv_observer_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return LPVMPC_OK;
}
int lpvmpc_observer_step_host(lpvmpc_handle *h, int32_t n, double *est, const double *y, const double *u, const double *lim_ls,
                              const double *gains_ls, const double *lim_hs, const double *gains_hs, const double *C_obs, double dt,
                              const int32_t *use_est, int32_t use_all) {
  if (!h || !est || !y || !u || !lim_ls || !gains_ls || !lim_hs || !gains_hs || !C_obs) return fail(h, LPVMPC_E_ARG, "null handle/state/measurement/tables");
  if (n < 0 || !(dt > 0)) return fail(h, LPVMPC_E_ARG, "negative element count or dt <= 0");
  if (n == 0) return LPVMPC_OK;
  ON_DEVICE(h);
  const size_t D = sizeof(double), nb = (size_t)n;
  const size_t b_e = 6 * D * nb, b_y = 5 * D * nb, b_u = 2 * D * nb, b_i = sizeof(int32_t) * nb, b_t = (12 + 480 + 12 + 480 + 30) * D;
  const size_t o_e = 0, o_y = o_e + align256(b_e), o_u = o_y + align256(b_y), o_i = o_u + align256(b_u), o_t = o_i + align256(b_i);
  if (o_t + align256(b_t) > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow (more elements than the handle's arena holds)");
  run_copies({{h->h_stage + o_e, est, b_e}, {h->h_stage + o_y, y, b_y}, {h->h_stage + o_u, u, b_u}});
  if (use_est) std::memcpy(h->h_stage + o_i, use_est, b_i);
  double *t = reinterpret_cast<double *>(h->h_stage + o_t);
  std::memcpy(t, lim_ls, 12 * D); std::memcpy(t + 12, gains_ls, 480 * D); std::memcpy(t + 492, lim_hs, 12 * D);
  std::memcpy(t + 504, gains_hs, 480 * D); std::memcpy(t + 984, C_obs, 30 * D);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, o_t + align256(b_t), cudaMemcpyHostToDevice, h->stream));
  const double *dt_ = reinterpret_cast<const double *>(h->d_stage + o_t);
  const int rc = lpvmpc_observer_step_dev(h, n, reinterpret_cast<double *>(h->d_stage + o_e), reinterpret_cast<const double *>(h->d_stage + o_y),
                                          reinterpret_cast<const double *>(h->d_stage + o_u), dt_, dt_ + 12, dt_ + 492, dt_ + 504, dt_ + 984, dt,
                                          use_est ? reinterpret_cast<const int32_t *>(h->d_stage + o_i) : nullptr, use_all, h->stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + o_e, h->d_stage + o_e, b_e, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  run_copies({{est, h->h_stage + o_e, b_e}});
  return LPVMPC_OK;
}

int lpvmpc_loop_view_dev(lpvmpc_handle *h, lpvmpc_loop_state *view, int32_t *B) {
  if (!h || !view) return LPVMPC_E_ARG;
  if (!h->d_loop || h->loop_B < 1) return fail(h, LPVMPC_E_ARG, "lpvmpc_loop_init_* first");
  view->sim = h->LP.sim; view->cmd = h->LP.cmd; view->u_pred = h->LP.u_pred; view->x_pred = h->d_loop_xpred;
  view->local = h->LP.local; view->stat = h->LP.stat; view->ctr = h->LP.ctr;
  if (B) *B = h->loop_B;
  return LPVMPC_OK;
}

int lpvmpc_loop_read_host(lpvmpc_handle *h, const lpvmpc_loop_state *dst) {
  if (!h || !dst) return LPVMPC_E_ARG;
  if (!h->d_loop || h->loop_B < 1) return fail(h, LPVMPC_E_ARG, "lpvmpc_loop_init_* first");
  ON_DEVICE(h);
  const size_t B = (size_t)h->loop_B, D = sizeof(double), N = (size_t)h->L.N;
  struct Item { void *dst; const void *src; size_t bytes; };
  const Item items[] = {{dst->sim, h->LP.sim, 8 * D * B}, {dst->cmd, h->LP.cmd, 2 * D * B}, {dst->u_pred, h->LP.u_pred, 2 * N * D * B},
                        {dst->x_pred, h->d_loop_xpred, 6 * (N + 1) * D * B}, {dst->local, h->LP.local, 6 * D * B},
                        {dst->stat, h->LP.stat, 4 * D * B}, {dst->ctr, h->LP.ctr, 8 * sizeof(int32_t) * B}};
  // through the pinned arena: one D2H per member, one wait (after the last run, whichever stream it was on)
  if (h->loop_ev) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->loop_ev, 0));
  size_t off = 0;
  for (const Item &it : items) {
    if (!it.dst) continue;
    if (off + it.bytes > h->stage_bytes) return fail(h, LPVMPC_E_ARG, "staging overflow");
    CUDA_TRY(h, cudaMemcpyAsync(h->h_stage + off, it.src, it.bytes, cudaMemcpyDeviceToHost, h->stream));
    off += align256(it.bytes);
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  off = 0;
  for (const Item &it : items) {
    if (!it.dst) continue;
    std::memcpy(it.dst, h->h_stage + off, it.bytes);
    off += align256(it.bytes);
  }
  return LPVMPC_OK;
}

int lpvmpc_schedule_host(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, int32_t *sched_err) {
  return run_host(h, B, a, sched_err, false);
}

}  // extern "C"
