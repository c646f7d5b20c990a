// Closed-loop tick around the controller QP, on the device (BASELINE configs[3]: Monte-Carlo fleet).
//
// One thread per vehicle does everything of the controller main loop that is NOT the QP: apply the last command to
// the simulator's vehicle model, localise the car on the track, lap bookkeeping, and write the inputs of the next
// batched solve (lpv_solve_*_kernel) in the layout lpvmpc_args expects.  Nothing goes back to the host per tick.
//
// Mirrors (paths under /root/reference/workspace/src/barc/src):
//   Simulator.f                   vehicleSimulator.py:164-199 (u = [motor, servo], :337)
//   Map.getLocalPosition          Utilities/trackInitialization.py:283-383, computeAngle :388-410
//   predicted_vectors_generation  controllerMain.py:510-553
//   main loop, lap 0              controllerMain.py:177-192, 252-257, 289-298, 310-331, 381-383
// Compiled with -fmad=false: every expression keeps the reference's operation order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "lpvmpc.h"
#include "lpv_model.cuh"

namespace lpv {
namespace loop {

constexpr double kPi = 3.141592653589793;

enum { C_FIRST_IT = 0, C_LAP, C_HALF, C_STATUS, C_ITERS, C_FAIL, C_FAIL_TICK, C_TICKS, C_COUNT };
enum { S_SOLVED = 0, S_ITERS, S_MAX_EY, S_LAP_TICK, S_COUNT };

struct LoopParams {
  lpvmpc_loop_cfg lc;
  double lf, lr, m, Iz;
  const double *track;
  int nseg;
  int N, B;
  // state
  double *sim;      // [B,8]  x y yaw vx vy psiDot ax ay
  double *cmd;      // [B,2]  delta a
  double *u_pred;   // [B,N,2] written by the solve
  double *local;    // [B,6]  measured state in the controller's slots
  int *ctr;         // [B,8]
  double *stat;     // [B,4]
  const int *status, *iters;  // [B] of the last solve
  // inputs of the next solve
  double *x0;       // [B,6]
  double *u_prev;   // [B,N,2]
  double *traj;     // [B,N,6]
  double *u_old;    // [B,2]
};

__device__ __forceinline__ void sim_f(double *st, double motor, double servo, const LoopParams &p) {
  const double x = st[0], y = st[1], yaw = st[2], vx = st[3], vy = st[4], psiDot = st[5], ax = st[6], ay = st[7];
  const double dt = p.lc.sim_dt;
  double a_F = 0.0, a_R = 0.0;
  if (fabs(vx) > 0.2) {
    a_F = servo - atan((vy + p.lf * psiDot) / fabs(vx));
    a_R = atan((-vy + p.lr * psiDot) / fabs(vx));
  }
  const double FyF = 60 * a_F, FyR = 60 * a_R;
  double sy, cy, ss, cs;
  sincos(yaw, &sy, &cy);
  sincos(servo, &ss, &cs);
  st[0] = x + dt * (cy * vx - sy * vy);
  st[1] = y + dt * (sy * vx + cy * vy);
  st[3] = fabs(vx + dt * (ax + psiDot * vy));
  st[4] = vy + dt * (ay - psiDot * vx);
  st[6] = motor - p.lc.sim_mu * vx - FyF / p.m * ss;
  st[7] = 1.0 / p.m * (FyF * cs + FyR);
  st[2] = yaw + dt * (psiDot);
  st[5] = psiDot + dt * (1.0 / p.Iz * (p.lf * FyF * cs - p.lr * FyR));
}

__device__ __forceinline__ double compute_angle(double p1x, double p1y, double ox, double oy, double p2x, double p2y) {
  const double v1x = p1x - ox, v1y = p1y - oy, v2x = p2x - ox, v2y = p2y - oy;
  const double dot = v1x * v2x + v1y * v2y;
  const double det = v1x * v2y - v1y * v2x;
  return atan2(det, dot);
}

// np.unwrap([a, b])[1]
__device__ __forceinline__ double unwrap_second(double a, double b) {
  const double period = 2 * kPi, hi = kPi, lo = -kPi;
  const double dd = b - a;
  double md = fmod(dd - lo, period);
  if (md != 0.0 && md < 0.0) md += period;
  double ddmod = md + lo;
  if (ddmod == lo && dd > 0) ddmod = hi;
  double corr = ddmod - dd;
  if (fabs(dd) < kPi) corr = 0.0;
  return b + corr;
}

__device__ __forceinline__ double norm2(double dx, double dy) { return sqrt(dx * dx + dy * dy); }
__device__ __forceinline__ double sgn0(double v) { return (double)((v > 0) - (v < 0)); }

// out = [s ey epsi]; returns CompletedFlag
__device__ __forceinline__ int local_position(const double *__restrict__ track, int nseg, double width, double x, double y,
                                              double psi, double *out) {
  int done = 0;
  double s = 0, ey = 0, epsi = 0;
  for (int i = 0; i < nseg && !done; ++i) {
    const double *Pi = track + i * 6;
    const double *Pm = track + ((i == 0) ? (nseg - 1) : (i - 1)) * 6;
    const double xf = Pi[0], yf = Pi[1], xs = Pm[0], ys = Pm[1];
    if (Pi[5] == 0.0) {
      epsi = unwrap_second(Pm[2], psi) - Pm[2];
      if (norm2(xs - x, ys - y) == 0) { s = Pi[3]; ey = 0; done = 1; }
      else if (norm2(xf - x, yf - y) == 0) { s = Pi[3] + Pi[4]; ey = 0; done = 1; }
      else if (fabs(compute_angle(x, y, xs, ys, xf, yf)) <= kPi / 2 && fabs(compute_angle(x, y, xf, yf, xs, ys)) <= kPi / 2) {
        const double v1 = norm2(x - xs, y - ys);
        const double angle = compute_angle(xf, yf, xs, ys, x, y);
        double sa, ca;
        sincos(angle, &sa, &ca);
        s = v1 * ca + Pi[3];
        ey = v1 * sa;
        if (fabs(ey) <= width) done = 1;
      }
    } else {
      const double r = 1 / Pi[5];
      const double direction = (r >= 0) ? 1 : -1;
      const double ang = Pm[2];
      double sc, cc;
      sincos(ang + direction * kPi / 2, &sc, &cc);
      const double CenterX = xs + fabs(r) * cc;
      const double CenterY = ys + fabs(r) * sc;
      if (norm2(xs - x, ys - y) == 0) { ey = 0; epsi = unwrap_second(ang, psi) - ang; s = Pi[3]; done = 1; }
      else if (norm2(xf - x, yf - y) == 0) { s = Pi[3] + Pi[4]; ey = 0; epsi = unwrap_second(Pi[2], psi) - Pi[2]; done = 1; }
      else {
        const double arc1 = Pi[4] * Pi[5];
        const double arc2 = compute_angle(xs, ys, CenterX, CenterY, x, y);
        if (sgn0(arc1) == sgn0(arc2) && fabs(arc1) >= fabs(arc2)) {
          const double vn = norm2(x - CenterX, y - CenterY);
          s = fabs(arc2) * fabs(r) + Pi[3];
          ey = -sgn0(direction) * (vn - fabs(r));
          epsi = unwrap_second(ang + arc2, psi) - (ang + arc2);
          if (fabs(ey) <= width) done = 1;
        }
      }
    }
  }
  if (!done) { s = 10000; ey = 10000; epsi = 10000; }
  out[0] = s; out[1] = ey; out[2] = epsi;
  return done;
}

__constant__ double kGuessDv[20] = {0.05, 0.2, 0.4, 0.6, 0.7, 0.8, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9};
__constant__ double kGuessDs[20] = {0, 0.01, 0.02, 0.04, 0.07, 0.1, 0.14, 0.18, 0.23, 0.55, 0.66, 0.77, 0.89, 1.00, 1.19, 1.39, 1.59, 1.79, 1.89, 1.999};
__constant__ double kGuessUa[20] = {0.0, 0.3, 0.5, 0.7, 0.8, 0.9, 0.9, 0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.30, 0.22, 0.18, 0.14, 0.1, 0.1, 0.1};

__device__ __forceinline__ bool feasible(int status) {
  return status == LPVMPC_SOLVED || status == LPVMPC_SOLVED_INACCURATE || status == LPVMPC_MAX_ITER_REACHED;
}

// do_apply: fold the last solve into the vehicle (status, command, `substeps` simulator steps).
// do_measure: localise, lap logic, write the next solve's inputs (`warmup` selects the _EstimateABC guess path).
__global__ void __launch_bounds__(128) lpv_loop_kernel(const __grid_constant__ LoopParams p, int do_apply, int do_measure,
                                                       int warmup) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int N = p.N;
  int *ct = p.ctr + (size_t)b * C_COUNT;
  double *sv = p.stat + (size_t)b * S_COUNT;
  double *S = p.sim + (size_t)b * 8;
  double *cm = p.cmd + (size_t)b * 2;
  int fail = ct[C_FAIL];
  if (do_apply && !fail) {
    const int status = p.status[b], it = p.iters[b];
    ct[C_STATUS] = status; ct[C_ITERS] = it;
    if (!feasible(status)) { fail = status; ct[C_FAIL] = status; ct[C_FAIL_TICK] = ct[C_TICKS]; }
    else {
      if (status == LPVMPC_SOLVED) sv[S_SOLVED] += 1;
      sv[S_ITERS] += it;
      const double delta = p.u_pred[(size_t)b * N * 2 + 0], acc = p.u_pred[(size_t)b * N * 2 + 1];
      cm[0] = delta; cm[1] = acc;
      double st[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) st[e] = S[e];
      for (int k = 0; k < p.lc.substeps; ++k) sim_f(st, acc, delta, p);
#pragma unroll
      for (int e = 0; e < 8; ++e) S[e] = st[e];
      ct[C_TICKS] += 1;
    }
  }
  if (!do_measure) return;
  double *x0 = p.x0 + (size_t)b * 6;
  if (!fail) {
    double loc[6], lp[3];
    loc[0] = S[3] < 0.01 ? 0.01 : S[3]; loc[1] = S[4]; loc[2] = S[5];
    const int ok = local_position(p.track, p.nseg, p.lc.half_width + p.lc.slack, S[0], S[1], S[2], lp);
    loc[4] = lp[0];
    if (p.lc.swap_ey_epsi) { loc[3] = lp[1]; loc[5] = lp[2]; } else { loc[3] = lp[2]; loc[5] = lp[1]; }
#pragma unroll
    for (int e = 0; e < 6; ++e) p.local[(size_t)b * 6 + e] = loc[e];
    if (!ok) { fail = LPVMPC_OFF_TRACK; ct[C_FAIL] = fail; ct[C_FAIL_TICK] = ct[C_TICKS]; ct[C_STATUS] = fail; }
    else {
      if (fabs(lp[1]) > sv[S_MAX_EY]) sv[S_MAX_EY] = fabs(lp[1]);
      const double TrackLength = p.track[(p.nseg - 1) * 6 + 3] + p.track[(p.nseg - 1) * 6 + 4];
      if (loc[4] >= 3 * TrackLength / 4) ct[C_HALF] = 1;
      if (ct[C_HALF] == 1 && loc[4] <= TrackLength / 4) {
        ct[C_HALF] = 0; ct[C_LAP] += 1;
        if (sv[S_LAP_TICK] < 0) sv[S_LAP_TICK] = ct[C_TICKS];
      }
      p.u_old[(size_t)b * 2 + 0] = cm[0]; p.u_old[(size_t)b * 2 + 1] = cm[1];
#pragma unroll
      for (int e = 0; e < 6; ++e) x0[e] = loc[e];
      double *up = p.u_prev + (size_t)b * N * 2;
      if (warmup) {
        double *tr = p.traj + (size_t)b * N * 6;
        for (int i = 0; i < N; ++i) {
          tr[i * 6 + 0] = loc[0] + kGuessDv[i]; tr[i * 6 + 1] = loc[1]; tr[i * 6 + 2] = loc[2];
          tr[i * 6 + 3] = 0.0001; tr[i * 6 + 4] = loc[4] + kGuessDs[i]; tr[i * 6 + 5] = 0.0001;
          up[i * 2 + 0] = 0.; up[i * 2 + 1] = kGuessUa[i];
        }
        ct[C_FIRST_IT] += 1;
      } else {
        const double *us = p.u_pred + (size_t)b * N * 2;
        for (int i = 0; i < 2 * N; ++i) up[i] = us[i];
      }
    }
  }
  if (fail) {  // retired vehicle: a NaN arc length makes the solve kernel leave at once with LPVMPC_SCHEDULE_ERROR
    const double qnan = nan("");
#pragma unroll
    for (int e = 0; e < 6; ++e) x0[e] = qnan;
    if (warmup) for (int i = 0; i < N; ++i) p.traj[((size_t)b * N + i) * 6 + 4] = qnan;
  }
}

// ------------------------------------------------------------------------------------------------
// Planner loop (plannerMain.py:128-224): re-plan from the previous plan's second state, integrate the arc lengths.
enum { PC_TICKS = 0, PC_STATUS, PC_ITERS, PC_FAIL, PC_FAIL_TICK, PC_COUNT = 8 };

struct PlanLoopParams {
  double dt, accel_rate;
  const double *track;
  int nseg;
  int N, B;
  const double *xstart;   // [B,5]   state of the first tick
  double *x_pred;         // [B,N+1,5] written by the solve
  double *u_pred;         // [B,N,2]
  double *SS;             // [B,N+1]
  int *ctr;               // [B,8]
  double *stat;           // [B,4]
  const int *status, *iters;
  // inputs of the next solve
  double *x0;             // [B,5]
  double *u_prev;         // [B,N,2]
  double *traj;           // [B,N,6]  warm-up tick: predicted_vectors_generation (plannerMain.py:465-505)
};

// do_apply: fold the last solve in (status, arc-length integration over the new plan, SS[0] = SS[1]).
// do_prepare: inputs of the next solve (first: the guess around xstart for _EstimateABC; later: x0 = xPred[1], uPred).
__global__ void __launch_bounds__(128) lpv_plan_loop_kernel(const __grid_constant__ PlanLoopParams p, int do_apply, int do_prepare,
                                                            int first) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int N = p.N;
  int *ct = p.ctr + (size_t)b * PC_COUNT;
  double *sv = p.stat + (size_t)b * 4;
  double *ss = p.SS + (size_t)b * (N + 1);
  const double *xP = p.x_pred + (size_t)b * (N + 1) * 5;
  int fail = ct[PC_FAIL];
  if (do_apply && !fail) {
    const int status = p.status[b], it = p.iters[b];
    ct[PC_STATUS] = status; ct[PC_ITERS] = it;
    if (!feasible(status)) { fail = status; ct[PC_FAIL] = status; ct[PC_FAIL_TICK] = ct[PC_TICKS]; }
    else {
      if (status == LPVMPC_SOLVED) sv[0] += 1;
      sv[1] += it;
      int err = 0;
      double s = ss[0];
      for (int j = 0; j < N; ++j) {   // plannerMain.py:201-208
        const double curv = curvature(p.track, p.nseg, s, err);
        const double *x = xP + j * 5;
        double se, ce;
        sincos(x[4], &se, &ce);
        s = (s + ((x[0] * ce - x[1] * se) / (1 - x[3] * curv)) * p.dt);
        ss[j + 1] = s;
      }
      ss[0] = ss[1];                  // :211
      ct[PC_TICKS] += 1;
      if (err) { fail = LPVMPC_SCHEDULE_ERROR; ct[PC_FAIL] = fail; ct[PC_FAIL_TICK] = ct[PC_TICKS]; ct[PC_STATUS] = fail; }
    }
  }
  if (!do_prepare) return;
  double *x0 = p.x0 + (size_t)b * 5;
  if (!fail) {
    if (first) {
      const double *xs = p.xstart + (size_t)b * 5;
      double *tr = p.traj + (size_t)b * N * 6, *up = p.u_prev + (size_t)b * N * 2;
      double Vx = xs[0], S = ss[0];
      double se, ce;
      sincos(xs[4], &se, &ce);
      for (int i = 0; i < N; ++i) {
        tr[i * 6 + 0] = Vx; tr[i * 6 + 1] = xs[1]; tr[i * 6 + 2] = xs[2]; tr[i * 6 + 3] = xs[3]; tr[i * 6 + 4] = xs[4]; tr[i * 6 + 5] = S;
        up[i * 2 + 0] = 0.0; up[i * 2 + 1] = 0.0;
        const double Accel = 0.1 + p.accel_rate * i;
        const double Vn = Vx + Accel * p.dt;
        S = S + ((Vx * ce - xs[1] * se) / (1 - xs[3] * 0.0)) * p.dt;
        Vx = Vn;
      }
#pragma unroll
      for (int e = 0; e < 5; ++e) x0[e] = xs[e];
    } else {
#pragma unroll
      for (int e = 0; e < 5; ++e) x0[e] = xP[5 + e];
      const double *us = p.u_pred + (size_t)b * N * 2;
      double *up = p.u_prev + (size_t)b * N * 2;
      for (int i = 0; i < 2 * N; ++i) up[i] = us[i];
    }
  } else {  // retired plan: a NaN arc length makes the solve leave at once with LPVMPC_SCHEDULE_ERROR
    const double qnan = nan("");
    p.SS[(size_t)b * (N + 1)] = qnan;
    if (first) for (int i = 0; i < N; ++i) p.traj[((size_t)b * N + i) * 6 + 5] = qnan;
#pragma unroll
    for (int e = 0; e < 5; ++e) x0[e] = qnan;
  }
}

// ------------------------------------------------------------------------------------------------
// Planner -> controller references (plannerMain.py:196-224, 257-280): global positions along the plan, vehicle pose and
// curvature per stage, then the cubic resampling to the controller's period and the zero-phase filter of the curvature.
// For the uniform grids involved both are linear maps, built once on the host (postproc.py): W for x, y, yaw, vx and
// Wc = F W for the curvature, [n_out, N] row-major.

__device__ __forceinline__ double wrap_angle(double a) {   // trackInitialization.py:413-421
  if (a < -kPi) return 2 * kPi + a;
  if (a > kPi) return a - 2 * kPi;
  return a;
}

// Map.getGlobalPosition(s, 0) (trackInitialization.py:205-260); err when no unique segment holds s
__device__ __forceinline__ void global_position0(const double *__restrict__ track, int nseg, double s, double *out, int &err) {
  const double TrackLength = track[(nseg - 1) * 6 + 3] + track[(nseg - 1) * 6 + 4];
  if (!(s <= kMaxLaps * TrackLength)) { err = 1; out[0] = out[1] = out[2] = nan(""); return; }   // NaN / runaway arc length (lpv_model.cuh)
  while (s > TrackLength) s = s - TrackLength;
  int i = -1, cnt = 0;
  for (int k = 0; k < nseg; ++k)
    if (s >= track[k * 6 + 3] && s < track[k * 6 + 3] + track[k * 6 + 4]) { if (i < 0) i = k; ++cnt; }
  if (cnt != 1) { err = 1; out[0] = out[1] = out[2] = nan(""); return; }
  const double *Pi = track + i * 6;
  const double *Pm = track + ((i == 0) ? (nseg - 1) : (i - 1)) * 6;
  const double ey = 0.0;
  if (Pi[5] == 0.0) {
    const double xf = Pi[0], yf = Pi[1], xs = Pm[0], ys = Pm[1], psi = Pi[2];
    const double deltaL = Pi[4], reltaL = s - Pi[3];
    out[0] = (1 - reltaL / deltaL) * xs + reltaL / deltaL * xf + ey * cos(psi + kPi / 2);
    out[1] = (1 - reltaL / deltaL) * ys + reltaL / deltaL * yf + ey * sin(psi + kPi / 2);
    out[2] = psi;
  } else {
    const double r = 1 / Pi[5], ang = Pm[2];
    const double direction = (r >= 0) ? 1 : -1;
    const double CenterX = Pm[0] + fabs(r) * cos(ang + direction * kPi / 2);
    const double CenterY = Pm[1] + fabs(r) * sin(ang + direction * kPi / 2);
    const double spanAng = (s - Pi[3]) / (kPi * fabs(r)) * kPi;
    const double angleNormal = wrap_angle(direction * kPi / 2 + ang);
    const double angle = -(kPi - fabs(angleNormal)) * ((angleNormal >= 0) ? 1 : -1);
    out[0] = CenterX + (fabs(r) - direction * ey) * cos(angle + direction * spanAng);
    out[1] = CenterY + (fabs(r) - direction * ey) * sin(angle + direction * spanAng);
    out[2] = ang + direction * spanAng;
  }
}

struct PlanRefsParams {
  const double *track;
  int nseg, N, n_out, B;
  const double *W, *Wc;       // [n_out, N]
  const double *x_pred;       // [B,N+1,5]
  const double *SS;           // [B,N+1]   arc lengths of the plan (entry 0 is not used: the pose of stage 0 is xyth0)
  const double *xyth0;        // [B,3]     Xlast, Ylast, Thetalast (plannerMain.py:196-198)
  double *refs;               // [B,5,n_out]  x_d, y_d, psi_d, vx_d, curv_d of My_Planning (plannerMain.py:299-303)
  int *err;                   // [B] optional: 1 when getGlobalPosition failed
};

// one CTA per plan: stage poses into shared memory, then one output sample per thread and signal
__global__ void __launch_bounds__(128) lpv_plan_refs_kernel(const __grid_constant__ PlanRefsParams p) {
  extern __shared__ double raw[];   // [5][N]: xp, yp, yaw, vel, curv
  const int b = blockIdx.x, N = p.N;
  __shared__ int s_err;
  if (threadIdx.x == 0) s_err = 0;
  __syncthreads();
  const double *xP = p.x_pred + (size_t)b * (N + 1) * 5;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double g[3];
    int err = 0;
    if (i == 0) { g[0] = p.xyth0[(size_t)b * 3]; g[1] = p.xyth0[(size_t)b * 3 + 1]; g[2] = p.xyth0[(size_t)b * 3 + 2]; }
    else global_position0(p.track, p.nseg, p.SS[(size_t)b * (N + 1) + i], g, err);
    if (err) s_err = 1;
    const double *x = xP + i * 5;
    const double yaw = g[2] + x[4];                         // :219
    raw[2 * N + i] = yaw;
    raw[0 * N + i] = g[0] - x[3] * sin(yaw);                // :220
    raw[1 * N + i] = g[1] + x[3] * cos(yaw);                // :221
    raw[3 * N + i] = x[0];                                  // :223
    raw[4 * N + i] = x[2] / x[0];                           // :224
  }
  __syncthreads();
  for (int r = threadIdx.x; r < p.n_out; r += blockDim.x) {
    const double *w = p.W + (size_t)r * N, *wc = p.Wc + (size_t)r * N;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    for (int i = 0; i < N; ++i) {
      const double wi = w[i];
      a0 = fma(wi, raw[i], a0); a1 = fma(wi, raw[N + i], a1); a2 = fma(wi, raw[2 * N + i], a2); a3 = fma(wi, raw[3 * N + i], a3);
      a4 = fma(wc[i], raw[4 * N + i], a4);
    }
    double *o = p.refs + (size_t)b * 5 * p.n_out + r;
    o[0] = a0; o[p.n_out] = a1; o[2 * p.n_out] = a2; o[3 * p.n_out] = a3; o[4 * p.n_out] = a4;
  }
  if (p.err && threadIdx.x == 0) p.err[b] = s_err;
}

// ------------------------------------------------------------------------------------------------
// Controller <- planner hand-off: the trajectory-tracking branch of the controller node,
// controllerMain.py:196-243 with Body_Frame_Errors (:495-506).
struct TrackInputsParams {
  int N, n_ref, B;
  double dt;
  const double *gstate;   // [B,6]  vx vy wz X Y psi (GlobalState; LocalState[0:3] = GlobalState[0:3], :190-192)
  const int *lap;         // [B] optional LapNumber (psi -= 2 pi LapNumber, :200)
  const double *s_prev;   // [B]   SS of the previous tick (:243)
  const double *refs;     // [B,5,n_ref]  x_d, y_d, psi_d, vx_d, curv_d (My_Planning; lpvmpc_plan_refs_*)
  const int *index;       // [B] optional window offset (:229-233), NULL = 0
  double *x0;             // [B,6]   LocalState [vx vy wz epsi s ey]
  double *vel_ref;        // [B,N+1] window of vx_d; entry N repeats entry N-1 (the reference's vel_ref[-1])
  double *curv_ref;       // [B,N]
  double *ex;             // [B] optional longitudinal error (Xerror)
};

__global__ void __launch_bounds__(128) lpv_track_inputs_kernel(const __grid_constant__ TrackInputsParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int N = p.N, o = p.index ? p.index[b] : 0;
  const double *g = p.gstate + (size_t)b * 6;
  const double *r = p.refs + (size_t)b * 5 * p.n_ref + o;
  const double xd = r[0], yd = r[p.n_ref], psid = r[2 * (size_t)p.n_ref], curv = r[4 * (size_t)p.n_ref];
  const double vx = g[0], vy = g[1], x = g[3], y = g[4];
  const double psi = wrap_angle(g[5] - 2 * kPi * (p.lap ? p.lap[b] : 0));        // :200-202
  const double c = cos(psid), s = sin(psid);
  const double exv = (x - xd) * c + (y - yd) * s;                                // :497
  const double ey = -(x - xd) * s + (y - yd) * c;                                // :499
  const double epsi = wrap_angle(psi - psid);                                    // :501
  const double sn = p.s_prev[b] + ((vx * cos(epsi) - vy * sin(epsi)) / (1 - ey * curv)) * p.dt;   // :504
  double *x0 = p.x0 + (size_t)b * 6;
  x0[0] = vx; x0[1] = vy; x0[2] = g[2]; x0[3] = epsi; x0[4] = sn; x0[5] = ey;    // :238-240
  if (p.ex) p.ex[b] = exv;
  const double *v = r + 3 * (size_t)p.n_ref, *k = r + 4 * (size_t)p.n_ref;
  for (int i = 0; i < N; ++i) { p.vel_ref[(size_t)b * (N + 1) + i] = v[i]; p.curv_ref[(size_t)b * N + i] = k[i]; }   // :232-233
  p.vel_ref[(size_t)b * (N + 1) + N] = v[N - 1];
}

}  // namespace loop
}  // namespace lpv
