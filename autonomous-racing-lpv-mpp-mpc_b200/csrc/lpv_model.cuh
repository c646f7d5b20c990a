// LPV bicycle-model scheduling on the device: Curvature lookup, the per-stage A(rho_k), B(rho_k)
// of the controller and the planner, and the serial roll-outs of LPVPrediction.
//
// Mirrors (paths under /root/reference/workspace/src/barc/src):
//   Curvature                    Utilities/utilities.py:31-50
//   controller stage matrices    ControllerObject/PathFollowingLPVMPC.py:203-246 (LPVPrediction), :761-802 (_EstimateABC)
//   planner stage matrices       PlannerObject/LPV_MPC_Planner.py:275-308 (LPVPrediction), :551-584 (_EstimateABC)
// Operation order follows the reference expressions term by term (the file is compiled with -fmad=false, so
// nothing is contracted behind our back); sin/cos are CUDA's fp64 libm (<= 2 ulp) instead of numpy's.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lpv {

struct Model {
  double dt;
  double lf, lr, m, Iz, Cf, Cr, mu;
  double max_vel, min_vel;
  double Q[36], R[4], dR[2], L_cf[5];
  const double *track;  // device, nseg x 6
  int nseg;
  int delay;
};

// Laps the arc-length reduction `while s > TrackLength: s -= TrackLength` (utilities.py:40-41) may take on the device.
// The reference's loop is unbounded (and never ends once s - TrackLength == s); one such element would stall a whole
// batch kernel, so beyond this many laps the element is a schedule error instead.  Below it the subtract loop is kept
// as written: bit parity with the reference / oracle.
constexpr double kMaxLaps = 1.0e4;

// utilities.py:36-48.  Returns NaN and sets err when no unique segment holds s (the reference raises), when s is not
// finite, or when it is more than kMaxLaps track lengths (lpvmpc_create guarantees 0 < TrackLength < inf).
__device__ __forceinline__ double curvature(const double *__restrict__ track, int nseg, double s, int &err) {
  const double track_len = track[(nseg - 1) * 6 + 3] + track[(nseg - 1) * 6 + 4];
  if (!(s <= kMaxLaps * track_len)) { err = 1; return nan(""); }   // also NaN
  while (s > track_len) s = s - track_len;
  int found = -1, cnt = 0;
  for (int i = 0; i < nseg; ++i) {
    const double s0 = track[i * 6 + 3];
    if (s >= s0 && s < s0 + track[i * 6 + 4]) { if (found < 0) found = i; ++cnt; }
  }
  if (cnt != 1) { err = 1; return nan(""); }
  return track[found * 6 + 5];
}

// Controller: states [vx vy wz epsi s ey].  Writes A (6x6, ld 6) and B (6x2) in discrete (Euler) form.
__device__ __forceinline__ void ctrl_stage(const Model &M, double Cf, double Cr, double vx, double vy, double epsi,
                                           double ey, double cur, double delta, double *Ai, double *Bi) {
  const double lf = M.lf, lr = M.lr, m = M.m, I = M.Iz, mu = M.mu, dt = M.dt;
  double sd, cd, se, ce;
  sincos(delta, &sd, &cd);
  sincos(epsi, &se, &ce);
  const double A11 = -mu;
  const double A12 = (sd * Cf) / (m * vx);
  const double A13 = (sd * Cf * lf) / (m * vx) + vy;
  const double A22 = -(Cr + Cf * cd) / (m * vx);
  const double A23 = -(lf * Cf * cd - lr * Cr) / (m * vx) - vx;
  const double A32 = -(lf * Cf * cd - lr * Cr) / (I * vx);
  const double A33 = -(lf * lf * Cf * cd + lr * lr * Cr) / (I * vx);
  const double A51 = (1 / (1 - ey * cur)) * (-ce * cur);
  const double A52 = (1 / (1 - ey * cur)) * (+se * cur);
  const double A61 = ce / (1 - ey * cur);
  const double A62 = se / (1 - ey * cur);
  const double A7 = se;
  const double A8 = ce;
  const double B11 = -(sd * Cf) / m;
  const double B21 = (cd * Cf) / m;
  const double B31 = (lf * Cf * cd) / I;
#pragma unroll
  for (int e = 0; e < 36; ++e) Ai[e] = ((e % 7) == 0) ? 1.0 : 0.0;
  Ai[0] = 1.0 + dt * A11; Ai[1] = 0.0 + dt * A12; Ai[2] = 0.0 + dt * A13;
  Ai[7] = 1.0 + dt * A22; Ai[8] = 0.0 + dt * A23;
  Ai[13] = 0.0 + dt * A32; Ai[14] = 1.0 + dt * A33;
  Ai[18] = 0.0 + dt * A51; Ai[19] = 0.0 + dt * A52; Ai[20] = 0.0 + dt * 1.0; Ai[21] = 1.0 + dt * 0.0;
  Ai[24] = 0.0 + dt * A61; Ai[25] = 0.0 + dt * A62; Ai[28] = 1.0 + dt * 0.0;
  Ai[30] = 0.0 + dt * A7; Ai[31] = 0.0 + dt * A8; Ai[35] = 1.0 + dt * 0.0;
#pragma unroll
  for (int e = 0; e < 12; ++e) Bi[e] = dt * 0.0;
  Bi[0] = dt * B11; Bi[1] = dt * 1.0; Bi[2] = dt * B21; Bi[4] = dt * B31;
}

// Planner: states [vx vy wz ey epsi].  A (5x5), B (5x2).
__device__ __forceinline__ void plan_stage(const Model &M, double vx, double vy, double ey, double epsi, double cur,
                                           double delta, double *Ai, double *Bi) {
  const double lf = M.lf, lr = M.lr, m = M.m, I = M.Iz, mu = M.mu, Cf = M.Cf, Cr = M.Cr, dt = M.dt;
  double sd, cd;
  sincos(delta, &sd, &cd);
  const double A5 = (sd * Cf) / (m * vx);
  const double A6 = (sd * Cf * lf) / (m * vx) + vy;
  const double A7 = -(Cr + Cf * cd) / (m * vx);
  const double A8 = -(lf * Cf * cd - lr * Cr) / (m * vx) - vx;
  const double A9 = -(lf * Cf * cd - lr * Cr) / (I * vx);
  const double A10 = -(lf * lf * Cf * cd + lr * lr * Cr) / (I * vx);
  const double A1 = (1 / (1 - ey * cur));
  const double A2 = sin(epsi);
  const double A4 = vx;
  const double B11 = -(sd * Cf) / m;
  const double B21 = (cd * Cf) / m;
  const double B31 = (lf * Cf * cd) / I;
#pragma unroll
  for (int e = 0; e < 25; ++e) Ai[e] = ((e % 6) == 0) ? 1.0 : 0.0;
  Ai[0] = 1.0 + dt * (-mu); Ai[1] = 0.0 + dt * A5; Ai[2] = 0.0 + dt * A6;
  Ai[6] = 1.0 + dt * A7; Ai[7] = 0.0 + dt * A8;
  Ai[11] = 0.0 + dt * A9; Ai[12] = 1.0 + dt * A10;
  Ai[16] = 0.0 + dt * 1.0; Ai[18] = 1.0 + dt * 0.0; Ai[19] = 0.0 + dt * A4;
  Ai[20] = 0.0 + dt * (-A1 * cur); Ai[21] = 0.0 + dt * (A1 * A2 * cur); Ai[22] = 0.0 + dt * 1.0; Ai[24] = 1.0 + dt * 0.0;
#pragma unroll
  for (int e = 0; e < 10; ++e) Bi[e] = dt * 0.0;
  Bi[0] = dt * B11; Bi[1] = dt * 1.0; Bi[2] = dt * B21; Bi[4] = dt * B31;
}

// x <- A x + B u, summed like the oracle's plain loops (np.dot semantics up to BLAS ordering)
template <int NX>
__device__ __forceinline__ void propagate(const double *Ai, const double *Bi, const double *u, double *st) {
  double nw[NX];
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < NX; ++k) a += Ai[r * NX + k] * st[k];
    double b = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) b += Bi[r * 2 + k] * u[k];
    nw[r] = a + b;
  }
#pragma unroll
  for (int r = 0; r < NX; ++r) st[r] = nw[r];
}

}  // namespace lpv
