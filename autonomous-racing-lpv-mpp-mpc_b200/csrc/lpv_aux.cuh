// Two auxiliary routines of the reference next to the hot path, batched, one thread per element:
//   lpv_anfis_abc_kernel   ControllerObject/PathFollowingLPVMPC.py:530-602  ABC_computation_5SV_new: 32-vertex TS-fuzzy (ANFIS)
//                          blend of vertex models A (32x3), B (32x2), C (32) with generalised-bell memberships of
//                          (vx, vy, omega, steer, accel)
//   lpv_observer_kernel    stateEstimator.py:349-492  GS_LPV_Est + Continuous_AB_Comp + L_Gain_Comp: one Euler step of the
//                          polytopic LPV observer, gain interpolated over 16 vertices of one of two polytopes (by vx)
// The vertex / gain tables are the caller's (the reference loads them from .mat files that are not in its repository); a
// CTA keeps them in shared memory.  HBM-bound by construction: 40 B in / 48 B out (blend), 104 B in / 48 B out (observer)
// per element; the arithmetic follows the reference expressions term by term (-fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace lpv {
namespace aux {

struct AnfisParams {
  const double *sched;                       // [n,5] vx vy omega steer accel
  const double *A_tab, *B_tab, *C_tab, *bell;  // [32,3] [32,2] [32] [10,3]
  double *A, *B, *C;                         // [n,3] [n,2] [n]
  int n;
};

__global__ void __launch_bounds__(128) lpv_anfis_abc_kernel(const AnfisParams p) {
  __shared__ double tab[32 * 6 + 30];   // per vertex: A (3) B (2) C (1); then the bell parameters
  for (int e = threadIdx.x; e < 32 * 6 + 30; e += blockDim.x) {
    double v;
    if (e < 192) { const int vtx = e / 6, j = e - vtx * 6; v = (j < 3) ? p.A_tab[vtx * 3 + j] : ((j < 5) ? p.B_tab[vtx * 2 + j - 3] : p.C_tab[vtx]); }
    else v = p.bell[e - 192];
    tab[e] = v;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const double *bell = tab + 192;
  double W[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const double in = p.sched[(size_t)i * 5 + (k >> 1)];
    W[k] = 1.0 / (1.0 + pow(fabs((in - bell[k * 3 + 2]) / bell[k * 3 + 0]), 2.0 * bell[k * 3 + 1]));
  }
  double w[32], sum = 0.0;
#pragma unroll
  for (int v = 0; v < 32; ++v) {
    w[v] = W[0 + ((v >> 4) & 1)] * W[2 + ((v >> 3) & 1)] * W[4 + ((v >> 2) & 1)] * W[6 + ((v >> 1) & 1)] * W[8 + (v & 1)];
    sum += w[v];
  }
  double a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, c = 0;
#pragma unroll
  for (int v = 0; v < 32; ++v) {
    const double nw = w[v] / sum;
    const double *t = tab + v * 6;
    a0 += nw * t[0]; a1 += nw * t[1]; a2 += nw * t[2]; b0 += nw * t[3]; b1 += nw * t[4]; c += nw * t[5];
  }
  p.A[(size_t)i * 3 + 0] = a0; p.A[(size_t)i * 3 + 1] = a1; p.A[(size_t)i * 3 + 2] = a2;
  p.B[(size_t)i * 2 + 0] = b0; p.B[(size_t)i * 2 + 1] = b1;
  p.C[i] = c;
}

struct ObserverParams {
  double *est;                               // [n,6] vx vy omega x y yaw, in place
  const double *y, *u;                       // [n,5] vx omega x y yaw; [n,2] steer accel
  const double *lim_ls, *gains_ls, *lim_hs, *gains_hs, *C_obs;   // [6,2] [6,5,16] [6,2] [6,5,16] [5,6]
  const int32_t *use_est;                    // [n] or NULL: schedule on the estimate (reference: curr_time > 0.02) or the measurement
  int use_all;                               // when use_est is NULL
  double dt;
  int n;
};

__global__ void __launch_bounds__(128) lpv_observer_kernel(const ObserverParams p) {
  __shared__ double G[2][480];   // gains of the low- / high-speed polytope
  __shared__ double lim[2][12];
  __shared__ double Cs[30];
  for (int e = threadIdx.x; e < 960; e += blockDim.x) G[e / 480][e % 480] = (e < 480) ? p.gains_ls[e] : p.gains_hs[e - 480];
  for (int e = threadIdx.x; e < 24; e += blockDim.x) lim[e / 12][e % 12] = (e < 12) ? p.lim_ls[e] : p.lim_hs[e - 12];
  for (int e = threadIdx.x; e < 30; e += blockDim.x) Cs[e] = p.C_obs[e];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  double est[6], y[5];
#pragma unroll
  for (int r = 0; r < 6; ++r) est[r] = p.est[(size_t)i * 6 + r];
#pragma unroll
  for (int r = 0; r < 5; ++r) y[r] = p.y[(size_t)i * 5 + r];
  const double u0 = p.u[(size_t)i * 2], u1 = p.u[(size_t)i * 2 + 1];
  const bool use = p.use_est ? (p.use_est[i] != 0) : (p.use_all != 0);
  const double steer = u0, Vx = use ? est[0] : y[0], Vy = use ? est[1] : 0.0, Th = use ? est[5] : y[4];
  const double lf = 0.125, lr = 0.125, m = 1.98, I = 0.03, Cf = 60, Cr = 60, mu = 0.05;
  const double sd = sin(steer), cd = cos(steer), st = sin(Th), ct = cos(Th);
  const double B11 = -(sd * Cf) / m, B21 = (cd * Cf) / m, B31 = (lf * Cf * cd) / I;
  const double A11 = -mu, A12 = (sd * Cf) / (m * Vx), A13 = (sd * Cf * lf) / (m * Vx) + Vy;
  const double A22 = -(Cr + Cf * cd) / (m * Vx), A23 = -(lf * Cf * cd - lr * Cr) / (m * Vx) - Vx;
  const double A32 = -(lf * Cf * cd - lr * Cr) / (I * Vx), A33 = -(lf * lf * Cf * cd + lr * lr * Cr) / (I * Vx);
  const int hs = Vx > lim[0][1];
  const double *lm = lim[hs], *Gp = G[hs];
  const double Mvx = (lm[1] - Vx) / (lm[1] - lm[0]), Mvy = (lm[3] - Vy) / (lm[3] - lm[2]);
  const double Mst = (lm[7] - steer) / (lm[7] - lm[6]), Mth = (lm[11] - Th) / (lm[11] - lm[10]);
  double muv[16];
#pragma unroll
  for (int v = 0; v < 16; ++v)
    muv[v] = ((v & 8) ? (1 - Mvx) : Mvx) * ((v & 4) ? (1 - Mvy) : Mvy) * ((v & 2) ? (1 - Mst) : Mst) * ((v & 1) ? (1 - Mth) : Mth);
  double out[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double L[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int v = 0; v < 16; ++v) acc += muv[v] * Gp[(r * 5 + k) * 16 + v];
      L[k] = acc;
    }
    // row r of A_obs (stateEstimator.py:417-422)
    double Ar[6] = {0, 0, 0, 0, 0, 0};
    if (r == 0) { Ar[0] = A11; Ar[1] = A12; Ar[2] = A13; }
    if (r == 1) { Ar[1] = A22; Ar[2] = A23; }
    if (r == 2) { Ar[1] = A32; Ar[2] = A33; }
    if (r == 3) { Ar[0] = ct; Ar[1] = -st; }
    if (r == 4) { Ar[0] = st; Ar[1] = ct; }
    if (r == 5) { Ar[2] = 1.0; }
    double t1 = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double lc = 0.0;
#pragma unroll
      for (int k = 0; k < 5; ++k) lc += L[k] * Cs[k * 6 + c];
      t1 += (Ar[c] + lc) * est[c];
    }
    const double br0 = (r == 0) ? B11 : ((r == 1) ? B21 : ((r == 2) ? B31 : 0.0)), br1 = (r == 0) ? 1.0 : 0.0;
    double t2 = 0.0;
    t2 += br0 * u0; t2 += br1 * u1;
    double t3 = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) t3 += L[k] * y[k];
    out[r] = est[r] + ((p.dt * t1 + p.dt * t2) - p.dt * t3);
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) p.est[(size_t)i * 6 + r] = out[r];
}

}  // namespace aux
}  // namespace lpv
