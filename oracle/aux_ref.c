/* TEST INFRASTRUCTURE (CPU oracle) -- see aux_ref.h. */
#include "aux_ref.h"

#include <math.h>

/* PathFollowingLPVMPC.py:551-552: inputs = [vx vx vy vy w w steer steer accel accel];
 * Weights = 1 / (1 + |(inputs - c) / a| ** (2 b));  :555-592 weights[i] = product over the five variables, vertex bit
 * order (vx, vy, omega, steer, accel) = bits 4..0;  :597-601 normalise, blend the vertex tables */
double anfis_abc_ref(const double *sched, const double *A_tab, const double *B_tab, const double *C_tab, const double *bell,
                     double *A, double *B) {
  double W[10], w[32], sum = 0.0;
  for (int i = 0; i < 10; ++i) {
    const double in = sched[i / 2];
    W[i] = 1.0 / (1.0 + pow(fabs((in - bell[i * 3 + 2]) / bell[i * 3 + 0]), 2.0 * bell[i * 3 + 1]));
  }
  for (int v = 0; v < 32; ++v) {
    const int b0 = (v >> 4) & 1, b1 = (v >> 3) & 1, b2 = (v >> 2) & 1, b3 = (v >> 1) & 1, b4 = v & 1;
    w[v] = W[0 + b0] * W[2 + b1] * W[4 + b2] * W[6 + b3] * W[8 + b4];
  }
  for (int v = 0; v < 32; ++v) sum += w[v];
  double a[3] = {0, 0, 0}, b[2] = {0, 0}, c = 0.0;
  for (int v = 0; v < 32; ++v) {
    const double nw = w[v] / sum;
    for (int j = 0; j < 3; ++j) a[j] += nw * A_tab[v * 3 + j];
    for (int j = 0; j < 2; ++j) b[j] += nw * B_tab[v * 2 + j];
    c += nw * C_tab[v];
  }
  for (int j = 0; j < 3; ++j) A[j] = a[j];
  for (int j = 0; j < 2; ++j) B[j] = b[j];
  return c;
}

void observer_step_ref(double *est, const double *y, const double *u, const double *lim_ls, const double *gains_ls,
                       const double *lim_hs, const double *gains_hs, const double *C_obs, double dt, int use_est) {
  /* stateEstimator.py:369-378 scheduling variables */
  const double steer = u[0];
  const double Vx = use_est ? est[0] : y[0], Vy = use_est ? est[1] : 0.0, Th = use_est ? est[5] : y[4];
  /* :392-430 Continuous_AB_Comp */
  const double lf = 0.125, lr = 0.125, m = 1.98, I = 0.03, Cf = 60, Cr = 60, mu = 0.05;
  double A[36] = {0}, Bm[12] = {0};
  Bm[0] = -(sin(steer) * Cf) / m; Bm[1] = 1.0;
  Bm[2] = (cos(steer) * Cf) / m;
  Bm[4] = (lf * Cf * cos(steer)) / I;
  A[0] = -mu;
  A[1] = (sin(steer) * Cf) / (m * Vx);
  A[2] = (sin(steer) * Cf * lf) / (m * Vx) + Vy;
  A[7] = -(Cr + Cf * cos(steer)) / (m * Vx);
  A[8] = -(lf * Cf * cos(steer) - lr * Cr) / (m * Vx) - Vx;
  A[13] = -(lf * Cf * cos(steer) - lr * Cr) / (I * Vx);
  A[14] = -(lf * lf * Cf * cos(steer) + lr * lr * Cr) / (I * Vx);
  A[18] = cos(Th); A[19] = -sin(Th);
  A[24] = sin(Th); A[25] = cos(Th);
  A[32] = 1.0;
  /* :433-492 L_Gain_Comp: the polytope by vx, 16 vertices over (vx, vy, steer, theta) */
  const int hs = Vx > lim_ls[0 * 2 + 1];
  const double *lim = hs ? lim_hs : lim_ls, *G = hs ? gains_hs : gains_ls;
  const double Mvx = (lim[1] - Vx) / (lim[1] - lim[0]);
  const double Mvy = (lim[3] - Vy) / (lim[3] - lim[2]);
  const double Mst = (lim[7] - steer) / (lim[7] - lim[6]);
  const double Mth = (lim[11] - Th) / (lim[11] - lim[10]);
  double L[30];
  for (int e = 0; e < 30; ++e) L[e] = 0.0;
  for (int v = 0; v < 16; ++v) {
    const double f0 = (v & 8) ? (1 - Mvx) : Mvx, f1 = (v & 4) ? (1 - Mvy) : Mvy, f2 = (v & 2) ? (1 - Mst) : Mst, f3 = (v & 1) ? (1 - Mth) : Mth;
    const double muv = f0 * f1 * f2 * f3;
    for (int e = 0; e < 30; ++e) L[e] += muv * G[e * 16 + v];
  }
  /* :384-386 est + dt (A + L C) est + dt B u - dt L y */
  double M[36];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) {
      double acc = 0.0;
      for (int k = 0; k < 5; ++k) acc += L[r * 5 + k] * C_obs[k * 6 + c];
      M[r * 6 + c] = A[r * 6 + c] + acc;
    }
  double out[6];
  for (int r = 0; r < 6; ++r) {
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int c = 0; c < 6; ++c) t1 += M[r * 6 + c] * est[c];
    for (int c = 0; c < 2; ++c) t2 += Bm[r * 2 + c] * u[c];
    for (int c = 0; c < 5; ++c) t3 += L[r * 5 + c] * y[c];
    out[r] = est[r] + ((dt * t1 + dt * t2) - dt * t3);
  }
  for (int r = 0; r < 6; ++r) est[r] = out[r];
}
