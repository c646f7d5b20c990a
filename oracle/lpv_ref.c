/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference's LPV scheduling and QP assembly.
 *
 * Follows (file:line under /root/reference/workspace/src/barc/src):
 *   Curvature                     Utilities/utilities.py:31-50
 *   controller LPVPrediction      ControllerObject/PathFollowingLPVMPC.py:166-258
 *   controller _EstimateABC       ControllerObject/PathFollowingLPVMPC.py:732-809
 *   controller _buildMatIneqConst ControllerObject/PathFollowingLPVMPC.py:329-378
 *   controller _buildMatCost      ControllerObject/PathFollowingLPVMPC.py:382-473
 *   controller _buildMatEqConst   ControllerObject/PathFollowingLPVMPC.py:477-529
 *   controller solve/osqp_solve_qp ControllerObject/PathFollowingLPVMPC.py:89-162,273-325
 *   planner LPVPrediction         PlannerObject/LPV_MPC_Planner.py:242-320
 *   planner _EstimateABC          PlannerObject/LPV_MPC_Planner.py:519-591
 *   planner solve (cost, bounds)  PlannerObject/LPV_MPC_Planner.py:86-236
 *   planner _buildMatEqConst      PlannerObject/LPV_MPC_Planner.py:434-486
 *
 * Pinned against the reference's own Python (imported headless in the build container) through the
 * committed fixtures in tests/golden/ (generator: tests/golden/make_golden.py).
 * The product path never links or calls this file.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared lpv_ref.c osqp_ref.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "lpv_ref.h"

#define NC 6 /* controller states [vx vy wz epsi s ey] */
#define NP 5 /* planner states    [vx vy wz ey epsi]   */
#define ND 2 /* inputs            [delta a]            */

double lpv_ref_curvature(double s, const double *track, int nseg, int *err) {
  /* utilities.py:36-40 : lap wrap */
  double TrackLength = track[(nseg - 1) * 6 + 3] + track[(nseg - 1) * 6 + 4];
  while (s > TrackLength) s = s - TrackLength;
  /* utilities.py:44-48 : the unique segment with s_i <= s < s_i + len_i */
  int found = -1, cnt = 0;
  for (int i = 0; i < nseg; i++) {
    if (s >= track[i * 6 + 3] && s < track[i * 6 + 3] + track[i * 6 + 4]) {
      if (found < 0) found = i;
      cnt++;
    }
  }
  if (cnt != 1) { if (err) *err = 1; return NAN; } /* int(np.where(...)[0]) raises */
  return track[found * 6 + 5];
}

/* continuous-time entries shared by all four schedulers (PathFollowingLPVMPC.py:203-218) */
static void ctrl_stage(const lpv_ref_vehicle *v, double Cf, double Cr, double dt, double vx, double vy, double epsi,
                       double ey, double cur, double delta, double *Ai, double *Bi) {
  double lf = v->lf, lr = v->lr, m = v->m, I = v->Iz, mu = v->mu;
  double A11 = -mu;
  double A12 = (sin(delta) * Cf) / (m * vx);
  double A13 = (sin(delta) * Cf * lf) / (m * vx) + vy;
  double A22 = -(Cr + Cf * cos(delta)) / (m * vx);
  double A23 = -(lf * Cf * cos(delta) - lr * Cr) / (m * vx) - vx;
  double A32 = -(lf * Cf * cos(delta) - lr * Cr) / (I * vx);
  double A33 = -(lf * lf * Cf * cos(delta) + lr * lr * Cr) / (I * vx);
  double A51 = (1 / (1 - ey * cur)) * (-cos(epsi) * cur);
  double A52 = (1 / (1 - ey * cur)) * (+sin(epsi) * cur);
  double A61 = cos(epsi) / (1 - ey * cur);
  double A62 = sin(epsi) / (1 - ey * cur);
  double A7 = sin(epsi);
  double A8 = cos(epsi);
  double B11 = -(sin(delta) * Cf) / m;
  double B21 = (cos(delta) * Cf) / m;
  double B31 = (lf * Cf * cos(delta)) / I;
  double Ac[36] = {A11, A12, A13, 0., 0., 0.,
                   0.,  A22, A23, 0., 0., 0.,
                   0.,  A32, A33, 0., 0., 0.,
                   A51, A52, 1.,  0., 0., 0.,
                   A61, A62, 0.,  0., 0., 0.,
                   A7,  A8,  0.,  0., 0., 0.};
  double Bc[12] = {B11, 1, B21, 0, B31, 0, 0, 0, 0, 0, 0, 0};
  for (int r = 0; r < NC; r++)
    for (int c = 0; c < NC; c++) Ai[r * NC + c] = (r == c ? 1.0 : 0.0) + dt * Ac[r * NC + c];
  for (int k = 0; k < NC * ND; k++) Bi[k] = dt * Bc[k];
}

static void plan_stage(const lpv_ref_vehicle *v, double dt, double vx, double vy, double ey, double epsi, double cur,
                       double delta, double *Ai, double *Bi) {
  /* LPV_MPC_Planner.py:275-308 */
  double lf = v->lf, lr = v->lr, m = v->m, I = v->Iz, mu = v->mu, Cf = v->Cf, Cr = v->Cr;
  double A5 = (sin(delta) * Cf) / (m * vx);
  double A6 = (sin(delta) * Cf * lf) / (m * vx) + vy;
  double A7 = -(Cr + Cf * cos(delta)) / (m * vx);
  double A8 = -(lf * Cf * cos(delta) - lr * Cr) / (m * vx) - vx;
  double A9 = -(lf * Cf * cos(delta) - lr * Cr) / (I * vx);
  double A10 = -(lf * lf * Cf * cos(delta) + lr * lr * Cr) / (I * vx);
  double A1 = (1 / (1 - ey * cur));
  double A2 = sin(epsi);
  double A4 = vx;
  double B11 = -(sin(delta) * Cf) / m;
  double B21 = (cos(delta) * Cf) / m;
  double B31 = (lf * Cf * cos(delta)) / I;
  double Ac[25] = {-mu,       A5,            A6,  0., 0.,
                   0.,        A7,            A8,  0., 0.,
                   0.,        A9,            A10, 0., 0.,
                   0.,        1.,            0.,  0., A4,
                   -A1 * cur, A1 * A2 * cur, 1.,  0., 0.};
  double Bc[10] = {B11, 1, B21, 0, B31, 0, 0, 0, 0, 0};
  for (int r = 0; r < NP; r++)
    for (int c = 0; c < NP; c++) Ai[r * NP + c] = (r == c ? 1.0 : 0.0) + dt * Ac[r * NP + c];
  for (int k = 0; k < NP * ND; k++) Bi[k] = dt * Bc[k];
}

int lpv_ref_ctrl_predict(const lpv_ref_cfg *c, const double *x, const double *u, const double *vel_ref,
                         const double *curv_ref, double Cf_new, int lap, double *states_out, double *A, double *B,
                         double *C) {
  double st[NC], nw[NC];
  int err = 0;
  memcpy(st, x, sizeof(st));
  for (int i = 0; i < c->N; i++) {
    double vy = st[1], epsi = st[3], s = st[4], ey = st[5];
    double cur = (lap == 0) ? lpv_ref_curvature(s, c->track, c->nseg, &err) : curv_ref[i];
    double vx = vel_ref[i];
    double delta = u[i * ND + 0];
    double *Ai = A + i * NC * NC, *Bi = B + i * NC * ND;
    ctrl_stage(&c->veh, Cf_new, Cf_new, c->dt, vx, vy, epsi, ey, cur, delta, Ai, Bi);
    if (C) for (int r = 0; r < NC; r++) C[i * NC + r] = c->dt * 0.0;
    for (int r = 0; r < NC; r++) {
      double a = 0.0;
      for (int k = 0; k < NC; k++) a += Ai[r * NC + k] * st[k];
      double b = 0.0;
      for (int k = 0; k < ND; k++) b += Bi[r * ND + k] * u[i * ND + k];
      nw[r] = a + b;
    }
    memcpy(st, nw, sizeof(st));
    if (states_out) memcpy(states_out + i * NC, st, sizeof(st));
  }
  return err;
}

int lpv_ref_ctrl_estimate(const lpv_ref_cfg *c, const double *traj, int ld_traj, const double *u, int ld_u,
                          double *A, double *B, double *C) {
  int err = 0;
  for (int i = 0; i < c->N; i++) {
    const double *t = traj + (size_t)i * ld_traj;
    double vy = t[1], epsi = t[3], s = t[4], ey = t[5];
    double cur = lpv_ref_curvature(s, c->track, c->nseg, &err);
    double vx = t[0];
    double delta = u[(size_t)i * ld_u];
    ctrl_stage(&c->veh, c->veh.Cf, c->veh.Cr, c->dt, vx, vy, epsi, ey, cur, delta, A + i * NC * NC, B + i * NC * ND);
    if (C) for (int r = 0; r < NC; r++) C[i * NC + r] = c->dt * 0.0;
  }
  return err;
}

int lpv_ref_plan_predict(const lpv_ref_cfg *c, const double *x, const double *SS, const double *u, double *states_out,
                         double *A, double *B, double *C) {
  double st[NP], nw[NP];
  int err = 0;
  memcpy(st, x, sizeof(st));
  for (int i = 0; i < c->N; i++) {
    double vx = st[0], vy = st[1], ey = st[3], epsi = st[4];
    double cur = lpv_ref_curvature(SS[i], c->track, c->nseg, &err);
    double delta = u[i * ND + 0];
    double *Ai = A + i * NP * NP, *Bi = B + i * NP * ND;
    plan_stage(&c->veh, c->dt, vx, vy, ey, epsi, cur, delta, Ai, Bi);
    if (C) for (int r = 0; r < NP; r++) C[i * NP + r] = c->dt * 0.0;
    for (int r = 0; r < NP; r++) {
      double a = 0.0;
      for (int k = 0; k < NP; k++) a += Ai[r * NP + k] * st[k];
      double b = 0.0;
      for (int k = 0; k < ND; k++) b += Bi[r * ND + k] * u[i * ND + k];
      nw[r] = a + b;
    }
    memcpy(st, nw, sizeof(st));
    if (states_out) memcpy(states_out + i * NP, st, sizeof(st));
  }
  return err;
}

int lpv_ref_plan_estimate(const lpv_ref_cfg *c, const double *traj, int ld_traj, const double *u, int ld_u,
                          double *A, double *B, double *C) {
  int err = 0;
  for (int i = 0; i < c->N; i++) {
    const double *t = traj + (size_t)i * ld_traj;
    double vx = t[0], vy = t[1], ey = t[3], epsi = t[4], s = t[5];
    double cur = lpv_ref_curvature(s, c->track, c->nseg, &err);
    double delta = u[(size_t)i * ld_u];
    plan_stage(&c->veh, c->dt, vx, vy, ey, epsi, cur, delta, A + i * NP * NP, B + i * NP * ND);
    if (C) for (int r = 0; r < NP; r++) C[i * NP + r] = c->dt * 0.0;
  }
  return err;
}

/* ------------------------------------------------------------------ QP assembly */
void lpv_ref_qp_free(lpv_ref_qp *qp) {
  free(qp->Pp); free(qp->Pi); free(qp->Px); free(qp->q);
  free(qp->Ap); free(qp->Ai); free(qp->Ax); free(qp->l); free(qp->u);
  memset(qp, 0, sizeof(*qp));
}

/* Mu (PathFollowingLPVMPC.py:401-425 / LPV_MPC_Planner.py:148-158): block-diag R + 2 diag(dR), last block
 * R + diag(dR), -dR on the +-2 off-diagonals.  Entry (i,j) of the (2N x 2N) matrix. */
static double mu_entry(const lpv_ref_cfg *c, int i, int j) {
  int N = c->N;
  int bi = i / ND, bj = j / ND, ri = i % ND, rj = j % ND;
  double v = 0.0;
  if (bi == bj) {
    v = c->R[ri * ND + rj] + (ri == rj ? 2 * c->dR[ri] : 0.0);
    if (bi == N - 1 && ri == rj) v = v - c->dR[ri];
  }
  /* np.fill_diagonal(Mu[2:], OffDiag) and (Mu[:, 2:], OffDiag) overwrite the +-2 diagonals */
  if (i - j == ND || j - i == ND) v = -c->dR[(i < j ? i : j) % ND];
  return v;
}

/* P = triu(csr_matrix(2*M0)) in csc, explicit zeros dropped; M0 = blkdiag(Q x (N+1), Mu) */
static void build_P(const lpv_ref_cfg *c, int n, lpv_ref_qp *qp) {
  int N = c->N, nx = n * (N + 1), nz = nx + ND * N;
  int cap = (N + 1) * n * n + 3 * ND * N * ND;
  qp->Pp = (int *)calloc((size_t)nz + 1, sizeof(int));
  qp->Pi = (int *)malloc(sizeof(int) * (size_t)cap);
  qp->Px = (double *)malloc(sizeof(double) * (size_t)cap);
  int k = 0;
  for (int j = 0; j < nz; j++) {
    qp->Pp[j] = k;
    if (j < nx) {
      int b = j / n, cj = j % n;
      for (int ri = 0; ri <= cj; ri++) {
        double v = 2 * c->Q[ri * n + cj];
        if (v != 0.0) { qp->Pi[k] = b * n + ri; qp->Px[k++] = v; }
      }
    } else {
      int uj = j - nx;
      int lo = uj - ND; if (lo < 0) lo = 0;
      for (int ui = lo; ui <= uj; ui++) {
        double v = 2 * mu_entry(c, ui, uj);
        if (v != 0.0) { qp->Pi[k] = nx + ui; qp->Px[k++] = v; }
      }
    }
  }
  qp->Pp[nz] = k; qp->pnz = k;
}

int lpv_ref_ctrl_qp(const lpv_ref_cfg *c, const double *A, const double *B, const double *C, const double *x0,
                    const double *vel_ref, int n_vel_ref, const double *old_steering, double old_accel,
                    lpv_ref_qp *qp) {
  const int n = NC, d = ND, N = c->N, delay = c->steering_delay;
  const int nx = n * (N + 1), nz = nx + d * N;
  const int mF = 2 * N + 4 * N, mG = nx + delay, m = mF + mG;
  memset(qp, 0, sizeof(*qp));
  qp->n = nz; qp->m = m;
  build_P(c, n, qp);
  /* q (PathFollowingLPVMPC.py:434-462): -2 * [xtrack, 0] . M0 ; slew term on u_0 */
  qp->q = (double *)calloc((size_t)nz, sizeof(double));
  for (int k = 0; k <= N; k++) {
    double vref = (k < N) ? vel_ref[k] : vel_ref[n_vel_ref - 1];
    for (int j = 0; j < n; j++) {
      /* row-vector times M0: sum_i xtrack_i * Q[i][j] with xtrack = [vref 0 0 0 0 0] */
      double acc = 0.0;
      for (int i = 0; i < n; i++) acc += (i == 0 ? vref : 0.0) * c->Q[i * n + j];
      qp->q[k * n + j] = -2 * acc;
    }
  }
  for (int j = 0; j < d * N; j++) qp->q[nx + j] = -2 * 0.0;
  {
    double uOld[2] = {old_steering[0], old_accel};
    for (int j = 0; j < d; j++) qp->q[nx + j] = -2 * (uOld[j] * c->dR[j]);
  }
  /* A = vstack([F, G]) in csc */
  int cap = 2 * N + 4 * N * 1 + nx + N * n * n + N * n * d + delay + 16;
  qp->Ap = (int *)calloc((size_t)nz + 1, sizeof(int));
  qp->Ai = (int *)malloc(sizeof(int) * (size_t)cap);
  qp->Ax = (double *)malloc(sizeof(double) * (size_t)cap);
  int k = 0;
  for (int col = 0; col < nz; col++) {
    qp->Ap[col] = k;
    if (col < nx) {
      int b = col / n, j = col % n;
      if (j == 0 && b < N) { /* Fx rows: -vx <= -0.01 ; vx <= max_vel (no terminal rows) */
        qp->Ai[k] = 2 * b; qp->Ax[k++] = -1.0;
        qp->Ai[k] = 2 * b + 1; qp->Ax[k++] = 1.0;
      }
      qp->Ai[k] = mF + col; qp->Ax[k++] = 1.0; /* Gx = I */
      if (b < N)
        for (int r = 0; r < n; r++) {
          double v = -A[(size_t)b * n * n + r * n + j];
          if (v != 0.0) { qp->Ai[k] = mF + (b + 1) * n + r; qp->Ax[k++] = v; }
        }
    } else {
      int b = (col - nx) / d, j = (col - nx) % d;
      qp->Ai[k] = 2 * N + 4 * b + 2 * j; qp->Ax[k++] = 1.0;
      qp->Ai[k] = 2 * N + 4 * b + 2 * j + 1; qp->Ax[k++] = -1.0;
      for (int r = 0; r < n; r++) {
        double v = -B[(size_t)b * n * d + r * d + j];
        if (v != 0.0) { qp->Ai[k] = mF + (b + 1) * n + r; qp->Ax[k++] = v; }
      }
      if (j == 0 && b < delay) { qp->Ai[k] = mF + nx + b; qp->Ax[k++] = 1.0; }
    }
  }
  qp->Ap[nz] = k; qp->anz = k;
  qp->l = (double *)malloc(sizeof(double) * (size_t)m);
  qp->u = (double *)malloc(sizeof(double) * (size_t)m);
  for (int b = 0; b < N; b++) {
    qp->u[2 * b] = -0.01; qp->u[2 * b + 1] = c->max_vel;
    qp->u[2 * N + 4 * b + 0] = 0.249; qp->u[2 * N + 4 * b + 1] = 0.249;
    qp->u[2 * N + 4 * b + 2] = 4.0; qp->u[2 * N + 4 * b + 3] = 1.0;
  }
  for (int i = 0; i < mF; i++) qp->l[i] = -INFINITY;
  /* beq = E x0 + L (np.add's third positional argument is `out`; Eu == 0) */
  for (int i = 0; i < mG; i++) {
    double e = (i < n) ? x0[i] : 0.0;
    double L = 0.0;
    if (i >= n && i < nx) L = C ? C[i - n] : 0.0;
    else if (i >= nx) L = old_steering[(i - nx) + 1];
    qp->l[mF + i] = qp->u[mF + i] = e + L;
  }
  return 0;
}

int lpv_ref_plan_qp(const lpv_ref_cfg *c, const double *A, const double *B, const double *C, const double *x0,
                    const double *u_old, double max_ey, const double *ey_lo, const double *ey_hi, lpv_ref_qp *qp) {
  const int n = NP, d = ND, N = c->N;
  const int nx = n * (N + 1), nz = nx + d * N, m = nx + nz;
  memset(qp, 0, sizeof(*qp));
  qp->n = nz; qp->m = m;
  build_P(c, n, qp);
  qp->q = (double *)calloc((size_t)nz, sizeof(double));
  for (int k = 0; k <= N; k++) for (int j = 0; j < n; j++) qp->q[k * n + j] = c->L_cf[j];
  for (int j = 0; j < d; j++) qp->q[nx + j] = -2 * (u_old[j] * c->dR[j]);
  int cap = nx + N * n * n + N * n * d + nz + 16;
  qp->Ap = (int *)calloc((size_t)nz + 1, sizeof(int));
  qp->Ai = (int *)malloc(sizeof(int) * (size_t)cap);
  qp->Ax = (double *)malloc(sizeof(double) * (size_t)cap);
  int k = 0;
  for (int col = 0; col < nz; col++) {
    qp->Ap[col] = k;
    if (col < nx) {
      int b = col / n, j = col % n;
      qp->Ai[k] = col; qp->Ax[k++] = 1.0;
      if (b < N)
        for (int r = 0; r < n; r++) {
          double v = -A[(size_t)b * n * n + r * n + j];
          if (v != 0.0) { qp->Ai[k] = (b + 1) * n + r; qp->Ax[k++] = v; }
        }
    } else {
      int b = (col - nx) / d, j = (col - nx) % d;
      for (int r = 0; r < n; r++) {
        double v = -B[(size_t)b * n * d + r * d + j];
        if (v != 0.0) { qp->Ai[k] = (b + 1) * n + r; qp->Ax[k++] = v; }
      }
    }
    qp->Ai[k] = nx + col; qp->Ax[k++] = 1.0; /* Aineq = I */
  }
  qp->Ap[nz] = k; qp->anz = k;
  qp->l = (double *)malloc(sizeof(double) * (size_t)m);
  qp->u = (double *)malloc(sizeof(double) * (size_t)m);
  for (int i = 0; i < nx; i++) {
    double e = (i < n) ? x0[i] : 0.0;
    double L = (i >= n && C) ? C[i - n] : 0.0;
    qp->l[i] = qp->u[i] = e + L;
  }
  {
    const double xmin[NP] = {c->min_vel, -1, -2, -max_ey, -0.8};
    const double xmax[NP] = {c->max_vel, 1, 2, max_ey, 0.8};
    const double umin[ND] = {-0.249, -0.7}, umax[ND] = {+0.249, +2.0};
    for (int b = 0; b <= N; b++)
      for (int j = 0; j < n; j++) {
        qp->l[nx + b * n + j] = xmin[j]; qp->u[nx + b * n + j] = xmax[j];
        if (j == 3 && ey_lo) qp->l[nx + b * n + j] = ey_lo[b];
        if (j == 3 && ey_hi) qp->u[nx + b * n + j] = ey_hi[b];
      }
    for (int b = 0; b < N; b++)
      for (int j = 0; j < d; j++) { qp->l[nx + nx + b * d + j] = umin[j]; qp->u[nx + nx + b * d + j] = umax[j]; }
  }
  return 0;
}

/* ------------------------------------------------------------------ whole path */
static int solve_and_unpack(const lpv_ref_qp *qp, const osqp_ref_settings *st, int n, int N, double *xPred,
                            double *uPred, lpv_ref_info *info, unsigned char *alo, unsigned char *aup, double *xs,
                            double *zs, double *ys) {
  osqp_ref_result r; memset(&r, 0, sizeof(r));
  double *x = (double *)malloc(sizeof(double) * (size_t)qp->n);
  double *y = (double *)malloc(sizeof(double) * (size_t)qp->m);
  r.x = x; r.y = y; r.active_lo = alo; r.active_up = aup; r.xs = xs; r.zs = zs; r.ys = ys;
  int rc = osqp_ref_solve(qp->n, qp->m, qp->Pp, qp->Pi, qp->Px, qp->q, qp->Ap, qp->Ai, qp->Ax, qp->l, qp->u, st, &r);
  if (rc == 0) {
    /* PathFollowingLPVMPC.py:157-162 : z = [x_0..x_N, u_0..u_{N-1}] */
    memcpy(xPred, x, sizeof(double) * (size_t)(n * (N + 1)));
    memcpy(uPred, x + n * (N + 1), sizeof(double) * (size_t)(ND * N));
    if (info) {
      info->status = r.status; info->iter = r.iter; info->rho_updates = r.rho_updates;
      info->status_polish = r.status_polish; info->n_factor = r.n_factor;
      info->obj_val = r.obj_val; info->pri_res = r.pri_res; info->dua_res = r.dua_res;
    }
  }
  free(x); free(y);
  return rc;
}

int lpv_ref_ctrl_solve(const lpv_ref_cfg *c, const osqp_ref_settings *st, int mode, const double *x0,
                       const double *A, const double *B, const double *C, const double *x_sched,
                       const double *u_prev, const double *vel_ref, int n_vel_ref, const double *curv_ref,
                       double Cf_new, int lap, const double *traj, const double *old_steering, double old_accel,
                       double *xPred, double *uPred, lpv_ref_info *info, unsigned char *active_lo,
                       unsigned char *active_up, double *xs, double *zs, double *ys) {
  int N = c->N, err = 0;
  double *Ab = 0, *Bb = 0, *Cb = 0;
  if (mode != 0) {
    Ab = (double *)malloc(sizeof(double) * (size_t)N * NC * NC);
    Bb = (double *)malloc(sizeof(double) * (size_t)N * NC * ND);
    Cb = (double *)calloc((size_t)N * NC, sizeof(double));
    if (mode == 1) err = lpv_ref_ctrl_predict(c, x_sched, u_prev, vel_ref, curv_ref, Cf_new, lap, 0, Ab, Bb, Cb);
    else err = lpv_ref_ctrl_estimate(c, traj, NC, u_prev, ND, Ab, Bb, Cb);
    A = Ab; B = Bb; C = Cb;
  }
  if (info) info->sched_err = err;
  lpv_ref_qp qp;
  lpv_ref_ctrl_qp(c, A, B, C, x0, vel_ref, n_vel_ref, old_steering, old_accel, &qp);
  int rc = solve_and_unpack(&qp, st, NC, N, xPred, uPred, info, active_lo, active_up, xs, zs, ys);
  lpv_ref_qp_free(&qp);
  free(Ab); free(Bb); free(Cb);
  return rc;
}

int lpv_ref_plan_solve(const lpv_ref_cfg *c, const osqp_ref_settings *st, int mode, const double *x0,
                       const double *A, const double *B, const double *C, const double *x_sched, const double *SS,
                       const double *u_prev, const double *traj, const double *u_old, double max_ey,
                       const double *ey_lo, const double *ey_hi, double *xPred, double *uPred, lpv_ref_info *info,
                       unsigned char *active_lo, unsigned char *active_up, double *xs, double *zs, double *ys) {
  int N = c->N, err = 0;
  double *Ab = 0, *Bb = 0, *Cb = 0;
  if (mode != 0) {
    Ab = (double *)malloc(sizeof(double) * (size_t)N * NP * NP);
    Bb = (double *)malloc(sizeof(double) * (size_t)N * NP * ND);
    Cb = (double *)calloc((size_t)N * NP, sizeof(double));
    if (mode == 1) err = lpv_ref_plan_predict(c, x_sched, SS, u_prev, 0, Ab, Bb, Cb);
    else err = lpv_ref_plan_estimate(c, traj, 6, u_prev, 1, Ab, Bb, Cb);
    A = Ab; B = Bb; C = Cb;
  }
  if (info) info->sched_err = err;
  lpv_ref_qp qp;
  lpv_ref_plan_qp(c, A, B, C, x0, u_old, max_ey, ey_lo, ey_hi, &qp);
  int rc = solve_and_unpack(&qp, st, NP, N, xPred, uPred, info, active_lo, active_up, xs, zs, ys);
  lpv_ref_qp_free(&qp);
  free(Ab); free(Bb); free(Cb);
  return rc;
}

int lpv_ref_ctrl_batch(const lpv_ref_cfg *c, const osqp_ref_settings *st, int B, const double *x0,
                       const double *u_prev, const double *vel_ref, const double *curv_ref, const int *lap,
                       const double *u_old, double Cf_new, int threads, double *xPred, double *uPred, int *status,
                       int *iters) {
  int N = c->N, solved = 0;
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : solved)
  for (int b = 0; b < B; b++) {
    lpv_ref_info info; memset(&info, 0, sizeof(info));
    double old_st[1] = {u_old[b * ND + 0]};
    lpv_ref_cfg cc = *c; cc.steering_delay = 0;
    lpv_ref_ctrl_solve(&cc, st, 1, x0 + (size_t)b * NC, 0, 0, 0, x0 + (size_t)b * NC, u_prev + (size_t)b * N * ND,
                       vel_ref + (size_t)b * (N + 1), N + 1, curv_ref + (size_t)b * N, Cf_new, lap ? lap[b] : 1, 0,
                       old_st, u_old[b * ND + 1], xPred + (size_t)b * (N + 1) * NC, uPred + (size_t)b * N * ND, &info,
                       0, 0, 0, 0, 0);
    if (status) status[b] = info.sched_err ? -20 : info.status;
    if (iters) iters[b] = info.iter;
    solved += (info.status == OSQP_REF_SOLVED);
  }
  return solved;
}

int lpv_ref_plan_batch(const lpv_ref_cfg *c, const osqp_ref_settings *st, int B, const double *x0, const double *SS,
                       const double *u_prev, const double *u_old, const double *max_ey, const double *ey_lo,
                       const double *ey_hi, int threads, double *xPred, double *uPred, int *status, int *iters) {
  int N = c->N, solved = 0;
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : solved)
  for (int b = 0; b < B; b++) {
    lpv_ref_info info; memset(&info, 0, sizeof(info));
    lpv_ref_plan_solve(c, st, 1, x0 + (size_t)b * NP, 0, 0, 0, x0 + (size_t)b * NP, SS + (size_t)b * (N + 1),
                       u_prev + (size_t)b * N * ND, 0, u_old + (size_t)b * ND, max_ey[b],
                       ey_lo ? ey_lo + (size_t)b * (N + 1) : 0, ey_hi ? ey_hi + (size_t)b * (N + 1) : 0,
                       xPred + (size_t)b * (N + 1) * NP, uPred + (size_t)b * N * ND, &info, 0, 0, 0, 0, 0);
    if (status) status[b] = info.sched_err ? -20 : info.status;
    if (iters) iters[b] = info.iter;
    solved += (info.status == OSQP_REF_SOLVED);
  }
  return solved;
}
