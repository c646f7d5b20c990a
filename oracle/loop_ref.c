/*
 * TEST INFRASTRUCTURE — CPU restatement of the closed-loop tick AROUND the controller QP (SURVEY.md 8f row 1,
 * BASELINE configs[3]): the simulator's vehicle model, the inertial -> curvilinear localisation, the warm-up
 * guess and the controller main loop's per-tick bookkeeping.
 *
 * Follows (file:line under /root/reference/workspace/src/barc/src):
 *   Simulator.f                    vehicleSimulator.py:164-199 (linear tyres Fy = 60 alpha; u = [motor, servo], :337)
 *   Map.getLocalPosition           Utilities/trackInitialization.py:283-383 (+ computeAngle :388-410)
 *   Map.getGlobalPosition          Utilities/trackInitialization.py:205-260
 *   predicted_vectors_generation   controllerMain.py:510-553
 *   controller main loop, lap 0    controllerMain.py:177-192 (measure, clamp, localise), :252-257 (lap logic),
 *                                  :289-298 (OldSteering / OldAccelera), :310-331 (warm-up / LPVPrediction + solve),
 *                                  :381-383 (command)
 *
 * Pinned against the reference's own Python (Simulator.f, Map.getLocalPosition, Map.getGlobalPosition and the
 * controller class run headless in the build container) through tests/golden/closed_loop.npz
 * (generator: tests/golden/make_golden_loop.py).  The product path never links or calls this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "loop_ref.h"

#define NC 6
#define ND 2

static const double kPi = 3.141592653589793;

void loop_ref_sim_f(double *st, const double *u, const lpv_ref_vehicle *v, double mu, double dt) {
  double x = st[0], y = st[1], yaw = st[2], vx = st[3], vy = st[4], psiDot = st[5], ax = st[6], ay = st[7];
  double a_F = 0.0, a_R = 0.0;
  if (fabs(vx) > 0.2) { /* :168-170 */
    a_F = u[1] - atan((vy + v->lf * psiDot) / fabs(vx));
    a_R = atan((-vy + v->lr * psiDot) / fabs(vx));
  }
  double FyF = 60 * a_F, FyR = 60 * a_R; /* :174-175 */
  st[0] = x + dt * (cos(yaw) * vx - sin(yaw) * vy);           /* :190 */
  st[1] = y + dt * (sin(yaw) * vx + cos(yaw) * vy);           /* :191 */
  st[3] = vx + dt * (ax + psiDot * vy);                       /* :192 */
  st[4] = vy + dt * (ay - psiDot * vx);                       /* :193 */
  st[6] = u[0] - mu * vx - FyF / v->m * sin(u[1]);            /* :194 */
  st[7] = 1.0 / v->m * (FyF * cos(u[1]) + FyR);               /* :195 */
  st[2] = yaw + dt * (psiDot);                                /* :196 */
  st[5] = psiDot + dt * (1.0 / v->Iz * (v->lf * FyF * cos(u[1]) - v->lr * FyR)); /* :197 */
  st[3] = fabs(st[3]);                                        /* :199 */
}

/* trackInitialization.py:388-410 */
static double compute_angle(double p1x, double p1y, double ox, double oy, double p2x, double p2y) {
  double v1x = p1x - ox, v1y = p1y - oy, v2x = p2x - ox, v2y = p2y - oy;
  double dot = v1x * v2x + v1y * v2y;
  double det = v1x * v2y - v1y * v2x;
  return atan2(det, dot);
}

/* np.unwrap([a, b])[1] (numpy >= 1.21: period 2 pi, discont pi) */
static double unwrap_second(double a, double b) {
  const double period = 2 * kPi, hi = kPi, lo = -kPi;
  double dd = b - a;
  double md = fmod(dd - lo, period);
  if (md != 0.0 && md < 0.0) md += period; /* python-style modulo with a positive divisor */
  double ddmod = md + lo;
  if (ddmod == lo && dd > 0) ddmod = hi;
  double corr = ddmod - dd;
  if (fabs(dd) < kPi) corr = 0.0;
  return b + corr;
}

static double norm2(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

static double sgn0(double v) { return (v > 0) - (v < 0); } /* np.sign */

int loop_ref_local_position(const double *track, int nseg, double half_width, double slack, double x, double y,
                            double psi, double *out) {
  int done = 0;
  double s = 0, ey = 0, epsi = 0;
  for (int i = 0; i < nseg && !done; i++) {
    const double *Pi = track + i * 6;
    const double *Pm = track + ((i == 0) ? (nseg - 1) : (i - 1)) * 6; /* python's [-1] wrap for i = 0 */
    double xf = Pi[0], yf = Pi[1], xs = Pm[0], ys = Pm[1];
    if (Pi[5] == 0.0) { /* straight, :295-325 */
      double psi_unwrap = unwrap_second(Pm[2], psi);
      epsi = psi_unwrap - Pm[2];
      if (norm2(xs - x, ys - y) == 0) { s = Pi[3]; ey = 0; done = 1; }
      else if (norm2(xf - x, yf - y) == 0) { s = Pi[3] + Pi[4]; ey = 0; done = 1; }
      else if (fabs(compute_angle(x, y, xs, ys, xf, yf)) <= kPi / 2 && fabs(compute_angle(x, y, xf, yf, xs, ys)) <= kPi / 2) {
        double v1 = norm2(x - xs, y - ys);
        double angle = compute_angle(xf, yf, xs, ys, x, y);
        double s_local = v1 * cos(angle);
        s = s_local + Pi[3];
        ey = v1 * sin(angle);
        if (fabs(ey) <= half_width + slack) done = 1;
      }
    } else { /* arc, :327-369 */
      double r = 1 / Pi[5];
      double direction = (r >= 0) ? 1 : -1;
      double ang = Pm[2];
      double CenterX = xs + fabs(r) * cos(ang + direction * kPi / 2);
      double CenterY = ys + fabs(r) * sin(ang + direction * kPi / 2);
      if (norm2(xs - x, ys - y) == 0) {
        ey = 0; epsi = unwrap_second(ang, psi) - ang; s = Pi[3]; done = 1;
      } else if (norm2(xf - x, yf - y) == 0) {
        s = Pi[3] + Pi[4]; ey = 0; epsi = unwrap_second(Pi[2], psi) - Pi[2]; done = 1;
      } else {
        double arc1 = Pi[4] * Pi[5];
        double arc2 = compute_angle(xs, ys, CenterX, CenterY, x, y);
        if (sgn0(arc1) == sgn0(arc2) && fabs(arc1) >= fabs(arc2)) {
          double vn = norm2(x - CenterX, y - CenterY);
          double s_local = fabs(arc2) * fabs(r);
          s = s_local + Pi[3];
          ey = -sgn0(direction) * (vn - fabs(r));
          epsi = unwrap_second(ang + arc2, psi) - (ang + arc2);
          if (fabs(ey) <= half_width + slack) done = 1;
        }
      }
    }
  }
  if (!done) { s = 10000; ey = 10000; epsi = 10000; } /* :375-378 */
  out[0] = s; out[1] = ey; out[2] = epsi;
  return done;
}

static double wrap_angle(double a) { /* trackInitialization.py:413-421 */
  if (a < -kPi) return 2 * kPi + a;
  if (a > kPi) return a - 2 * kPi;
  return a;
}

int loop_ref_global_position(const double *track, int nseg, double s, double ey, double *out) {
  double TrackLength = track[(nseg - 1) * 6 + 3] + track[(nseg - 1) * 6 + 4];
  while (s > TrackLength) s = s - TrackLength; /* :211-212 */
  int i = -1, cnt = 0;
  for (int k = 0; k < nseg; k++)
    if (s >= track[k * 6 + 3] && s < track[k * 6 + 3] + track[k * 6 + 4]) { if (i < 0) i = k; cnt++; }
  if (cnt != 1) { out[0] = out[1] = out[2] = NAN; return 1; }
  const double *Pi = track + i * 6;
  const double *Pm = track + ((i == 0) ? (nseg - 1) : (i - 1)) * 6;
  if (Pi[5] == 0.0) { /* :221-235 */
    double xf = Pi[0], yf = Pi[1], xs = Pm[0], ys = Pm[1], psi = Pi[2];
    double deltaL = Pi[4], reltaL = s - Pi[3];
    out[0] = (1 - reltaL / deltaL) * xs + reltaL / deltaL * xf + ey * cos(psi + kPi / 2);
    out[1] = (1 - reltaL / deltaL) * ys + reltaL / deltaL * yf + ey * sin(psi + kPi / 2);
    out[2] = psi;
  } else { /* :236-258 */
    double r = 1 / Pi[5], ang = Pm[2];
    double direction = (r >= 0) ? 1 : -1;
    double CenterX = Pm[0] + fabs(r) * cos(ang + direction * kPi / 2);
    double CenterY = Pm[1] + fabs(r) * sin(ang + direction * kPi / 2);
    double spanAng = (s - Pi[3]) / (kPi * fabs(r)) * kPi;
    double angleNormal = wrap_angle(direction * kPi / 2 + ang);
    double angle = -(kPi - fabs(angleNormal)) * ((angleNormal >= 0) ? 1 : -1);
    out[0] = CenterX + (fabs(r) - direction * ey) * cos(angle + direction * spanAng);
    out[1] = CenterY + (fabs(r) - direction * ey) * sin(angle + direction * spanAng);
    out[2] = ang + direction * spanAng;
  }
  return 0;
}

static const double kGuessDv[20] = {0.05, 0.2, 0.4, 0.6, 0.7, 0.8, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9};
static const double kGuessDs[20] = {0, 0.01, 0.02, 0.04, 0.07, 0.1, 0.14, 0.18, 0.23, 0.55, 0.66, 0.77, 0.89, 1.00, 1.19, 1.39, 1.59, 1.79, 1.89, 1.999};
static const double kGuessUa[20] = {0.0, 0.3, 0.5, 0.7, 0.8, 0.9, 0.9, 0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.30, 0.22, 0.18, 0.14, 0.1, 0.1, 0.1};

void loop_ref_guess(const double *local, int N, double *xx, double *uu) {
  for (int i = 0; i < N && i < 20; i++) {
    xx[i * 6 + 0] = local[0] + kGuessDv[i]; xx[i * 6 + 1] = local[1]; xx[i * 6 + 2] = local[2];
    xx[i * 6 + 3] = 0.0001; xx[i * 6 + 4] = local[4] + kGuessDs[i]; xx[i * 6 + 5] = 0.0001;
    uu[i * 2 + 0] = 0.; uu[i * 2 + 1] = kGuessUa[i];
  }
}

enum { C_FIRST_IT = 0, C_LAP, C_HALF, C_STATUS, C_ITERS, C_FAIL, C_FAIL_TICK, C_TICKS };
enum { S_SOLVED = 0, S_ITERS, S_MAX_EY, S_LAP_TICK };

static int feasible(int status) { return status == 1 || status == 2 || status == -2; } /* PathFollowingLPVMPC.py:322-324 */

long loop_ref_run(const lpv_ref_cfg *c, const osqp_ref_settings *st, const loop_ref_cfg *lc, int B, int n_ticks,
                  double *sim, double *cmd, double *u_pred, double *local, int *ctr, double *stat, double *x_pred,
                  int threads) {
  const int N = c->N;
  if (N > 20 || N < 1) return -1;
  if (threads < 1) threads = 1;
  const double TrackLength = c->track[(c->nseg - 1) * 6 + 3] + c->track[(c->nseg - 1) * 6 + 4];
  long solved_total = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : solved_total)
  for (int b = 0; b < B; b++) {
    double *S = sim + (size_t)b * 8, *cm = cmd + (size_t)b * 2, *up = u_pred + (size_t)b * N * ND, *loc = local + (size_t)b * 6;
    int *ct = ctr + (size_t)b * 8;
    double *sv = stat + (size_t)b * 4;
    double vel_ref[21], curv_ref[20], xx[20 * 6], uu[20 * 2], xP[21 * 6], uP[20 * 2];
    lpv_ref_cfg cc = *c; cc.steering_delay = 0;
    for (int k = 0; k <= N; k++) vel_ref[k] = lc->vel_ref;
    for (int k = 0; k < N; k++) curv_ref[k] = 0.0;
    for (int t = 0; t < n_ticks; t++) {
      if (ct[C_FAIL]) break;
      /* measure: [vx vy wz x y psi], vx clamp (controllerMain.py:179-184) */
      double lp[3];
      loc[0] = S[3] < 0.01 ? 0.01 : S[3]; loc[1] = S[4]; loc[2] = S[5];
      int ok = loop_ref_local_position(c->track, c->nseg, lc->half_width, lc->slack, S[0], S[1], S[2], lp);
      /* :188 LocalState[4], LocalState[3], LocalState[5] = s, ey, epsi */
      loc[4] = lp[0];
      if (lc->swap_ey_epsi) { loc[3] = lp[1]; loc[5] = lp[2]; } else { loc[3] = lp[2]; loc[5] = lp[1]; }
      if (!ok) { ct[C_FAIL] = LOOP_REF_OFF_TRACK; ct[C_FAIL_TICK] = ct[C_TICKS]; ct[C_STATUS] = LOOP_REF_OFF_TRACK; break; }
      if (fabs(lp[1]) > sv[S_MAX_EY]) sv[S_MAX_EY] = fabs(lp[1]);
      /* lap logic :190-192, :252-257 */
      if (loc[4] >= 3 * TrackLength / 4) ct[C_HALF] = 1;
      if (ct[C_HALF] == 1 && loc[4] <= TrackLength / 4) {
        ct[C_HALF] = 0; ct[C_LAP] += 1;
        if (sv[S_LAP_TICK] < 0) sv[S_LAP_TICK] = ct[C_TICKS];
      }
      /* :289-298 : uOld = the command applied during the last period */
      double old_st[1] = {cm[0]}, old_acc = cm[1];
      lpv_ref_info info; memset(&info, 0, sizeof(info));
      if (ct[C_FIRST_IT] <= lc->warmup_ticks) { /* :310-315 (first_it < 10) */
        loop_ref_guess(loc, N, xx, uu);
        lpv_ref_ctrl_solve(&cc, st, 2, loc, 0, 0, 0, 0, uu, vel_ref, N, 0, lc->Cf_new, 0, xx, old_st, old_acc, xP, uP,
                           &info, 0, 0, 0, 0, 0);
        ct[C_FIRST_IT] += 1;
      } else { /* :326-331 : x0 = LPV_States_Prediction[0,:] */
        double states[20 * 6], Ab[20 * 36], Bb[20 * 12], Cb[20 * 6];
        int err = lpv_ref_ctrl_predict(&cc, loc, up, vel_ref, curv_ref, lc->Cf_new, 0, states, Ab, Bb, Cb);
        if (err) { info.sched_err = 1; }
        else lpv_ref_ctrl_solve(&cc, st, 0, states, Ab, Bb, Cb, 0, 0, vel_ref, N + 1, 0, lc->Cf_new, 0, 0, old_st, old_acc,
                                xP, uP, &info, 0, 0, 0, 0, 0);
      }
      int status = info.sched_err ? LOOP_REF_SCHEDULE_ERROR : info.status;
      ct[C_STATUS] = status; ct[C_ITERS] = info.iter;
      if (!feasible(status)) { ct[C_FAIL] = status; ct[C_FAIL_TICK] = ct[C_TICKS]; break; }
      if (status == 1) { sv[S_SOLVED] += 1; solved_total += 1; }
      sv[S_ITERS] += info.iter;
      memcpy(up, uP, sizeof(double) * (size_t)N * ND);
      if (x_pred) memcpy(x_pred + (size_t)b * (N + 1) * NC, xP, sizeof(double) * (size_t)(N + 1) * NC);
      cm[0] = uP[0]; cm[1] = uP[1]; /* :381-383, delays 0 */
      double u[2] = {cm[1], cm[0]}; /* ecu: [motor, servo] */
      for (int k = 0; k < lc->substeps; k++) loop_ref_sim_f(S, u, &c->veh, lc->sim_mu, lc->sim_dt);
      ct[C_TICKS] += 1;
    }
  }
  return solved_total;
}

/* ------------------------------------------------------------------------------------------------
 * planner loop: plannerMain.py:128-224 */
#define NPL 5

void plan_loop_ref_guess(const double *x0, int N, double accel_rate, double dt, double s0, double *xx) {
  /* plannerMain.py:465-505: Vx integrates Accel = 0.1 + accel_rate i, the other states are held, S integrates with curv = 0 */
  double Vx = x0[0], S = s0;
  const double curv = 0;
  for (int i = 0; i <= N; i++) {
    double *row = xx + i * 6;
    row[0] = Vx; row[1] = x0[1]; row[2] = x0[2]; row[3] = x0[3]; row[4] = x0[4]; row[5] = S;
    if (i < N) {
      double Accel = 0.1 + accel_rate * i;
      double Vn = Vx + Accel * dt;
      S = S + ((Vx * cos(x0[4]) - x0[1] * sin(x0[4])) / (1 - x0[3] * curv)) * dt;
      Vx = Vn;
    }
  }
}

enum { PC_TICKS = 0, PC_STATUS, PC_ITERS, PC_FAIL, PC_FAIL_TICK };

long plan_loop_ref_run(const lpv_ref_cfg *c, const osqp_ref_settings *st, int B, int n_ticks, double max_ey,
                       double accel_rate, const double *xstart, double *x_pred, double *u_pred, double *SS, int *ctr,
                       double *stat, int threads) {
  const int N = c->N;
  if (threads < 1) threads = 1;
  long solved_total = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : solved_total)
  for (int b = 0; b < B; b++) {
    double *xP = x_pred + (size_t)b * (N + 1) * NPL, *uP = u_pred + (size_t)b * N * ND, *ss = SS + (size_t)b * (N + 1);
    int *ct = ctr + (size_t)b * 8;
    double *sv = stat + (size_t)b * 4;
    double *xx = (double *)malloc(sizeof(double) * (size_t)(N + 1) * 6);
    double *uu = (double *)calloc((size_t)N, sizeof(double));
    double *nx = (double *)malloc(sizeof(double) * (size_t)(N + 1) * NPL), *nu = (double *)malloc(sizeof(double) * (size_t)N * ND);
    const double u_old[2] = {0.0, 0.0};
    for (int t = 0; t < n_ticks; t++) {
      if (ct[PC_FAIL]) break;
      lpv_ref_info info; memset(&info, 0, sizeof(info));
      if (ct[PC_TICKS] == 0) { /* first_it == 1 */
        plan_loop_ref_guess(xstart + (size_t)b * NPL, N, accel_rate, c->dt, ss[0], xx);
        lpv_ref_plan_solve(c, st, 2, xstart + (size_t)b * NPL, 0, 0, 0, 0, 0, uu, xx, u_old, max_ey, 0, 0, nx, nu, &info, 0, 0, 0, 0, 0);
      } else {
        double x1[NPL];
        for (int q = 0; q < NPL; q++) x1[q] = xP[NPL + q];
        lpv_ref_plan_solve(c, st, 1, x1, 0, 0, 0, x1, ss, uP, 0, u_old, max_ey, 0, 0, nx, nu, &info, 0, 0, 0, 0, 0);
      }
      int status = info.sched_err ? LOOP_REF_SCHEDULE_ERROR : info.status;
      ct[PC_STATUS] = status; ct[PC_ITERS] = info.iter;
      if (!feasible(status)) { ct[PC_FAIL] = status; ct[PC_FAIL_TICK] = ct[PC_TICKS]; break; }
      if (status == 1) { sv[0] += 1; solved_total += 1; }
      sv[1] += info.iter;
      memcpy(xP, nx, sizeof(double) * (size_t)(N + 1) * NPL);
      memcpy(uP, nu, sizeof(double) * (size_t)N * ND);
      /* plannerMain.py:201-211 */
      int err = 0;
      for (int j = 0; j < N; j++) {
        double curv = lpv_ref_curvature(ss[j], c->track, c->nseg, &err);
        const double *x = xP + j * NPL;
        ss[j + 1] = (ss[j] + ((x[0] * cos(x[4]) - x[1] * sin(x[4])) / (1 - x[3] * curv)) * c->dt);
      }
      ss[0] = ss[1];
      ct[PC_TICKS] += 1;
      if (err) { ct[PC_FAIL] = LOOP_REF_SCHEDULE_ERROR; ct[PC_FAIL_TICK] = ct[PC_TICKS]; ct[PC_STATUS] = LOOP_REF_SCHEDULE_ERROR; break; }
    }
    free(xx); free(uu); free(nx); free(nu);
  }
  return solved_total;
}

/* ------------------------------------------------------------------------------------------------
 * controller <- planner hand-off: controllerMain.py:196-243 (lap >= 1 branch) with Body_Frame_Errors (:495-506) */
double track_inputs_ref(const double *g, int lap, double s_prev, const double *refs, int n_ref, int index, int N, double dt,
                        double *x0, double *vel_ref, double *curv_ref) {
  const double *x_d = refs + index, *y_d = refs + n_ref + index, *psi_d = refs + 2 * n_ref + index;
  const double *vx_d = refs + 3 * n_ref + index, *curv_d = refs + 4 * n_ref + index;
  double psi = g[5] - 2 * kPi * lap;                                  /* :200 */
  psi = wrap_angle(psi);                                              /* :202 */
  const double x = g[3], y = g[4], vx = g[0], vy = g[1];
  const double xd = x_d[0], yd = y_d[0], psid = psi_d[0], curv = curv_d[0];   /* :238-240: x_ref[0], ..., curv_ref[0] */
  const double ex = (x - xd) * cos(psid) + (y - yd) * sin(psid);      /* :497 */
  const double ey = -(x - xd) * sin(psid) + (y - yd) * cos(psid);     /* :499 */
  const double epsi = wrap_angle(psi - psid);                         /* :501 */
  const double s = s_prev + ((vx * cos(epsi) - vy * sin(epsi)) / (1 - ey * curv)) * dt;   /* :504 */
  x0[0] = vx; x0[1] = vy; x0[2] = g[2]; x0[3] = epsi; x0[4] = s; x0[5] = ey;
  for (int i = 0; i < N; ++i) { vel_ref[i] = vx_d[i]; curv_ref[i] = curv_d[i]; }   /* :232-233 */
  vel_ref[N] = vx_d[N - 1];
  return ex;
}
