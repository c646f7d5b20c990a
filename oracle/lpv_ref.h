/* TEST INFRASTRUCTURE — CPU restatement of the reference's LPV scheduling + QP build
 * (see lpv_ref.c).  Never linked into the product library. */
#ifndef LPV_REF_H
#define LPV_REF_H

#include "osqp_ref.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double lf, lr, m, Iz, Cf, Cr, mu;
} lpv_ref_vehicle;

typedef struct {
  int N;              /* horizon */
  double dt;
  double Q[36];       /* n x n row-major (controller n=6, planner n=5 uses the leading 25) */
  double R[4];        /* 2 x 2 */
  double dR[2];
  double L_cf[5];     /* planner linear cost */
  double max_vel, min_vel;
  int steering_delay; /* controller only */
  lpv_ref_vehicle veh;
  int nseg;           /* rows of PointAndTangent */
  const double *track;/* nseg x 6 */
} lpv_ref_cfg;

/* utilities.py:31-50.  *err = 1 when no (or no unique) segment contains s (the reference raises). */
double lpv_ref_curvature(double s, const double *track, int nseg, int *err);

/* PathFollowingLPVMPC.py:166-258 */
int lpv_ref_ctrl_predict(const lpv_ref_cfg *c, const double *x, const double *u, const double *vel_ref,
                         const double *curv_ref, double Cf_new, int lap, double *states, double *A, double *B,
                         double *C);
/* PathFollowingLPVMPC.py:732-809 ; traj has ld_traj columns (>=6), u has ld_u columns */
int lpv_ref_ctrl_estimate(const lpv_ref_cfg *c, const double *traj, int ld_traj, const double *u, int ld_u,
                          double *A, double *B, double *C);
/* LPV_MPC_Planner.py:242-320 */
int lpv_ref_plan_predict(const lpv_ref_cfg *c, const double *x, const double *SS, const double *u, double *states,
                         double *A, double *B, double *C);
/* LPV_MPC_Planner.py:519-591 ; traj rows [vx vy w ey epsi s], u = steering per stage (stride ld_u) */
int lpv_ref_plan_estimate(const lpv_ref_cfg *c, const double *traj, int ld_traj, const double *u, int ld_u,
                          double *A, double *B, double *C);

typedef struct {
  int n, m, pnz, anz;
  int *Pp, *Pi, *Ap, *Ai;
  double *Px, *q, *Ax, *l, *u;
} lpv_ref_qp;
void lpv_ref_qp_free(lpv_ref_qp *qp);

/* PathFollowingLPVMPC.py:89-148,273-313,329-529 : OSQP-form QP exactly as handed to osqp.setup
 * (P upper triangle, rows [F; G], l = -inf on F rows).  old_steering has 1+delay entries. */
int lpv_ref_ctrl_qp(const lpv_ref_cfg *c, const double *A, const double *B, const double *C, const double *x0,
                    const double *vel_ref, int n_vel_ref, const double *old_steering, double old_accel,
                    lpv_ref_qp *qp);
/* LPV_MPC_Planner.py:86-205 : rows [Aeq; I].  ey_lo/ey_hi optional per-stage overrides (N+1 each) of the
 * lateral-error box (our "obstacle" hook, SURVEY.md 8d cfg 3); NULL => +-max_ey. */
int lpv_ref_plan_qp(const lpv_ref_cfg *c, const double *A, const double *B, const double *C, const double *x0,
                    const double *u_old, double max_ey, const double *ey_lo, const double *ey_hi, lpv_ref_qp *qp);

typedef struct {
  int status, iter, rho_updates, status_polish, n_factor, sched_err;
  double obj_val, pri_res, dua_res;
} lpv_ref_info;

/* Whole path for one controller QP: (optional schedule) + build + OSQP + unpack.
 * mode 0: A,B,C given;  1: LPVPrediction fused (x_sched,u_prev,vel_ref,curv_ref,lap);  2: _EstimateABC.
 * Outputs xPred (N+1,6), uPred (N,2); active_lo/up (m) optional; zs.. optional scaled iterates. */
int lpv_ref_ctrl_solve(const lpv_ref_cfg *c, const osqp_ref_settings *st, int mode, const double *x0,
                       const double *A, const double *B, const double *C, const double *x_sched,
                       const double *u_prev, const double *vel_ref, int n_vel_ref, const double *curv_ref,
                       double Cf_new, int lap, const double *traj, const double *old_steering, double old_accel,
                       double *xPred, double *uPred, lpv_ref_info *info, unsigned char *active_lo,
                       unsigned char *active_up, double *xs, double *zs, double *ys);
int lpv_ref_plan_solve(const lpv_ref_cfg *c, const osqp_ref_settings *st, int mode, const double *x0,
                       const double *A, const double *B, const double *C, const double *x_sched, const double *SS,
                       const double *u_prev, const double *traj, const double *u_old, double max_ey,
                       const double *ey_lo, const double *ey_hi, double *xPred, double *uPred, lpv_ref_info *info,
                       unsigned char *active_lo, unsigned char *active_up, double *xs, double *zs, double *ys);

/* The reference's per-instance loop over a batch (mode 1 for every QP), `threads` OpenMP threads over
 * disjoint index ranges.  Row-major [B, ...] arrays.  Returns number of SOLVED QPs. */
int lpv_ref_ctrl_batch(const lpv_ref_cfg *c, const osqp_ref_settings *st, int B, const double *x0,
                       const double *u_prev, const double *vel_ref, const double *curv_ref, const int *lap,
                       const double *u_old, double Cf_new, int threads, double *xPred, double *uPred, int *status,
                       int *iters);
int lpv_ref_plan_batch(const lpv_ref_cfg *c, const osqp_ref_settings *st, int B, const double *x0, const double *SS,
                       const double *u_prev, const double *u_old, const double *max_ey, const double *ey_lo,
                       const double *ey_hi, int threads, double *xPred, double *uPred, int *status, int *iters);

#ifdef __cplusplus
}
#endif
#endif
