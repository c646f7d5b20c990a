/* TEST INFRASTRUCTURE — CPU restatement of the closed-loop tick around the controller QP (see loop_ref.c).
 * Never linked into the product library. */
#ifndef LOOP_REF_H
#define LOOP_REF_H

#include "lpv_ref.h"

#ifdef __cplusplus
extern "C" {
#endif

/* status codes the loop adds to OSQP's */
#define LOOP_REF_SCHEDULE_ERROR (-20)
#define LOOP_REF_OFF_TRACK (-22)

typedef struct {
  double sim_dt;       /* simulator/dt, MAIN_LAUNCH.launch:60 */
  int substeps;        /* Simulator.f steps per controller tick */
  int warmup_ticks;    /* ticks linearised around predicted_vectors_generation (controllerMain.py:310-320) */
  int swap_ey_epsi;    /* 1: controllerMain.py:188 slot assignment (ey -> epsi slot, epsi -> ey slot) */
  int reserved;
  double vel_ref;      /* controllerMain.py:326 */
  double Cf_new;       /* controllerMain.py:77 */
  double half_width;   /* Map.halfWidth */
  double slack;        /* Map.slack */
  double sim_mu;       /* simulator/mu, MAIN_LAUNCH.launch:72 */
} loop_ref_cfg;

/* vehicleSimulator.py:164-199.  st = [x y yaw vx vy psiDot ax ay]; u = [motor, servo] (vehicleSimulator.py:337). */
void loop_ref_sim_f(double *st, const double *u, const lpv_ref_vehicle *v, double mu, double dt);

/* trackInitialization.py:283-383.  out = [s ey epsi]; returns CompletedFlag (0: the 10000 sentinels). */
int loop_ref_local_position(const double *track, int nseg, double half_width, double slack, double x, double y,
                            double psi, double *out);

/* trackInitialization.py:205-260.  out = [x y theta]; returns 1 when no unique segment holds s. */
int loop_ref_global_position(const double *track, int nseg, double s, double ey, double *out);

/* controllerMain.py:510-553, first N rows: xx [N,6], uu [N,2] */
void loop_ref_guess(const double *local, int N, double *xx, double *uu);

/* n_ticks closed-loop ticks for B vehicles (OpenMP over vehicles).  Per tick (controllerMain.py:177-454, lap 0):
 * measure (true state, vx >= 0.01) -> getLocalPosition -> lap logic -> OldSteering/OldAccelera <- cmd ->
 * warm-up (_EstimateABC around the guess) or LPVPrediction -> solve -> cmd = uPred[0] -> `substeps` x Simulator.f.
 * Arrays (row-major, updated in place): sim [B,8], cmd [B,2], u_pred [B,N,2], local [B,6],
 * ctr [B,8] ints = [first_it, lap, half_track, last_status, last_iters, fail_status (0 = alive), fail_tick, ticks_done],
 * stat [B,4] doubles = [solved ticks, total ADMM iterations, max |ey| (true), tick of the first lap completion or -1].
 * x_pred [B,N+1,6] optional (last tick).  Returns the number of SOLVED ticks. */
long loop_ref_run(const lpv_ref_cfg *c, const osqp_ref_settings *st, const loop_ref_cfg *lc, int B, int n_ticks,
                  double *sim, double *cmd, double *u_pred, double *local, int *ctr, double *stat, double *x_pred,
                  int threads);

/* ---- planner loop (plannerMain.py:128-224, Testing-style: re-plan from the previous plan's second state) ---- */
/* plannerMain.py:465-505 : xx [N+1,6] = [Vx Vy W Ey Epsi S], uu = zeros; s0 = 0 in the reference */
void plan_loop_ref_guess(const double *x0, int N, double accel_rate, double dt, double s0, double *xx);

/* n_ticks planner ticks for B vehicles.  Tick: first tick _EstimateABC around the guess (plannerMain.py:152-164), later
 * LPVPrediction(xPred[1], SS, uPred) + solve(xPred[1], ...) (:175-176); uOld = 0 (OldSteering is never popped, :186-187);
 * then the arc-length integration SS[j+1] = SS[j] + ((vx cos(epsi) - vy sin(epsi)) / (1 - ey kappa(SS[j]))) dt over the
 * new plan and SS[0] = SS[1] (:201-211).
 * Arrays (row-major, in place): xstart [B,5], x_pred [B,N+1,5], u_pred [B,N,2], SS [B,N+1] (SS[:,0] = start arc length),
 * ctr [B,8] = [ticks_done, last_status, last_iters, fail_status, fail_tick, 0, 0, 0], stat [B,4] = [solved ticks,
 * total ADMM iterations, 0, 0].  Returns the number of SOLVED ticks. */
long plan_loop_ref_run(const lpv_ref_cfg *c, const osqp_ref_settings *st, int B, int n_ticks, double max_ey,
                       double accel_rate, const double *xstart, double *x_pred, double *u_pred, double *SS, int *ctr,
                       double *stat, int threads);

/* ---- controller <- planner hand-off (controllerMain.py:196-243, Body_Frame_Errors :495-506) ---- */
/* One vehicle: g [6] = vx vy wz X Y psi (GlobalState), refs [5, n_ref] = x_d y_d psi_d vx_d curv_d, window offset
 * `index`; out: x0 [6] = LocalState [vx vy wz epsi s ey], vel_ref [N+1] (entry N = entry N-1: the reference's
 * vel_ref[-1]), curv_ref [N]; returns ex. */
double track_inputs_ref(const double *g, int lap, double s_prev, const double *refs, int n_ref, int index, int N, double dt,
                        double *x0, double *vel_ref, double *curv_ref);

#ifdef __cplusplus
}
#endif
#endif
