/*
 * lpvmpc.h — C-ABI of the B200-native batched LPV-MPC QP solver.
 *
 * Drop-in boundary for ONE hot path of euge2838/Autonomous-Racing-LPV-MPP-MPC: per control tick,
 * LPV scheduling A(rho_k)/B(rho_k) over the horizon -> QP assembly -> OSQP ADMM solve, for a batch
 * of independent problems.  The reference has no FFI of its own for this path (it is Python calling
 * the PyPI `osqp` C extension); the interfaces each entry point replaces are the Python methods below
 * (paths relative to /root/reference/workspace/src/barc/src):
 *
 *   lpvmpc_create            PathFollowingLPV_MPC.__init__   ControllerObject/PathFollowingLPVMPC.py:35-84
 *                            LPV_MPC_Planner.__init__        PlannerObject/LPV_MPC_Planner.py:34-82
 *                            (+ rospy params MAIN_LAUNCH.launch:5-11,40-44 and Map().PointAndTangent,
 *                             Utilities/trackInitialization.py:90-202)
 *   lpvmpc_schedule_*        .LPVPrediction                  PathFollowingLPVMPC.py:166-258, LPV_MPC_Planner.py:242-320
 *                            _EstimateABC                    PathFollowingLPVMPC.py:732-809, LPV_MPC_Planner.py:519-591
 *   lpvmpc_solve_*           .solve + osqp_solve_qp          PathFollowingLPVMPC.py:89-162,273-325
 *                            .solve (OSQP().setup/.solve)    LPV_MPC_Planner.py:86-236
 *                            (OSQP itself: PyPI `osqp` 0.6.x, setup+solve+polish; not vendored)
 *   lpvmpc_loop_*            the controller node's main loop + the simulator node, for a fleet (see below)
 *                            controllerMain.py:177-454, vehicleSimulator.py:164-199,
 *                            Utilities/trackInitialization.py:283-383
 *
 * Conventions: plain pointers and sizes only; every call returns 0 or a negative LPVMPC_E_* code and
 * never throws; lpvmpc_last_error() gives the message.  A handle is bound to one CUDA device and is not
 * thread-safe; `_dev` calls take DEVICE pointers and are asynchronous on `stream` (a cudaStream_t passed
 * as void*, NULL = legacy default stream); `_host` calls take HOST pointers, stage through pinned
 * buffers owned by the handle, and return after the results are in the caller's arrays.
 * All arrays are row-major fp64 unless stated.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with LPVMPC_E_CUDA.
 *
 * Sizes: controller n=6 states [vx vy wz epsi s ey], planner n=5 states [vx vy wz ey epsi]; d=2 inputs
 * [delta a]; nz = n(N+1)+dN decision variables ordered [x_0..x_N, u_0..u_{N-1}] as in the reference
 * (PathFollowingLPVMPC.py:157-158); m constraint rows in the REFERENCE's OSQP row order:
 *   controller: 2N state rows, 4N input rows, n(N+1) dynamics rows, then `steering_delay` rows
 *               (PathFollowingLPVMPC.py:304-308, 329-378, 518-527)
 *   planner   : n(N+1) dynamics rows, then nz box rows (LPV_MPC_Planner.py:200-202)
 */
#ifndef LPVMPC_H
#define LPVMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPVMPC_ABI_VERSION 1

enum { LPVMPC_CONTROLLER = 0, LPVMPC_PLANNER = 1 };

/* error codes (function return values) */
enum {
  LPVMPC_OK = 0,
  LPVMPC_E_ARG = -1,      /* bad argument / NULL where required / batch > max_batch */
  LPVMPC_E_CUDA = -2,     /* CUDA runtime error or no device */
  LPVMPC_E_UNSUPPORTED = -3
};

/* per-problem status: OSQP's own integers (osqp 0.6 constants.h) plus two of ours */
enum {
  LPVMPC_SOLVED = 1,
  LPVMPC_SOLVED_INACCURATE = 2,
  LPVMPC_PRIMAL_INFEASIBLE_INACCURATE = 3,
  LPVMPC_DUAL_INFEASIBLE_INACCURATE = 4,
  LPVMPC_MAX_ITER_REACHED = -2,
  LPVMPC_PRIMAL_INFEASIBLE = -3,
  LPVMPC_DUAL_INFEASIBLE = -4,
  LPVMPC_NON_CVX = -7,
  LPVMPC_UNSOLVED = -10,
  LPVMPC_SCHEDULE_ERROR = -20, /* Curvature(s) had no (unique) segment: the reference raises (utilities.py:46) */
  LPVMPC_DATA_ERROR = -21,     /* l > u: upstream osqp.setup() refuses the problem */
  LPVMPC_OFF_TRACK = -22       /* closed loop: getLocalPosition found no segment (the reference returns 10000 sentinels) */
};

/* scheduling modes of lpvmpc_solve_* */
enum {
  LPVMPC_SCHED_GIVEN = 0,   /* A,B(,C) supplied by the caller (what .solve(..., A_L,B_L,C_L, first_it>=10) does) */
  LPVMPC_SCHED_PREDICT = 1, /* fused .LPVPrediction roll-out inside the solve kernel */
  LPVMPC_SCHED_ESTIMATE = 2 /* fused _EstimateABC from a given trajectory (warm-up ticks) */
};

typedef struct {
  double rho, sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double delta;                  /* polish regularisation */
  double adaptive_rho_tolerance;
  int32_t max_iter;
  int32_t check_termination;     /* 0 = never check before max_iter (fixed-iteration mode) */
  int32_t scaling;               /* Ruiz passes */
  int32_t adaptive_rho;
  int32_t adaptive_rho_interval; /* 0 = OSQP's deterministic fallback: 4 * check_termination (or 100) */
  int32_t polish;
  int32_t polish_refine_iter;
  int32_t scaled_termination;
} lpvmpc_settings;

typedef struct {
  int32_t abi_version;           /* LPVMPC_ABI_VERSION */
  int32_t kind;                  /* LPVMPC_CONTROLLER / LPVMPC_PLANNER */
  int32_t N;                     /* horizon */
  int32_t steering_delay;        /* controller only; extra equality rows u_k[0] = old_steering[k+1] */
  double dt;
  double Q[36];                  /* n x n row-major (planner uses the leading 25 entries) */
  double R[4];                   /* d x d */
  double dR[2];                  /* input-rate weights */
  double L_cf[5];                /* planner linear (lap-time) cost */
  double lf, lr, m, Iz, Cf, Cr, mu; /* rospy params "lf".."mu" */
  double max_vel, min_vel;       /* /TrajectoryPlanner/max_vel, min_vel */
  int32_t n_track_seg;           /* rows of PointAndTangent */
  int32_t max_batch;             /* workspaces are allocated for this many problems */
  const double *track;           /* HOST pointer, n_track_seg x 6 [x y psi s len kappa]; copied */
  int32_t device;                /* CUDA device ordinal */
  int32_t variant;               /* 0 = auto; 1 = generic warp-per-QP kernel; see lpvmpc_info.variant */
  lpvmpc_settings settings;
} lpvmpc_cfg;

typedef struct {
  int32_t n, d, N, nz, m;        /* problem dimensions */
  int32_t variant;               /* kernel actually selected */
  int32_t workspace_in_smem;     /* 1: per-QP workspace lives in shared memory; 0: global (L2) workspace */
  int32_t smem_bytes_per_qp;
  int64_t workspace_bytes;       /* device bytes owned by the handle */
  int64_t kernel_launches;       /* kernels launched by this handle so far */
} lpvmpc_info;

/* Batched arguments of one solve.  Unused pointers may be NULL.  `_dev`: device pointers; `_host`: host. */
typedef struct {
  int32_t sched_mode;            /* LPVMPC_SCHED_* */
  int32_t x0_from_prediction;    /* controller lap-0 quirk: QP x0 := first predicted state (controllerMain.py:329-331) */
  int32_t lap_all;               /* used when `lap` is NULL (controller PREDICT: 0 => Curvature(s), else curv_ref) */
  int32_t natural_order;         /* 0 (default): the kernel may visit the problems of a batch in its own order (controller:
                                    grouped by vel_ref[0] - vx0, which predicts the ADMM iteration count, so that the QPs
                                    sharing a warp finish together; results do not depend on it); 1: batch order */
  double Cf_new;                 /* controller PREDICT tyre stiffness (controllerMain.py:77: 60) */
  /* inputs */
  const double *x0;              /* [B,n]      QP initial state (and scheduling state unless x_sched given) */
  const double *x_sched;         /* [B,n]      optional roll-out start (PREDICT) */
  const double *A;               /* [B,N,n,n]  GIVEN */
  const double *Bm;              /* [B,N,n,d]  GIVEN */
  const double *C;               /* [B,N,n]    GIVEN, optional (reference: zeros) */
  const double *u_prev;          /* [B,N,d]    PREDICT / ESTIMATE: previous input sequence (steering schedules A,B) */
  const double *vel_ref;         /* [B,N+1]    controller: tracking reference; entry N is the terminal one */
  const double *curv_ref;        /* [B,N]      controller PREDICT with lap != 0 */
  const double *SS;              /* [B,N+1]    planner PREDICT: arc-length per stage */
  const int32_t *lap;            /* [B]        controller PREDICT, optional */
  const double *traj;            /* [B,N,6]    ESTIMATE: controller rows [vx vy wz epsi s ey]; planner [vx vy w ey epsi s] */
  const double *u_old;           /* [B,d]      slew-rate anchor [OldSteering[0], OldAccelera[0]]; NULL = 0 */
  const double *old_steering;    /* [B,delay]  controller steering_delay rows: OldSteering[1..delay] */
  const double *max_ey;          /* [B]        planner lateral box */
  const double *ey_lo, *ey_hi;   /* [B,N+1]    planner optional per-stage lateral box ("obstacles") */
  /* outputs */
  double *x_pred;                /* [B,N+1,n] */
  double *u_pred;                /* [B,N,d]   */
  int32_t *status;               /* [B] LPVMPC_* status */
  int32_t *iters;                /* [B] ADMM iterations */
  int32_t *rho_updates;          /* [B] optional */
  int32_t *polish_status;        /* [B] optional: 1 ok, -1 unsuccessful, 0 not run */
  double *obj;                   /* [B] optional objective 0.5 z'Pz + q'z */
  double *pri_res, *dua_res;     /* [B] optional */
  uint8_t *active_lo, *active_up;/* [B,m] optional polish active-set guess, reference row order */
  double *y;                     /* [B,m] optional dual solution, reference row order */
  double *A_out, *B_out;         /* [B,N,n,n], [B,N,n,d] optional: scheduled matrices (PREDICT/ESTIMATE) */
  double *states_out;            /* [B,N,n] optional: roll-out states (PREDICT) */
  double *xs, *zs, *ys;          /* [B,nz],[B,m],[B,m] optional scaled ADMM iterates before polish (parity tests) */
  const int32_t *order_hint;     /* [B] optional input: expected ADMM iterations per problem (e.g. `iters` of the same problems
                                    at the previous tick); used instead of vel_ref[0] - vx0 to group the batch */
} lpvmpc_args;

typedef struct lpvmpc_handle lpvmpc_handle;

int lpvmpc_abi_version(void);
void lpvmpc_default_settings(lpvmpc_settings *s); /* OSQP 0.6 defaults + polish=1 (what the reference runs) */
int lpvmpc_device_count(void);

int lpvmpc_create(const lpvmpc_cfg *cfg, lpvmpc_handle **out);
void lpvmpc_destroy(lpvmpc_handle *h);
const char *lpvmpc_last_error(const lpvmpc_handle *h); /* h may be NULL: last create() error */
int lpvmpc_get_info(const lpvmpc_handle *h, lpvmpc_info *info);
int lpvmpc_update_settings(lpvmpc_handle *h, const lpvmpc_settings *s);

/* LPVPrediction / _EstimateABC for a batch: A_out [B,N,n,n], B_out [B,N,n,d], states_out [B,N,n] (PREDICT),
 * sched_err [B] (1 where Curvature failed).  Uses x0 (or x_sched), u_prev, vel_ref, curv_ref/SS, lap, traj. */
int lpvmpc_schedule_dev(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, int32_t *sched_err, void *stream);
int lpvmpc_schedule_host(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, int32_t *sched_err);

/* schedule (per sched_mode) + build + OSQP solve (+ polish) + unpack */
int lpvmpc_solve_dev(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, void *stream);
int lpvmpc_solve_host(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a);
/* As lpvmpc_solve_host, but the results are NOT copied into the caller's arrays: every output requested in `a` (non-NULL
 * pointer; its value is not used) comes back in `views` as a pointer into pinned host memory owned by the handle -- one of
 * two result arenas used in turn, so a result stays valid until the call after the next one on this handle.  Removes the
 * last host copy of the call (2.4 MB per 4,096 controller QPs); the kernel has written the arena directly. */
int lpvmpc_solve_host_view(lpvmpc_handle *h, int32_t B, const lpvmpc_args *a, lpvmpc_args *views);

/* ------------------------------------------------------------------------------------------------
 * Closed-loop fleet (BASELINE configs[3], SURVEY 8b `lpvmpc_step_closed_loop`): B independent vehicles, each running the
 * controller main loop of the reference in path-tracking mode (lap 0) against the reference simulator's vehicle
 * model, entirely on the device.  One tick =
 *     measure the true state, vx >= 0.01                       controllerMain.py:179-184
 *     Map.getLocalPosition(x, y, psi) -> s, ey, epsi           controllerMain.py:188, trackInitialization.py:283-383
 *     lap bookkeeping                                          controllerMain.py:190-192, 252-257
 *     OldSteering / OldAccelera <- last command                controllerMain.py:289-298
 *     warm-up ticks: _EstimateABC around the hard-coded guess  controllerMain.py:310-315, 510-553
 *     later ticks : LPVPrediction, x0 = first predicted state  controllerMain.py:326-331
 *     solve (schedule + build + OSQP + polish)                 = lpvmpc_solve_dev on the same handle
 *     command = uPred[0]                                       controllerMain.py:381-383
 *     `substeps` x Simulator.f(u = [motor, servo])             vehicleSimulator.py:164-199, 337
 * Differences from the reference, on purpose: a vehicle whose QP is not feasible ({SOLVED, SOLVED_INACCURATE,
 * MAX_ITER_REACHED}, PathFollowingLPVMPC.py:322-324) or that leaves the track is retired (ctr[5] = the status, its
 * state frozen) instead of carrying on with res.x / the 10000 sentinels; the state estimator is bypassed (true-state
 * feedback; its gain tables are not in the reference repository).  The handle must be a controller with N <= 20 and
 * steering_delay = 0.
 */
typedef struct {
  double sim_dt;          /* simulator/dt (MAIN_LAUNCH.launch:60): 0.005 */
  int32_t substeps;       /* Simulator.f steps per controller tick (33 ms / 5 ms, rounded up: 7) */
  int32_t warmup_ticks;   /* first_it < 10 (controllerMain.py:310): 9 */
  int32_t swap_ey_epsi;   /* 1 = controllerMain.py:188 as written (ey lands in the epsi slot and vice versa); 0 = as its comment intends */
  int32_t reserved;
  double vel_ref;         /* controllerMain.py:326: 1.0 */
  double Cf_new;          /* controllerMain.py:77: 60 */
  double half_width;      /* Map.halfWidth (trackInitialization.py:20) */
  double slack;           /* Map.slack (trackInitialization.py:30,45,54) */
  double sim_mu;          /* simulator/mu (MAIN_LAUNCH.launch:72) */
} lpvmpc_loop_cfg;

/* Fleet state, row-major.  ctr = [first_it, lap, half_track, last_status, last_iters, fail_status (0 = running),
 * fail_tick, ticks_done]; stat = [SOLVED ticks, total ADMM iterations, max |ey| (true), tick of the first lap
 * completion or -1]. */
typedef struct {
  double *sim;      /* [B,8]     x y yaw vx vy psiDot ax ay */
  double *cmd;      /* [B,2]     last command [delta a] */
  double *u_pred;   /* [B,N,2]   last predicted inputs */
  double *x_pred;   /* [B,N+1,6] last predicted states */
  double *local;    /* [B,6]     last measured state in the controller's slots [vx vy wz epsi s ey] */
  double *stat;     /* [B,4] */
  int32_t *ctr;     /* [B,8] */
} lpvmpc_loop_state;

void lpvmpc_loop_default_cfg(lpvmpc_loop_cfg *c); /* launch-file values, L_shape track widths */
/* (Re)start a fleet of B <= max_batch vehicles from simulator states sim0 [B,8] (first_it = 1, command 0). */
int lpvmpc_loop_init_dev(lpvmpc_handle *h, int32_t B, const lpvmpc_loop_cfg *c, const double *sim0, void *stream);
int lpvmpc_loop_init_host(lpvmpc_handle *h, int32_t B, const lpvmpc_loop_cfg *c, const double *sim0);
/* Enqueue n_ticks ticks (2 kernels per tick + 1; nothing returns to the host).  _host runs on the handle's stream
 * and waits. */
int lpvmpc_loop_run_dev(lpvmpc_handle *h, int32_t n_ticks, void *stream);
int lpvmpc_loop_run_host(lpvmpc_handle *h, int32_t n_ticks);
/* Device pointers of the fleet state (owned by the handle, valid until the next init / destroy). */
int lpvmpc_loop_view_dev(lpvmpc_handle *h, lpvmpc_loop_state *view, int32_t *B);
/* Copy the non-NULL members of `dst` (HOST pointers) out of the device state. */
int lpvmpc_loop_read_host(lpvmpc_handle *h, const lpvmpc_loop_state *dst);

/* ------------------------------------------------------------------------------------------------
 * Planner loop for a fleet (plannerMain.py:128-224): every plan is re-planned from its own second
 * predicted state.  One tick =
 *     first tick : _EstimateABC around predicted_vectors_generation(N, x0, accel_rate, dt)   plannerMain.py:152-164, 465-505
 *     later ticks: LPVPrediction(xPred[1], SS, uPred); solve(xPred[1], ...)                    plannerMain.py:175-176
 *     uOld = 0 (OldSteering / OldAccelera are appended but never popped)                     plannerMain.py:186-187
 *     SS[j+1] = SS[j] + ((vx cos(epsi) - vy sin(epsi)) / (1 - ey kappa(SS[j]))) dt over the new plan, SS[0] = SS[1]
 *                                                                                            plannerMain.py:201-211
 * A plan whose QP is not feasible, or whose arc length leaves Curvature()'s domain, is retired (ctr[3] = the status).
 * The handle must be a planner.  The reference starts at s = 0 with [1, 0, 0, 0, 0] (Testing mode, :145-146); here every
 * plan has its own start state and start arc length.
 * State, row-major: x_pred [B,N+1,5], u_pred [B,N,2], SS [B,N+1], ctr [B,8] = [ticks_done, last_status, last_iters,
 * fail_status (0 = running), fail_tick, 0, 0, 0], stat [B,4] = [SOLVED ticks, total ADMM iterations, 0, 0].
 */
typedef struct {
  double *x_pred, *u_pred, *SS, *stat;
  int32_t *ctr;
} lpvmpc_plan_loop_state;

/* xstart [B,5] = [vx vy wz ey epsi], s0 [B] start arc lengths (NULL = 0), max_ey = lateral box (plannerMain.py:52 HW),
 * accel_rate = 0.2 in the reference (plannerMain.py:158). */
int lpvmpc_plan_loop_init_host(lpvmpc_handle *h, int32_t B, const double *xstart, const double *s0, double max_ey, double accel_rate);
int lpvmpc_plan_loop_init_dev(lpvmpc_handle *h, int32_t B, const double *xstart, const double *s0, double max_ey, double accel_rate,
                              void *stream);
int lpvmpc_plan_loop_run_host(lpvmpc_handle *h, int32_t n_ticks);
int lpvmpc_plan_loop_run_dev(lpvmpc_handle *h, int32_t n_ticks, void *stream);
int lpvmpc_plan_loop_view_dev(lpvmpc_handle *h, lpvmpc_plan_loop_state *view, int32_t *B);
int lpvmpc_plan_loop_read_host(lpvmpc_handle *h, const lpvmpc_plan_loop_state *dst);

/* ------------------------------------------------------------------------------------------------
 * Planner -> controller references (plannerMain.py:196-224, 257-280, My_Planning of :299-303): for every
 * plan the stage poses Xref/Yref/Thetaref = Map.getGlobalPosition(SS[j], 0) (stage 0: xyth0 = Xlast, Ylast, Thetalast),
 * yaw = Thetaref + epsi, xp = Xref - ey sin(yaw), yp = Yref + ey cos(yaw), vel = vx, curv = wz / vx over the first N
 * stages; then interp1d(kind='cubic') from the planner grid to n_out samples and, for the curvature, filtfilt with the
 * 4th-order elliptic filter.  Both are linear for the uniform grids involved; the caller supplies them as matrices W
 * (x, y, yaw, vx) and Wc (curvature), [n_out, N] row-major (postproc.py builds them).  Planner handles only.
 * refs [B,5,n_out] = x_d, y_d, psi_d, vx_d, curv_d; err [B] optional (1: getGlobalPosition found no segment).
 */
int lpvmpc_plan_refs_setup(lpvmpc_handle *h, int32_t n_out, const double *W, const double *Wc); /* HOST matrices; copied */
int lpvmpc_plan_refs_dev(lpvmpc_handle *h, int32_t B, const double *x_pred, const double *SS, const double *xyth0, double *refs,
                         int32_t *err, void *stream);
int lpvmpc_plan_refs_host(lpvmpc_handle *h, int32_t B, const double *x_pred, const double *SS, const double *xyth0, double *refs,
                          int32_t *err);

/* ------------------------------------------------------------------------------------------------
 * Controller <- planner hand-off: the trajectory-tracking branch of the controller node's main
 * loop, controllerMain.py:196-243, with Body_Frame_Errors (controllerMain.py:495-506).  For every vehicle: psi =
 * wrap(psi - 2 pi lap); the window [index, index + N) of the planner's references (controllerMain.py:229-233; the
 * reference's own index state machine is ReferenceWindow in handoff.py); ex, ey, epsi against the first reference pose
 * of the window and s = s_prev + (vx cos(epsi) - vy sin(epsi)) / (1 - ey curv) dt.  Outputs are the inputs of
 * lpvmpc_solve_* with LPVMPC_SCHED_PREDICT and lap != 0 (controllerMain.py:361-363): x0 = LocalState [vx vy wz epsi s ey],
 * vel_ref (entry N repeats entry N-1: the reference hands N samples and its cost uses vel_ref[-1] for the terminal
 * stage), curv_ref.  Controller handles only.
 *   gstate [B,6] vx vy wz X Y psi (GlobalState); lap [B] optional; s_prev [B]; refs [B,5,n_ref] = x_d, y_d, psi_d, vx_d,
 *   curv_d (the layout lpvmpc_plan_refs_* writes); index [B] optional (NULL = 0), `_dev`: index_max = upper bound of its
 *   entries (n_ref >= N + index_max is checked; `_host` checks the entries themselves); ex [B] optional.
 */
int lpvmpc_track_inputs_dev(lpvmpc_handle *h, int32_t B, const double *gstate, const int32_t *lap, const double *s_prev, const double *refs,
                            int32_t n_ref, const int32_t *index, int32_t index_max, double *x0, double *vel_ref, double *curv_ref, double *ex,
                            void *stream);
int lpvmpc_track_inputs_host(lpvmpc_handle *h, int32_t B, const double *gstate, const int32_t *lap, const double *s_prev, const double *refs,
                             int32_t n_ref, const int32_t *index, double *x0, double *vel_ref, double *curv_ref, double *ex);

/* ------------------------------------------------------------------------------------------------
 * two auxiliary routines of the reference next to the hot path, batched, one thread per element.  Their
 * vertex / gain tables are the caller's: the reference loads them from .mat files that are not in its repository.  Any
 * handle serves (it supplies the device, the stream of the `_host` variants and their staging arena: n is limited by it).
 *
 * lpvmpc_anfis_abc_*: the TS-fuzzy (ANFIS) scheduling blend ABC_computation_5SV_new,
 * ControllerObject/PathFollowingLPVMPC.py:530-602.  sched [n,5] = vx vy omega steer accel; bell [10,3] = (a, b, c) of the
 * two generalised-bell memberships 1 / (1 + |(x - c) / a|^(2b)) of every scheduling variable (:551-552); vertex v of 32 has
 * bits (vx, vy, omega, steer, accel) = bits 4..0 (:555-592); outputs = normalised-weight blends of the vertex tables
 * A_tab [32,3], B_tab [32,2], C_tab [32] (:597-601): A [n,3], B [n,2], C [n].
 */
int lpvmpc_anfis_abc_dev(lpvmpc_handle *h, int32_t n, const double *sched, const double *A_tab, const double *B_tab, const double *C_tab,
                         const double *bell, double *A, double *B, double *C, void *stream);
int lpvmpc_anfis_abc_host(lpvmpc_handle *h, int32_t n, const double *sched, const double *A_tab, const double *B_tab, const double *C_tab,
                          const double *bell, double *A, double *B, double *C);
/*
 * lpvmpc_observer_step_*: one step of the polytopic LPV observer, stateEstimator.py:349-492 (GS_LPV_Est with
 * Continuous_AB_Comp :392-430 and L_Gain_Comp :433-492):  est <- est + dt (A_obs + L C_obs) est + dt B_obs u - dt L y.
 * est [n,6] = vx vy omega x y yaw (in place); y [n,5] = vx omega x y yaw; u [n,2] = steer accel; lim_* [6,2]
 * (SchedVars_Limits) and gains_* [6,5,16] (Llmi) of the low- and the high-speed polytope (the high-speed one when the
 * scheduling vx > lim_ls[0][1]); C_obs [5,6]; use_est [n] (or NULL: use_all for every element): schedule on the
 * estimate (the reference's curr_time > 0.02 branch, :369-373) or on the measurement (:374-378).
 */
int lpvmpc_observer_step_dev(lpvmpc_handle *h, int32_t n, double *est, const double *y, const double *u, const double *lim_ls,
                             const double *gains_ls, const double *lim_hs, const double *gains_hs, const double *C_obs, double dt,
                             const int32_t *use_est, int32_t use_all, void *stream);
int lpvmpc_observer_step_host(lpvmpc_handle *h, int32_t n, double *est, const double *y, const double *u, const double *lim_ls,
                              const double *gains_ls, const double *lim_hs, const double *gains_hs, const double *C_obs, double dt,
                              const int32_t *use_est, int32_t use_all);

#ifdef __cplusplus
}
#endif
#endif /* LPVMPC_H */
