/* TEST INFRASTRUCTURE (CPU oracle) -- not part of the product path.
 * Restatement of two auxiliary routines of the reference, pinned by tests/golden/aux.npz, which the
 * reference's own Python text produced (tests/golden/make_golden_aux.py):
 *   anfis_abc_ref      ControllerObject/PathFollowingLPVMPC.py:530-602  ABC_computation_5SV_new
 *   observer_step_ref  stateEstimator.py:349-492  GS_LPV_Est + Continuous_AB_Comp + L_Gain_Comp
 */
#ifndef AUX_REF_H
#define AUX_REF_H
#ifdef __cplusplus
extern "C" {
#endif

/* sched [5] = vx vy omega steer accel; A_tab [32,3], B_tab [32,2], C_tab [32], bell [10,3] = (a, b, c) of the two
 * generalised-bell memberships of every scheduling variable -> A [3], B [2], returns C */
double anfis_abc_ref(const double *sched, const double *A_tab, const double *B_tab, const double *C_tab, const double *bell,
                     double *A, double *B);

/* One observer step: est [6] = vx vy omega x y yaw (in place), y [5] = vx omega x y yaw, u [2] = steer accel;
 * lim_* [6,2], gains_* [6,5,16] (the low- / high-speed polytope), C_obs [5,6]; use_est: schedule on the estimate (1: the
 * reference's curr_time > 0.02 branch) or on the measurement (0) */
void observer_step_ref(double *est, const double *y, const double *u, const double *lim_ls, const double *gains_ls,
                       const double *lim_hs, const double *gains_hs, const double *C_obs, double dt, int use_est);

#ifdef __cplusplus
}
#endif
#endif
