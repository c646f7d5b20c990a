"""Drop-in for the reference's controller object, backed by the B200 library.

Same constructor, methods, attributes and quirks as ``PathFollowingLPV_MPC``
(ControllerObject/PathFollowingLPVMPC.py:30-258); the work happens in liblpvmpc.so.
"""
import datetime

import numpy as np

from . import _native as nat
from .solver import BatchSolver, VEHICLE_DEFAULTS


def _ros_params(keys, overrides):
    """rospy.get_param when a ROS master is there (as the reference does), else launch-file values."""
    out = {}
    try:
        import rospy  # noqa: F401
        getter = rospy.get_param
    except Exception:
        getter = None
    defaults = dict(VEHICLE_DEFAULTS)
    defaults.update({"/TrajectoryPlanner/max_vel": 5.0, "/TrajectoryPlanner/min_vel": 0.9})
    for k in keys:
        if overrides and k in overrides:
            out[k] = overrides[k]
        elif getter is not None:
            try:
                out[k] = getter(k)
            except Exception:
                out[k] = defaults[k]
        else:
            out[k] = defaults[k]
    return out


class PathFollowingLPV_MPC(object):
    """``PathFollowingLPV_MPC(Q, R, dR, N, vt, dt, map, Solver, steeringDelay, velocityDelay)``.

    Extra keyword-only knobs (not in the reference): ``params`` (dict replacing rospy.get_param),
    ``device``, and OSQP settings (the reference runs the library defaults with polish=True).
    """

    def __init__(self, Q, R, dR, N, vt, dt, map, Solver, steeringDelay, velocityDelay, params=None, device=0,
                 **osqp_settings):
        prm = _ros_params(["lf", "lr", "m", "Iz", "Cf", "Cr", "mu", "/TrajectoryPlanner/max_vel"], params)
        self.lf, self.lr, self.m, self.I = prm["lf"], prm["lr"], prm["m"], prm["Iz"]
        self.Cf, self.Cr, self.mu = prm["Cf"], prm["Cr"], prm["mu"]
        self.g = 9.81
        self.max_vel = prm["/TrajectoryPlanner/max_vel"]
        self.A, self.B, self.C = [], [], []
        self.N = N
        self.n = Q.shape[0]
        self.d = R.shape[0]
        self.vt = vt
        self.Q, self.R, self.dR = Q, R, dR
        self.LinPoints = np.zeros((self.N + 2, self.n))
        self.dt = dt
        self.map = map
        self.halfWidth = map.halfWidth
        self.first_it = 1
        self.steeringDelay = steeringDelay
        self.velocityDelay = velocityDelay
        self.OldSteering = [0.0] * int(1 + steeringDelay)
        self.OldAccelera = [0.0] * int(1)
        self.OldPredicted = [0.0] * int(1 + steeringDelay + N)
        self.Solver = Solver
        if Solver != "OSQP":
            raise NotImplementedError("only the OSQP path of the reference is implemented (the CVX branch is dead code)")
        self.feasible = 1
        self.status_val = None
        self.info = {}
        veh = dict(lf=self.lf, lr=self.lr, m=self.m, Iz=self.I, Cf=self.Cf, Cr=self.Cr, mu=self.mu)
        self._solver = BatchSolver("controller", N, dt, Q, R, dR, map.PointAndTangent, vehicle=veh, max_vel=self.max_vel,
                                   steering_delay=int(steeringDelay), max_batch=1, device=device, **osqp_settings)

    # ------------------------------------------------------------------ .solve (PathFollowingLPVMPC.py:89-162)
    def solve(self, x0, Last_xPredicted, uPred, NN_LPV_MPC, vel_ref, A_L, B_L, C_L, first_it):
        startTimer = datetime.datetime.now()
        N, n, d = self.N, self.n, self.d
        vel_ref = np.atleast_1d(np.asarray(vel_ref, dtype=np.float64))
        vref = np.empty((1, N + 1))
        vref[0, :N] = vel_ref[:N]
        vref[0, N] = vel_ref[-1]
        kw = dict(vel_ref=vref, u_old=np.array([[self.OldSteering[0], self.OldAccelera[0]]], dtype=np.float64))
        if self.steeringDelay > 0:
            kw["old_steering"] = np.asarray(self.OldSteering[1:1 + int(self.steeringDelay)], dtype=np.float64)[None, :]
        if (NN_LPV_MPC == False) and (first_it < 10):  # noqa: E712  (reference spelling)
            mode = nat.SCHED_ESTIMATE
            kw["traj"] = np.asarray(Last_xPredicted, dtype=np.float64)[None, :N, :6]
            kw["u_prev"] = np.asarray(uPred, dtype=np.float64)[None, :N, :d]
            extra = ("A_out", "B_out")
        else:
            mode = nat.SCHED_GIVEN
            kw["A"] = np.asarray(A_L, dtype=np.float64).reshape(1, N, n, n)
            kw["Bm"] = np.asarray(B_L, dtype=np.float64).reshape(1, N, n, d)
            kw["C"] = np.asarray(C_L, dtype=np.float64).reshape(1, N, n)
            self.A, self.B, self.C = A_L, B_L, C_L
            extra = ()
        self.linearizationTime = datetime.datetime.now() - startTimer
        startTimer = datetime.datetime.now()
        res = self._solver.solve(np.asarray(x0, dtype=np.float64).reshape(1, n), sched_mode=mode,
                                 extra_outputs=extra + ("active_lo", "active_up"), **kw)
        if mode == nat.SCHED_ESTIMATE:
            self.A = [res.A_out[0, k] for k in range(N)]
            self.B = [res.B_out[0, k] for k in range(N)]
            self.C = [np.zeros((n, 1)) for _ in range(N)]
        status = int(res.status[0])
        if status == -20:
            raise TypeError("only length-1 arrays can be converted to Python scalars")  # Curvature() failure
        self.status_val = status
        self.info = {k: res[k][0] for k in ("iters", "rho_updates", "polish_status", "obj", "pri_res", "dua_res")}
        self.active_lo, self.active_up = res.active_lo[0], res.active_up[0]
        if status != 1:
            print("OSQP exited with status '%s'" % nat.STATUS_NAMES.get(status, status))
        self.feasible = 1 if status in (1, 2, -2) else 0
        if self.feasible == 0:
            print('QUIT...')
        self.solverTime = datetime.datetime.now() - startTimer
        self.xPred = res.x_pred[0].copy()
        self.uPred = res.u_pred[0].copy()
        self.LinPoints = np.concatenate((self.xPred[1:, :], np.array([self.xPred[-1, :]])), axis=0)

    # ------------------------------------------------------------------ .LPVPrediction (:166-258)
    def LPVPrediction(self, x, u, vel_ref, curv_ref, Cf_new, LapNumber):
        N, n, d = self.N, self.n, self.d
        vel_ref = np.atleast_1d(np.asarray(vel_ref, dtype=np.float64))
        vref = np.zeros((1, N + 1))
        vref[0, :N] = vel_ref[:N]
        vref[0, N] = vel_ref[-1]
        kw = dict(x0=np.asarray(x, dtype=np.float64).reshape(1, n), u_prev=np.asarray(u, dtype=np.float64)[None, :N, :d],
                  vel_ref=vref, lap=np.array([int(LapNumber)], dtype=np.int32))
        kw["curv_ref"] = np.zeros((1, N)) if LapNumber == 0 else np.asarray(curv_ref, dtype=np.float64).reshape(-1)[None, :N]
        res = self._solver.schedule(sched_mode=nat.SCHED_PREDICT, Cf_new=float(Cf_new), **kw)
        if int(res.sched_err[0]):
            raise TypeError("only length-1 arrays can be converted to Python scalars")  # utilities.py:46
        Atv = [res.A_out[0, k].copy() for k in range(N)]
        Btv = [res.B_out[0, k].copy() for k in range(N)]
        Ctv = [np.zeros((n, 1)) for _ in range(N)]
        return res.states_out[0].copy(), Atv, Btv, Ctv
