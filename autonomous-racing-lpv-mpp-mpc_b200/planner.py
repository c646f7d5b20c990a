"""Drop-in for the reference's planner object (PlannerObject/LPV_MPC_Planner.py:29-320) on the B200 library."""
import datetime

import numpy as np

from . import _native as nat
from .controller import _ros_params
from .solver import BatchSolver


class LPV_MPC_Planner(object):
    """``LPV_MPC_Planner(Q, R, dR, L_cf, N, dt, map, Solver)`` with ``.LPVPrediction`` and ``.solve``."""

    def __init__(self, Q, R, dR, L_cf, N, dt, map, Solver, params=None, device=0, **osqp_settings):
        self.A, self.B, self.C = [], [], []
        self.N = N
        self.nx = Q.shape[0]
        self.nu = R.shape[0]
        self.Q, self.QN, self.R, self.dR, self.L_cf = Q, Q, R, dR, L_cf
        self.LinPoints = np.zeros((self.N + 2, self.nx))
        self.dt = dt
        self.map = map
        self.halfWidth = map.halfWidth
        self.first_it = 1
        self.Solver = Solver
        self.steeringDelay = 0
        self.OldSteering = [0.0] * int(1)
        self.OldAccelera = [0.0] * int(1)
        prm = _ros_params(["lf", "lr", "m", "Iz", "Cf", "Cr", "mu", "/TrajectoryPlanner/max_vel",
                           "/TrajectoryPlanner/min_vel"], params)
        self.lf, self.lr, self.m, self.I = prm["lf"], prm["lr"], prm["m"], prm["Iz"]
        self.Cf, self.Cr, self.mu = prm["Cf"], prm["Cr"], prm["mu"]
        self.g = 9.81
        self.epss = 0.00000001
        self.max_vel = prm["/TrajectoryPlanner/max_vel"]
        self.min_vel = prm["/TrajectoryPlanner/min_vel"]
        self.feasible = 1
        self.status_val = None
        self.info = {}
        veh = dict(lf=self.lf, lr=self.lr, m=self.m, Iz=self.I, Cf=self.Cf, Cr=self.Cr, mu=self.mu)
        self._solver = BatchSolver("planner", N, dt, Q, R, dR, map.PointAndTangent, L_cf=L_cf, vehicle=veh,
                                   max_vel=self.max_vel, min_vel=self.min_vel, max_batch=1, device=device, **osqp_settings)

    # LPV_MPC_Planner.py:86-236
    def solve(self, x0, Last_xPredicted, uPred, A_LPV, B_LPV, C_LPV, first_it, max_ey):
        startTimer = datetime.datetime.now()
        N, n, d = self.N, self.nx, self.nu
        kw = dict(u_old=np.array([[self.OldSteering[0], self.OldAccelera[0]]], dtype=np.float64),
                  max_ey=np.array([float(max_ey)]))
        if first_it < 2:
            mode = nat.SCHED_ESTIMATE
            traj = np.asarray(Last_xPredicted, dtype=np.float64)[:N, :6]
            uu = np.asarray(uPred, dtype=np.float64)
            steer = uu.reshape(uu.shape[0], -1)[:N, 0]  # (Hp,1) or (Hp,) or (Hp,2): steering column
            up = np.zeros((1, N, d))
            up[0, :, 0] = steer
            kw["traj"], kw["u_prev"] = traj[None], up
            extra = ("A_out", "B_out")
        else:
            mode = nat.SCHED_GIVEN
            kw["A"] = np.asarray(A_LPV, dtype=np.float64).reshape(1, N, n, n)
            kw["Bm"] = np.asarray(B_LPV, dtype=np.float64).reshape(1, N, n, d)
            kw["C"] = np.asarray(C_LPV, dtype=np.float64).reshape(1, N, n)
            self.A, self.B, self.C = A_LPV, B_LPV, C_LPV
            extra = ()
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
        self.feasible = 1 if status in (1, 2, -2) else 0
        if self.feasible == 0:
            print('QUIT...')
        self.solverTime = datetime.datetime.now() - startTimer
        self.xPred = res.x_pred[0].copy()
        self.uPred = res.u_pred[0].copy()
        self.LinPoints = np.concatenate((self.xPred[1:, :], np.array([self.xPred[-1, :]])), axis=0)

    # LPV_MPC_Planner.py:242-320
    def LPVPrediction(self, x, SS, u):
        N, n, d = self.N, self.nx, self.nu
        SSv = np.zeros((1, N + 1))
        SSa = np.asarray(SS, dtype=np.float64).reshape(-1)
        SSv[0, :min(N + 1, SSa.size)] = SSa[:N + 1]
        res = self._solver.schedule(sched_mode=nat.SCHED_PREDICT, x0=np.asarray(x, dtype=np.float64).reshape(1, n),
                                    u_prev=np.asarray(u, dtype=np.float64)[None, :N, :d], SS=SSv)
        if int(res.sched_err[0]):
            raise TypeError("only length-1 arrays can be converted to Python scalars")
        Atv = [res.A_out[0, k].copy() for k in range(N)]
        Btv = [res.B_out[0, k].copy() for k in range(N)]
        Ctv = [np.zeros((n, 1)) for _ in range(N)]
        return res.states_out[0].copy(), Atv, Btv, Ctv
