"""ClosedLoopFleet: B independent vehicles running the reference's controller main loop (path tracking, lap 0)
against the reference simulator's vehicle model, entirely on one B200 (BASELINE configs[3]).

Host-side mirror of what ``controllerMain.py:177-454`` + ``vehicleSimulator.py:164-199`` do for ONE car, over the
``lpvmpc_loop_*`` entry points of the C-ABI: per tick localise (``Map.getLocalPosition``), schedule, build, solve,
apply the command to the simulator — nothing crosses PCIe until ``read()``.
"""
import ctypes as C

import numpy as np

from . import _native as nat
from .solver import BatchSolver
from .track import Map
from .postproc import n_out_for, reference_matrices
from .workloads import CTRL_DT, CTRL_PT, PLAN, PLAN_DT

CTR_FIELDS = ("first_it", "lap", "half_track", "status", "iters", "fail_status", "fail_tick", "ticks")
STAT_FIELDS = ("solved_ticks", "admm_iterations", "max_abs_ey", "lap_tick")


def global_position(track_map, s, ey):
    """Vectorised ``Map.getGlobalPosition`` (trackInitialization.py:205-260) for 0 <= s < TrackLength: (x, y, theta)."""
    pt = track_map.PointAndTangent
    s = np.asarray(s, dtype=np.float64)
    ey = np.asarray(ey, dtype=np.float64)
    i = np.clip(np.searchsorted(pt[:, 3], s, side="right") - 1, 0, pt.shape[0] - 1)
    cur, prev = pt[i], pt[i - 1]
    straight = cur[:, 5] == 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = (s - cur[:, 3]) / cur[:, 4]
        xs_ = (1 - rel) * prev[:, 0] + rel * cur[:, 0] + ey * np.cos(cur[:, 2] + np.pi / 2)
        ys_ = (1 - rel) * prev[:, 1] + rel * cur[:, 1] + ey * np.sin(cur[:, 2] + np.pi / 2)
        r = 1.0 / cur[:, 5]
        ang = prev[:, 2]
        direction = np.where(r >= 0, 1.0, -1.0)
        cx = prev[:, 0] + np.abs(r) * np.cos(ang + direction * np.pi / 2)
        cy = prev[:, 1] + np.abs(r) * np.sin(ang + direction * np.pi / 2)
        span = (s - cur[:, 3]) / (np.pi * np.abs(r)) * np.pi
        normal = direction * np.pi / 2 + ang
        normal = np.where(normal < -np.pi, 2 * np.pi + normal, np.where(normal > np.pi, normal - 2 * np.pi, normal))
        a0 = -(np.pi - np.abs(normal)) * np.where(normal >= 0, 1.0, -1.0)
        xc = cx + (np.abs(r) - direction * ey) * np.cos(a0 + direction * span)
        yc = cy + (np.abs(r) - direction * ey) * np.sin(a0 + direction * span)
        th = ang + direction * span
    return np.where(straight, xs_, xc), np.where(straight, ys_, yc), np.where(straight, cur[:, 2], th)


def fleet_start(B, seed=2, track_map=None):
    """SURVEY 8d config 4 start states: s0 ~ U(0, 19.2), ey0 ~ U(-.1, .1), vx0 ~ U(.2, 1); sim rows [x y yaw vx vy psiDot ax ay]."""
    m = track_map if track_map is not None else Map("L_shape")
    rng = np.random.default_rng(seed)
    s0 = rng.uniform(0.0, 19.2, B)
    ey0 = rng.uniform(-0.1, 0.1, B)
    vx0 = rng.uniform(0.2, 1.0, B)
    x, y, th = global_position(m, s0, ey0)
    sim = np.zeros((B, 8))
    sim[:, 0], sim[:, 1], sim[:, 2], sim[:, 3] = x, y, th, vx0
    return sim


class ClosedLoopFleet(object):
    def __init__(self, track_map=None, N=8, dt=CTRL_DT, tune=None, max_fleet=8192, device=0, substeps=7,
                 swap_ey_epsi=1, warmup_ticks=9, vel_ref=1.0, Cf_new=60.0, sim_dt=0.005, sim_mu=0.05, variant=0,
                 **osqp_settings):
        self.map = track_map if track_map is not None else Map("L_shape")
        tune = dict(CTRL_PT if tune is None else tune)
        self.solver = BatchSolver("controller", N, dt, track=self.map.PointAndTangent, max_batch=max_fleet, device=device,
                                  variant=variant, **tune, **osqp_settings)
        self.N = int(N)
        self.device = int(device)
        lc = nat.LoopCfg()
        nat.lib().lpvmpc_loop_default_cfg(C.byref(lc))
        lc.substeps, lc.swap_ey_epsi, lc.warmup_ticks = int(substeps), int(swap_ey_epsi), int(warmup_ticks)
        lc.vel_ref, lc.Cf_new, lc.sim_dt, lc.sim_mu = float(vel_ref), float(Cf_new), float(sim_dt), float(sim_mu)
        lc.half_width, lc.slack = float(self.map.halfWidth), float(self.map.slack)
        self.cfg = lc
        self.B = 0

    @property
    def _h(self):
        return self.solver._h

    def close(self):
        self.solver.close()

    def start(self, sim0):
        """(Re)start the fleet from simulator states ``sim0`` [B,8] (numpy: staged H2D; torch CUDA tensor: device copy)."""
        L = nat.lib()
        if type(sim0).__module__.startswith("torch"):
            import torch
            t = sim0.to(device=torch.device("cuda", self.device), dtype=torch.float64).contiguous()
            self.B = int(t.shape[0])
            stream = torch.cuda.current_stream(self.device).cuda_stream
            nat.check(L.lpvmpc_loop_init_dev(self._h, self.B, C.byref(self.cfg), C.c_void_p(t.data_ptr()), C.c_void_p(stream)), self._h)
        else:
            a = np.ascontiguousarray(sim0, dtype=np.float64)
            if a.ndim != 2 or a.shape[1] != 8:
                raise ValueError("sim0 must be [B,8] = [x y yaw vx vy psiDot ax ay]")
            self.B = int(a.shape[0])
            nat.check(L.lpvmpc_loop_init_host(self._h, self.B, C.byref(self.cfg), C.c_void_p(a.ctypes.data)), self._h)
        return self

    def run(self, n_ticks, stream=None):
        """Advance every vehicle by ``n_ticks`` controller ticks.  ``stream`` None: the handle's stream, waits;
        otherwise a CUDA stream handle (int): asynchronous."""
        L = nat.lib()
        if stream is None:
            nat.check(L.lpvmpc_loop_run_host(self._h, int(n_ticks)), self._h)
        else:
            nat.check(L.lpvmpc_loop_run_dev(self._h, int(n_ticks), C.c_void_p(int(stream))), self._h)
        return self

    def read(self, fields=("sim", "cmd", "u_pred", "x_pred", "local", "stat", "ctr")):
        """Copy the fleet state to host numpy arrays (dict)."""
        B, N = self.B, self.N
        shapes = dict(sim=((B, 8), "f8"), cmd=((B, 2), "f8"), u_pred=((B, N, 2), "f8"), x_pred=((B, N + 1, 6), "f8"),
                      local=((B, 6), "f8"), stat=((B, 4), "f8"), ctr=((B, 8), "i4"))
        out = {}
        st = nat.LoopState()
        for k in fields:
            shp, dt = shapes[k]
            out[k] = np.empty(shp, dtype=dt)
            setattr(st, k, out[k].ctypes.data)
        nat.check(nat.lib().lpvmpc_loop_read_host(self._h, C.byref(st)), self._h)
        return out

    def view(self):
        """Device pointers (ints) of the fleet state, for callers that keep everything on the GPU."""
        st = nat.LoopState()
        B = C.c_int32(0)
        nat.check(nat.lib().lpvmpc_loop_view_dev(self._h, C.byref(st), C.byref(B)), self._h)
        return {k: getattr(st, k) for k, _ in nat.LoopState._fields_}, B.value


class PlannerFleet(object):
    """B independent plans advanced by the reference's planner main loop (plannerMain.py:128-224) on one B200: every
    tick re-plans from the previous plan's second state (``LPVPrediction(xPred[1], SS, uPred)`` + ``solve``) and
    integrates the arc lengths ``SS`` over the new plan; the first tick linearises around
    ``predicted_vectors_generation`` (plannerMain.py:465-505).  ``lpvmpc_plan_loop_*`` of the C-ABI."""

    def __init__(self, track_map=None, N=40, dt=PLAN_DT, tune=None, max_fleet=4096, device=0, max_ey=0.2, accel_rate=0.2,
                 variant=0, **osqp_settings):
        self.map = track_map if track_map is not None else Map("L_shape")
        tune = dict(PLAN if tune is None else tune)
        self.solver = BatchSolver("planner", N, dt, track=self.map.PointAndTangent, max_batch=max_fleet, device=device,
                                  variant=variant, **tune, **osqp_settings)
        self.N, self.device = int(N), int(device)
        self.max_ey, self.accel_rate = float(max_ey), float(accel_rate)
        self.B = 0

    @property
    def _h(self):
        return self.solver._h

    def close(self):
        self.solver.close()

    def start(self, xstart, s0=None):
        """xstart [B,5] = [vx vy wz ey epsi]; s0 [B] start arc lengths (None: 0, the reference's Testing mode)."""
        x = np.ascontiguousarray(xstart, dtype=np.float64)
        if x.ndim != 2 or x.shape[1] != 5:
            raise ValueError("xstart must be [B,5]")
        self.B = int(x.shape[0])
        s = None if s0 is None else np.ascontiguousarray(s0, dtype=np.float64).reshape(self.B)
        nat.check(nat.lib().lpvmpc_plan_loop_init_host(self._h, self.B, C.c_void_p(x.ctypes.data),
                                                       C.c_void_p(s.ctypes.data) if s is not None else None,
                                                       self.max_ey, self.accel_rate), self._h)
        return self

    def run(self, n_ticks, stream=None):
        L = nat.lib()
        if stream is None:
            nat.check(L.lpvmpc_plan_loop_run_host(self._h, int(n_ticks)), self._h)
        else:
            nat.check(L.lpvmpc_plan_loop_run_dev(self._h, int(n_ticks), C.c_void_p(int(stream))), self._h)
        return self

    def read(self, fields=("x_pred", "u_pred", "SS", "stat", "ctr")):
        B, N = self.B, self.N
        shapes = dict(x_pred=((B, N + 1, 5), "f8"), u_pred=((B, N, 2), "f8"), SS=((B, N + 1), "f8"), stat=((B, 4), "f8"),
                      ctr=((B, 8), "i4"))
        out = {}
        st = nat.PlanLoopState()
        for k in fields:
            shp, dt = shapes[k]
            out[k] = np.empty(shp, dtype=dt)
            setattr(st, k, out[k].ctypes.data)
        nat.check(nat.lib().lpvmpc_plan_loop_read_host(self._h, C.byref(st)), self._h)
        return out

    def references(self, x_pred, SS, xyth0):
        """Planner -> controller references of ``My_Planning`` (plannerMain.py:196-224, 257-303) for a batch of plans:
        x_pred [B,N+1,5], SS [B,N+1] (arc lengths of the plan), xyth0 [B,3] (pose of stage 0: Xlast, Ylast, Thetalast).
        Returns (refs [B,5,n_out] = x_d, y_d, psi_d, vx_d, curv_d; err [B])."""
        L = nat.lib()
        if not getattr(self, "_refs_ready", False):
            W, Wc = reference_matrices(self.N, self.solver._cfg.dt)
            self.n_out = W.shape[0]
            nat.check(L.lpvmpc_plan_refs_setup(self._h, self.n_out, C.c_void_p(W.ctypes.data), C.c_void_p(Wc.ctypes.data)), self._h)
            self._refs_ready = True
        x = np.ascontiguousarray(x_pred, dtype=np.float64)
        s_ = np.ascontiguousarray(SS, dtype=np.float64)
        p0 = np.ascontiguousarray(xyth0, dtype=np.float64)
        B = int(x.shape[0])
        refs = np.empty((B, 5, self.n_out))
        err = np.zeros(B, dtype=np.int32)
        nat.check(L.lpvmpc_plan_refs_host(self._h, B, C.c_void_p(x.ctypes.data), C.c_void_p(s_.ctypes.data), C.c_void_p(p0.ctypes.data),
                                          C.c_void_p(refs.ctypes.data), C.c_void_p(err.ctypes.data)), self._h)
        return refs, err
