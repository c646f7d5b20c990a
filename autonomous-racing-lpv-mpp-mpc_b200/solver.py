"""BatchSolver: the batched schedule -> build -> OSQP-ADMM path on one B200, over the C-ABI.

Host arrays (numpy) go through ``lpvmpc_solve_host`` (pinned staging, one H2D, one kernel, one D2H);
CUDA tensors (torch) go through ``lpvmpc_solve_dev`` on torch's current stream with no copies.  PyTorch is
only the owner of device memory and streams here.
"""
import ctypes as C

import numpy as np

from . import _native as nat

VEHICLE_DEFAULTS = dict(lf=0.125, lr=0.125, m=1.98, Iz=0.03, Cf=60.0, Cr=60.0, mu=0.05)  # MAIN_LAUNCH.launch:5-11

_F64_IN = ("x0", "x_sched", "A", "Bm", "C", "u_prev", "vel_ref", "curv_ref", "SS", "traj", "u_old", "old_steering",
           "max_ey", "ey_lo", "ey_hi")


class BatchResult(dict):
    """dict with attribute access: x_pred, u_pred, status, iters, ... (numpy arrays or torch tensors)."""
    __getattr__ = dict.__getitem__


class BatchSolver(object):
    """One handle = one problem family (kind, horizon, weights, bounds, track, OSQP settings) on one GPU."""

    def __init__(self, kind, N, dt, Q, R, dR, track, L_cf=None, vehicle=None, max_vel=5.0, min_vel=0.9,
                 steering_delay=0, max_batch=4096, device=0, settings=None, variant=0, **osqp_settings):
        self.kind = nat.CONTROLLER if kind in ("controller", nat.CONTROLLER) else nat.PLANNER
        self.n = 6 if self.kind == nat.CONTROLLER else 5
        self.d = 2
        self.N = int(N)
        n = self.n
        Q = np.asarray(Q, dtype=np.float64)
        R = np.asarray(R, dtype=np.float64)
        if Q.shape != (n, n) or R.shape != (2, 2):
            raise ValueError("Q must be %dx%d and R 2x2" % (n, n))
        cfg = nat.Cfg()
        cfg.abi_version = nat.ABI_VERSION
        cfg.kind = self.kind
        cfg.N = self.N
        cfg.steering_delay = int(steering_delay)
        cfg.dt = float(dt)
        Qf = np.zeros(36)
        Qf[:n * n] = Q.reshape(-1)
        cfg.Q[:] = Qf.tolist()
        cfg.R[:] = R.reshape(-1).tolist()
        cfg.dR[:] = np.asarray(dR, dtype=np.float64).reshape(2).tolist()
        cfg.L_cf[:] = (np.zeros(5) if L_cf is None else np.asarray(L_cf, dtype=np.float64).reshape(5)).tolist()
        veh = dict(VEHICLE_DEFAULTS)
        veh.update(vehicle or {})
        for k, v in veh.items():
            setattr(cfg, k, float(v))
        cfg.max_vel = float(max_vel)
        cfg.min_vel = float(min_vel)
        self._track = np.ascontiguousarray(track, dtype=np.float64)
        if self._track.ndim != 2 or self._track.shape[1] != 6:
            raise ValueError("track must be the [S,6] PointAndTangent table")
        cfg.n_track_seg = self._track.shape[0]
        cfg.track = self._track.ctypes.data_as(nat.c_double_p)
        cfg.max_batch = int(max_batch)
        cfg.device = int(device)
        cfg.variant = int(variant)  # 0 auto, 1 generic warp-per-QP kernel, 2 T8 kernel (controller, N=8)
        cfg.settings = settings if settings is not None else nat.default_settings(**osqp_settings)
        self.steering_delay = int(steering_delay)
        self.max_batch = int(max_batch)
        self.device = int(device)
        self._h = C.c_void_p()
        L = nat.lib()
        rc = L.lpvmpc_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            nat.check(rc, None)
        self._cfg = cfg
        info = self.info()
        self.nz, self.m = info["nz"], info["m"]

    # ------------------------------------------------------------------ housekeeping
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            nat.lib().lpvmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = nat.Info()
        nat.check(nat.lib().lpvmpc_get_info(self._h, C.byref(i)), self._h)
        return {k: getattr(i, k) for k, _ in nat.Info._fields_}

    def update_settings(self, **kw):
        s = self._cfg.settings
        for k, v in kw.items():
            if not hasattr(s, k):
                raise KeyError(k)
            setattr(s, k, v)
        nat.check(nat.lib().lpvmpc_update_settings(self._h, C.byref(s)), self._h)

    # ------------------------------------------------------------------ argument plumbing
    def _shapes(self, B):
        n, d, N, nz, m = self.n, self.d, self.N, self.nz, self.m
        return dict(x0=(B, n), x_sched=(B, n), A=(B, N, n, n), Bm=(B, N, n, d), C=(B, N, n), u_prev=(B, N, d),
                    vel_ref=(B, N + 1), curv_ref=(B, N), SS=(B, N + 1), lap=(B,), traj=(B, N, 6), u_old=(B, d),
                    old_steering=(B, max(self.steering_delay, 1)), max_ey=(B,), ey_lo=(B, N + 1), ey_hi=(B, N + 1), order_hint=(B,),
                    x_pred=(B, N + 1, n), u_pred=(B, N, d), status=(B,), iters=(B,), rho_updates=(B,),
                    polish_status=(B,), obj=(B,), pri_res=(B,), dua_res=(B,), active_lo=(B, m), active_up=(B, m),
                    y=(B, m), A_out=(B, N, n, n), B_out=(B, N, n, d), states_out=(B, N, n), xs=(B, nz), zs=(B, m),
                    ys=(B, m))

    _OUT_DTYPES = dict(x_pred="f8", u_pred="f8", status="i4", iters="i4", rho_updates="i4", polish_status="i4",
                       obj="f8", pri_res="f8", dua_res="f8", active_lo="u1", active_up="u1", y="f8", A_out="f8",
                       B_out="f8", states_out="f8", xs="f8", zs="f8", ys="f8")
    _DEFAULT_OUT = ("x_pred", "u_pred", "status", "iters", "rho_updates", "polish_status", "obj", "pri_res", "dua_res")

    def _is_torch(self, v):
        return type(v).__module__.startswith("torch")

    def _prepare(self, inputs, outputs, sched_mode, x0_from_prediction, lap_all, Cf_new):
        import numpy as _np
        B = None
        use_torch = any(self._is_torch(v) for v in inputs.values() if v is not None)
        for k in ("x0", "x_sched", "traj", "u_prev"):
            if inputs.get(k) is not None:
                B = int(inputs[k].shape[0])
                break
        if B is None:
            raise ValueError("cannot infer the batch size: give x0")
        shapes = self._shapes(B)
        a = nat.Args()
        a.sched_mode = int(sched_mode)
        a.x0_from_prediction = int(bool(x0_from_prediction))
        a.lap_all = int(lap_all)
        a.Cf_new = float(Cf_new)
        keep = []
        if use_torch:
            import torch
            dev = torch.device("cuda", self.device)
        for k, v in inputs.items():
            if v is None:
                continue
            if k not in shapes:
                raise KeyError("unknown input %r" % k)
            want = "i4" if k in ("lap", "order_hint") else "f8"
            if use_torch:
                tdt = torch.int32 if want == "i4" else torch.float64
                t = v if self._is_torch(v) else torch.as_tensor(_np.asarray(v))
                t = t.to(device=dev, dtype=tdt).contiguous()
                if tuple(t.shape) != shapes[k]:
                    t = t.reshape(shapes[k])
                keep.append(t)
                setattr(a, k, t.data_ptr())
            else:
                arr = _np.ascontiguousarray(v, dtype=want)
                if arr.shape != shapes[k]:
                    arr = arr.reshape(shapes[k])
                keep.append(arr)
                setattr(a, k, arr.ctypes.data)
        res = BatchResult()
        for k in outputs:
            dt = self._OUT_DTYPES[k]
            if use_torch:
                tdt = {"f8": torch.float64, "i4": torch.int32, "u1": torch.uint8}[dt]
                t = torch.empty(shapes[k], dtype=tdt, device=dev)
                res[k] = t
                setattr(a, k, t.data_ptr())
            else:
                arr = _np.empty(shapes[k], dtype=dt)
                res[k] = arr
                setattr(a, k, arr.ctypes.data)
        return B, a, res, keep, use_torch

    # ------------------------------------------------------------------ the hot path
    def solve(self, x0, sched_mode=nat.SCHED_PREDICT, extra_outputs=(), x0_from_prediction=False, lap_all=1,
              Cf_new=60.0, host_views=False, **inputs):
        """Schedule (per ``sched_mode``) + build + OSQP solve for a batch.

        numpy inputs -> host path, torch CUDA tensors -> device path (asynchronous on torch's current stream).
        Inputs by keyword as in ``lpvmpc_args``: A, Bm, C | u_prev, vel_ref, curv_ref, SS, lap, x_sched | traj;
        u_old, old_steering, max_ey, ey_lo, ey_hi.

        ``host_views=True`` (host path): the result arrays are numpy VIEWS of the handle's pinned result arena
        (``lpvmpc_solve_host_view``: no copy into fresh arrays); they stay valid until the call after the next one.
        """
        inputs = dict(inputs)
        inputs["x0"] = x0
        outs = tuple(self._DEFAULT_OUT) + tuple(o for o in extra_outputs if o not in self._DEFAULT_OUT)
        if host_views and not any(self._is_torch(v) for v in inputs.values() if v is not None):
            return self._solve_views(inputs, outs, sched_mode, x0_from_prediction, lap_all, Cf_new)
        B, a, res, keep, use_torch = self._prepare(inputs, outs, sched_mode, x0_from_prediction, lap_all, Cf_new)
        L = nat.lib()
        if use_torch:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
            nat.check(L.lpvmpc_solve_dev(self._h, B, C.byref(a), C.c_void_p(stream)), self._h)
            res["_keepalive"] = keep
        else:
            nat.check(L.lpvmpc_solve_host(self._h, B, C.byref(a)), self._h)
        return res

    def _solve_views(self, inputs, outs, sched_mode, x0_from_prediction, lap_all, Cf_new):
        import numpy as _np
        B, a, res, keep, _ = self._prepare(inputs, (), sched_mode, x0_from_prediction, lap_all, Cf_new)
        shapes = self._shapes(B)
        for k in outs:
            setattr(a, k, 1)                      # requested (the pointer's value is not used)
        views = nat.Args()
        nat.check(nat.lib().lpvmpc_solve_host_view(self._h, B, C.byref(a), C.byref(views)), self._h)
        ctype = {"f8": C.c_double, "i4": C.c_int32, "u1": C.c_uint8}
        for k in outs:
            n = int(_np.prod(shapes[k]))
            ptr = getattr(views, k)
            if n == 0 or not ptr:
                res[k] = _np.empty(shapes[k], dtype=self._OUT_DTYPES[k])
                continue
            res[k] = _np.ctypeslib.as_array((ctype[self._OUT_DTYPES[k]] * n).from_address(ptr)).reshape(shapes[k])
        return res

    def schedule(self, sched_mode=nat.SCHED_PREDICT, lap_all=1, Cf_new=60.0, **inputs):
        """Batched LPVPrediction (PREDICT) / _EstimateABC (ESTIMATE): returns A_out, B_out, states_out, sched_err."""
        outs = ("A_out", "B_out") + (("states_out",) if sched_mode == nat.SCHED_PREDICT else ())
        B, a, res, keep, use_torch = self._prepare(inputs, outs, sched_mode, False, lap_all, Cf_new)
        L = nat.lib()
        if use_torch:
            import torch
            err = torch.empty(B, dtype=torch.int32, device=torch.device("cuda", self.device))   # every entry is written
            stream = torch.cuda.current_stream(self.device).cuda_stream
            nat.check(L.lpvmpc_schedule_dev(self._h, B, C.byref(a), C.c_void_p(err.data_ptr()), C.c_void_p(stream)), self._h)
            res["_keepalive"] = keep
        else:
            err = np.zeros(B, dtype=np.int32)
            nat.check(L.lpvmpc_schedule_host(self._h, B, C.byref(a), C.c_void_p(err.ctypes.data)), self._h)
        res["sched_err"] = err
        return res
