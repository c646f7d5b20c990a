"""SURVEY.md 8f row 4 over the C-ABI: the TS-fuzzy (ANFIS) scheduling blend and the polytopic LPV observer, batched.

* ``anfis_abc``      = ``ABC_computation_5SV_new`` (ControllerObject/PathFollowingLPVMPC.py:530-602) for n operating points;
* ``observer_step``  = ``GS_LPV_Est`` with ``Continuous_AB_Comp`` and ``L_Gain_Comp`` (stateEstimator.py:349-492) for n vehicles.

numpy in -> numpy out (``lpvmpc_*_host``), torch CUDA tensors in -> torch CUDA tensors out on torch's current stream
(``lpvmpc_*_dev``; the small tables may be numpy either way).  The vertex / gain tables are the caller's: the reference
reads them from .mat files that are not part of its repository.
"""
import ctypes as C

import numpy as np

from . import _native as nat


def _is_torch(v):
    return type(v).__module__.startswith("torch")


def _np(a, shape, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), a.shape))
    return a


def anfis_abc(solver, sched, A_tab, B_tab, C_tab, bell):
    """sched [n,5] = vx vy omega steer accel -> {"A": [n,3], "B": [n,2], "C": [n]}."""
    L = nat.lib()
    A_tab, B_tab, C_tab, bell = _np(A_tab, (32, 3), "A_tab"), _np(B_tab, (32, 2), "B_tab"), _np(C_tab, (32,), "C_tab"), _np(bell, (10, 3), "bell")
    if _is_torch(sched):
        import torch
        dev = torch.device("cuda", solver.device)
        sd = sched.to(device=dev, dtype=torch.float64).contiguous().reshape(-1, 5)
        n = int(sd.shape[0])
        tabs = torch.as_tensor(np.concatenate([A_tab.ravel(), B_tab.ravel(), C_tab.ravel(), bell.ravel()])).to(dev)
        out = {"A": torch.empty((n, 3), dtype=torch.float64, device=dev), "B": torch.empty((n, 2), dtype=torch.float64, device=dev),
               "C": torch.empty((n,), dtype=torch.float64, device=dev)}
        p = tabs.data_ptr()
        nat.check(L.lpvmpc_anfis_abc_dev(solver._h, n, sd.data_ptr(), p, p + 96 * 8, p + 160 * 8, p + 192 * 8, out["A"].data_ptr(),
                                         out["B"].data_ptr(), out["C"].data_ptr(), C.c_void_p(torch.cuda.current_stream(solver.device).cuda_stream)),
                  solver._h)
        out["_keepalive"] = (sd, tabs)
        return out
    sd = np.ascontiguousarray(sched, dtype=np.float64).reshape(-1, 5)
    n = int(sd.shape[0])
    out = {"A": np.empty((n, 3)), "B": np.empty((n, 2)), "C": np.empty((n,))}
    nat.check(L.lpvmpc_anfis_abc_host(solver._h, n, sd.ctypes.data, A_tab.ctypes.data, B_tab.ctypes.data, C_tab.ctypes.data, bell.ctypes.data,
                                      out["A"].ctypes.data, out["B"].ctypes.data, out["C"].ctypes.data), solver._h)
    return out


def observer_step(solver, est, y, u, lim_ls, gains_ls, lim_hs, gains_hs, C_obs, dt, use_estimate=1):
    """One observer step for n vehicles: est [n,6] (vx vy omega x y yaw), y [n,5] (vx omega x y yaw), u [n,2] (steer accel)
    -> new est [n,6].  ``use_estimate``: scalar or [n]; nonzero = schedule on the estimate (the reference once
    curr_time > 0.02 s), zero = on the measurement."""
    L = nat.lib()
    lim_ls, lim_hs = _np(lim_ls, (6, 2), "lim_ls"), _np(lim_hs, (6, 2), "lim_hs")
    gains_ls, gains_hs = _np(gains_ls, (6, 5, 16), "gains_ls"), _np(gains_hs, (6, 5, 16), "gains_hs")
    C_obs = _np(C_obs, (5, 6), "C_obs")
    scalar = np.ndim(use_estimate) == 0 and not _is_torch(use_estimate)
    if _is_torch(est):
        import torch
        dev = torch.device("cuda", solver.device)
        e = est.to(device=dev, dtype=torch.float64).contiguous().reshape(-1, 6).clone()
        n = int(e.shape[0])
        yd = y.to(device=dev, dtype=torch.float64).contiguous().reshape(n, 5) if _is_torch(y) else torch.as_tensor(_np(y, (n, 5), "y")).to(dev)
        ud = u.to(device=dev, dtype=torch.float64).contiguous().reshape(n, 2) if _is_torch(u) else torch.as_tensor(_np(u, (n, 2), "u")).to(dev)
        tabs = torch.as_tensor(np.concatenate([lim_ls.ravel(), gains_ls.ravel(), lim_hs.ravel(), gains_hs.ravel(), C_obs.ravel()])).to(dev)
        use = None if scalar else torch.as_tensor(np.asarray(use_estimate.cpu() if _is_torch(use_estimate) else use_estimate)).to(device=dev, dtype=torch.int32).contiguous()
        p = tabs.data_ptr()
        nat.check(L.lpvmpc_observer_step_dev(solver._h, n, e.data_ptr(), yd.data_ptr(), ud.data_ptr(), p, p + 12 * 8, p + 492 * 8, p + 504 * 8,
                                             p + 984 * 8, float(dt), None if use is None else use.data_ptr(), int(bool(use_estimate)) if scalar else 0,
                                             C.c_void_p(torch.cuda.current_stream(solver.device).cuda_stream)), solver._h)
        e._keepalive = (yd, ud, tabs, use)   # the launch is asynchronous: its operands live as long as the result
        return e
    e = np.array(est, dtype=np.float64).reshape(-1, 6).copy()
    n = int(e.shape[0])
    yd, ud = _np(y, (n, 5), "y"), _np(u, (n, 2), "u")
    use = None if scalar else np.ascontiguousarray(use_estimate, dtype=np.int32).reshape(n)
    nat.check(L.lpvmpc_observer_step_host(solver._h, n, e.ctypes.data, yd.ctypes.data, ud.ctypes.data, lim_ls.ctypes.data, gains_ls.ctypes.data,
                                          lim_hs.ctypes.data, gains_hs.ctypes.data, C_obs.ctypes.data, float(dt),
                                          None if use is None else use.ctypes.data, int(bool(use_estimate)) if scalar else 0), solver._h)
    return e
