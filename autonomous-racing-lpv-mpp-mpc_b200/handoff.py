"""Controller <- planner hand-off: the trajectory-tracking branch of the controller node's main loop
(controllerMain.py:196-243) for a batch of vehicles.

* ``ReferenceWindow`` is the reference's own index state machine around the planner message (controllerMain.py:
  217-235): with ``max_window = 0`` (the value in the file) a fresh window is latched on every second tick and the
  tick in between keeps the previous one.
* ``track_inputs`` runs ``lpv_track_inputs_kernel`` over the C-ABI (``lpvmpc_track_inputs_*``): yaw unwrapping by lap,
  ``Body_Frame_Errors`` (controllerMain.py:495-506) against the first pose of the window, the arc-length update and the
  ``vel_ref`` / ``curv_ref`` windows, i.e. the inputs of ``BatchSolver.solve(..., lap_all=1)`` (controllerMain.py:361-363).

numpy arrays go through the ``_host`` entry point, torch CUDA tensors through ``_dev`` on torch's current stream.
"""
import ctypes as C

import numpy as np

from . import _native as nat


class ReferenceWindow(object):
    """controllerMain.py:225-235 as written.  ``update(x_d, y_d, psi_d, vx_d, curv_d)`` is called once per tick with
    the newest planner message and returns ``(x_ref, y_ref, yaw_ref, vel_ref, curv_ref)`` (N samples each) — on the
    ticks where the reference's ``else: index = 0`` branch runs these are the PREVIOUS tick's windows, exactly as the
    reference leaves its variables untouched."""

    def __init__(self, N, max_window=0):
        self.N = int(N)
        self.max_window = int(max_window)
        self.index = 0
        self._vec = None
        self._cur = None

    def update(self, x_d, y_d, psi_d, vx_d, curv_d):
        N, mw = self.N, self.max_window
        if self.index <= mw:
            if self.index == 0:
                self._vec = tuple(np.array(v[0:N + mw], dtype=np.float64) for v in (x_d, y_d, psi_d, vx_d, curv_d))
            i = self.index
            self._cur = tuple(v[i:i + N] for v in self._vec)
            self.index += 1
        else:
            self.index = 0
        if self._cur is None:
            raise RuntimeError("no reference window latched yet")
        return self._cur


def _ptr(v):
    if v is None:
        return C.c_void_p()
    if type(v).__module__.startswith("torch"):
        return C.c_void_p(v.data_ptr())
    return C.c_void_p(v.ctypes.data)


def track_inputs(solver, gstate, s_prev, refs, lap=None, index=None):
    """Batched hand-off for a controller ``BatchSolver``.

    gstate [B,6] = vx vy wz X Y psi; s_prev [B]; refs [B,5,n_ref] = x_d y_d psi_d vx_d curv_d (what
    ``PlannerFleet.references`` returns); lap [B] int32 optional; index [B] int32 optional window offsets.
    Returns dict(x0 [B,6], vel_ref [B,N+1], curv_ref [B,N], ex [B]).
    """
    L = nat.lib()
    N = solver.N
    use_torch = type(gstate).__module__.startswith("torch")
    if use_torch:
        import torch
        dev = gstate.device
        f = lambda v: v.to(device=dev, dtype=torch.float64).contiguous()  # noqa: E731
        i = lambda v: None if v is None else v.to(device=dev, dtype=torch.int32).contiguous()  # noqa: E731
        g, s_, r, lp, ix = f(gstate), f(s_prev), f(refs), i(lap), i(index)
        B, n_ref = int(g.shape[0]), int(r.shape[2])
        imax = 0 if ix is None or B == 0 else int(ix.max().item())
        if ix is not None and B and int(ix.min().item()) < 0:
            raise ValueError("negative window index")
        out = dict(x0=torch.empty((B, 6), dtype=torch.float64, device=dev), vel_ref=torch.empty((B, N + 1), dtype=torch.float64, device=dev),
                   curv_ref=torch.empty((B, N), dtype=torch.float64, device=dev), ex=torch.empty((B,), dtype=torch.float64, device=dev))
        stream = torch.cuda.current_stream(dev).cuda_stream
        nat.check(L.lpvmpc_track_inputs_dev(solver._h, B, _ptr(g), _ptr(lp), _ptr(s_), _ptr(r), n_ref, _ptr(ix), imax, _ptr(out["x0"]),
                                            _ptr(out["vel_ref"]), _ptr(out["curv_ref"]), _ptr(out["ex"]), C.c_void_p(stream)), solver._h)
        out["_keepalive"] = (g, s_, r, lp, ix)
        return out
    g = np.ascontiguousarray(gstate, dtype=np.float64)
    s_ = np.ascontiguousarray(s_prev, dtype=np.float64)
    r = np.ascontiguousarray(refs, dtype=np.float64)
    lp = None if lap is None else np.ascontiguousarray(lap, dtype=np.int32)
    ix = None if index is None else np.ascontiguousarray(index, dtype=np.int32)
    if g.ndim != 2 or g.shape[1] != 6 or r.ndim != 3 or r.shape[1] != 5 or r.shape[0] != g.shape[0] or s_.shape != (g.shape[0],):
        raise ValueError("gstate [B,6], s_prev [B], refs [B,5,n_ref] expected")
    B, n_ref = int(g.shape[0]), int(r.shape[2])
    out = dict(x0=np.empty((B, 6)), vel_ref=np.empty((B, N + 1)), curv_ref=np.empty((B, N)), ex=np.empty(B))
    nat.check(L.lpvmpc_track_inputs_host(solver._h, B, _ptr(g), _ptr(lp), _ptr(s_), _ptr(r), n_ref, _ptr(ix), _ptr(out["x0"]),
                                         _ptr(out["vel_ref"]), _ptr(out["curv_ref"]), _ptr(out["ex"])), solver._h)
    return out
