"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/lpv_ref.c + oracle/osqp_ref.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does (and fails loudly without its CUDA library).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblpv_oracle.so")
_SOURCES = ["lpv_ref.c", "osqp_ref.c", "loop_ref.c", "aux_ref.c", "lpv_ref.h", "osqp_ref.h", "loop_ref.h", "aux_ref.h", "Makefile"]

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)

STATUS_NAMES = {1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate",
                4: "dual infeasible inaccurate", -2: "maximum iterations reached", -3: "primal infeasible",
                -4: "dual infeasible", -7: "problem non convex", -10: "unsolved", -20: "schedule error"}


class Settings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double), ("eps_abs", C.c_double),
                ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("delta", C.c_double), ("adaptive_rho_tolerance", C.c_double),
                ("max_iter", C.c_int), ("check_termination", C.c_int), ("scaling", C.c_int),
                ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int),
                ("polish", C.c_int), ("polish_refine_iter", C.c_int), ("scaled_termination", C.c_int),
                ("linsys", C.c_int), ("cache_ordering", C.c_int)]


class Result(C.Structure):
    _fields_ = [("x", c_double_p), ("y", c_double_p), ("z", c_double_p),
                ("xs", c_double_p), ("zs", c_double_p), ("ys", c_double_p),
                ("active_lo", c_ubyte_p), ("active_up", c_ubyte_p),
                ("D", c_double_p), ("E", c_double_p), ("Ps", c_double_p), ("As", c_double_p), ("qs", c_double_p),
                ("status", C.c_int), ("iter", C.c_int), ("rho_updates", C.c_int), ("status_polish", C.c_int),
                ("n_factor", C.c_int),
                ("obj_val", C.c_double), ("pri_res", C.c_double), ("dua_res", C.c_double),
                ("rho_estimate", C.c_double), ("rho_final", C.c_double), ("c", C.c_double)]


class Vehicle(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("lf", "lr", "m", "Iz", "Cf", "Cr", "mu")]


class Cfg(C.Structure):
    _fields_ = [("N", C.c_int), ("dt", C.c_double), ("Q", C.c_double * 36), ("R", C.c_double * 4),
                ("dR", C.c_double * 2), ("L_cf", C.c_double * 5), ("max_vel", C.c_double), ("min_vel", C.c_double),
                ("steering_delay", C.c_int), ("veh", Vehicle), ("nseg", C.c_int), ("track", c_double_p)]


class QP(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("pnz", C.c_int), ("anz", C.c_int),
                ("Pp", c_int_p), ("Pi", c_int_p), ("Ap", c_int_p), ("Ai", c_int_p),
                ("Px", c_double_p), ("q", c_double_p), ("Ax", c_double_p), ("l", c_double_p), ("u", c_double_p)]


class Info(C.Structure):
    _fields_ = [("status", C.c_int), ("iter", C.c_int), ("rho_updates", C.c_int), ("status_polish", C.c_int),
                ("n_factor", C.c_int), ("sched_err", C.c_int),
                ("obj_val", C.c_double), ("pri_res", C.c_double), ("dua_res", C.c_double)]


_lib = None


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile) when missing or stale."""
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SOURCES)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "CC=gcc"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.lpv_ref_curvature.restype = C.c_double
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _bp(a):
    return None if a is None else a.ctypes.data_as(c_ubyte_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def default_settings(**kw):
    s = Settings()
    lib().osqp_ref_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def make_cfg(kind, N, dt, Q, R, dR, track, L_cf=None, veh=None, max_vel=5.0, min_vel=0.9, steering_delay=0):
    """kind: 'controller' (n=6) or 'planner' (n=5).  Keeps numpy buffers alive on the returned object."""
    n = 6 if kind == "controller" else 5
    cfg = Cfg()
    cfg.N = int(N)
    cfg.dt = float(dt)
    Qf = np.zeros(36)
    Qf[:n * n] = np.asarray(Q, dtype=np.float64).reshape(n * n)
    cfg.Q[:] = Qf.tolist()
    cfg.R[:] = np.asarray(R, dtype=np.float64).reshape(4).tolist()
    cfg.dR[:] = np.asarray(dR, dtype=np.float64).reshape(2).tolist()
    cfg.L_cf[:] = (np.zeros(5) if L_cf is None else np.asarray(L_cf, dtype=np.float64).reshape(5)).tolist()
    cfg.max_vel = float(max_vel)
    cfg.min_vel = float(min_vel)
    cfg.steering_delay = int(steering_delay)
    v = dict(lf=0.125, lr=0.125, m=1.98, Iz=0.03, Cf=60.0, Cr=60.0, mu=0.05)
    v.update(veh or {})
    for k, val in v.items():
        setattr(cfg.veh, k, float(val))
    cfg._track = _f64(track)
    cfg.nseg = cfg._track.shape[0]
    cfg.track = _dp(cfg._track)
    cfg._kind = kind
    cfg._n = n
    return cfg


def curvature(s, track):
    tr = _f64(track)
    err = C.c_int(0)
    k = lib().lpv_ref_curvature(C.c_double(float(s)), _dp(tr), C.c_int(tr.shape[0]), C.byref(err))
    if err.value:
        raise TypeError("only length-1 arrays can be converted to Python scalars")
    return k


def ctrl_predict(cfg, x, u, vel_ref, curv_ref, Cf_new, lap):
    N = cfg.N
    x, u, vel_ref = _f64(x), _f64(u), _f64(vel_ref)
    curv_ref = _f64(np.zeros(N) if curv_ref is None else curv_ref)
    st, A, B, Cc = np.zeros((N, 6)), np.zeros((N, 6, 6)), np.zeros((N, 6, 2)), np.zeros((N, 6))
    err = lib().lpv_ref_ctrl_predict(C.byref(cfg), _dp(x), _dp(u), _dp(vel_ref), _dp(curv_ref),
                                     C.c_double(Cf_new), C.c_int(lap), _dp(st), _dp(A), _dp(B), _dp(Cc))
    return st, A, B, Cc, err


def ctrl_estimate(cfg, traj, u):
    N = cfg.N
    traj, u = _f64(traj), _f64(u)
    A, B, Cc = np.zeros((N, 6, 6)), np.zeros((N, 6, 2)), np.zeros((N, 6))
    err = lib().lpv_ref_ctrl_estimate(C.byref(cfg), _dp(traj), C.c_int(traj.shape[1]), _dp(u),
                                      C.c_int(u.shape[1] if u.ndim == 2 else 1), _dp(A), _dp(B), _dp(Cc))
    return A, B, Cc, err


def plan_predict(cfg, x, SS, u):
    N = cfg.N
    x, SS, u = _f64(x), _f64(SS), _f64(u)
    st, A, B, Cc = np.zeros((N, 5)), np.zeros((N, 5, 5)), np.zeros((N, 5, 2)), np.zeros((N, 5))
    err = lib().lpv_ref_plan_predict(C.byref(cfg), _dp(x), _dp(SS), _dp(u), _dp(st), _dp(A), _dp(B), _dp(Cc))
    return st, A, B, Cc, err


def plan_estimate(cfg, traj, u):
    N = cfg.N
    traj, u = _f64(traj), _f64(u)
    A, B, Cc = np.zeros((N, 5, 5)), np.zeros((N, 5, 2)), np.zeros((N, 5))
    err = lib().lpv_ref_plan_estimate(C.byref(cfg), _dp(traj), C.c_int(traj.shape[1]), _dp(u),
                                      C.c_int(u.shape[1] if u.ndim == 2 else 1), _dp(A), _dp(B), _dp(Cc))
    return A, B, Cc, err


def _qp_to_py(qp):
    from scipy import sparse
    n, m = qp.n, qp.m
    Pp = np.ctypeslib.as_array(qp.Pp, (n + 1,)).copy()
    Pi = np.ctypeslib.as_array(qp.Pi, (max(qp.pnz, 1),))[:qp.pnz].copy()
    Px = np.ctypeslib.as_array(qp.Px, (max(qp.pnz, 1),))[:qp.pnz].copy()
    Ap = np.ctypeslib.as_array(qp.Ap, (n + 1,)).copy()
    Ai = np.ctypeslib.as_array(qp.Ai, (max(qp.anz, 1),))[:qp.anz].copy()
    Ax = np.ctypeslib.as_array(qp.Ax, (max(qp.anz, 1),))[:qp.anz].copy()
    out = dict(P=sparse.csc_matrix((Px, Pi, Pp), shape=(n, n)), A=sparse.csc_matrix((Ax, Ai, Ap), shape=(m, n)),
               q=np.ctypeslib.as_array(qp.q, (n,)).copy(), l=np.ctypeslib.as_array(qp.l, (m,)).copy(),
               u=np.ctypeslib.as_array(qp.u, (m,)).copy())
    lib().lpv_ref_qp_free(C.byref(qp))
    return out


def ctrl_qp(cfg, A, B, Cc, x0, vel_ref, old_steering, old_accel):
    A, B, x0, vel_ref = _f64(A), _f64(B), _f64(x0), _f64(vel_ref)
    Cc = _f64(np.zeros((cfg.N, 6)) if Cc is None else Cc)
    olds = _f64(np.atleast_1d(old_steering))
    qp = QP()
    lib().lpv_ref_ctrl_qp(C.byref(cfg), _dp(A), _dp(B), _dp(Cc), _dp(x0), _dp(vel_ref), C.c_int(vel_ref.shape[0]),
                          _dp(olds), C.c_double(old_accel), C.byref(qp))
    return _qp_to_py(qp)


def plan_qp(cfg, A, B, Cc, x0, u_old, max_ey, ey_lo=None, ey_hi=None):
    A, B, x0, u_old = _f64(A), _f64(B), _f64(x0), _f64(u_old)
    Cc = _f64(np.zeros((cfg.N, 5)) if Cc is None else Cc)
    ey_lo, ey_hi = _f64(ey_lo), _f64(ey_hi)
    qp = QP()
    lib().lpv_ref_plan_qp(C.byref(cfg), _dp(A), _dp(B), _dp(Cc), _dp(x0), _dp(u_old), C.c_double(max_ey),
                          _dp(ey_lo), _dp(ey_hi), C.byref(qp))
    return _qp_to_py(qp)


def osqp_solve(P, q, A, l, u, settings=None, want_scaled=False, **kw):
    """Generic OSQP restatement on (P, q, A, l, u); P may be full symmetric (upper triangle is taken, as
    the python wrapper of upstream does)."""
    from scipy import sparse
    P = sparse.triu(sparse.csc_matrix(P), format="csc")
    P.sort_indices()
    A = sparse.csc_matrix(A)
    A.sort_indices()
    n, m = P.shape[0], A.shape[0]
    st = settings if settings is not None else default_settings(**kw)
    q, l, u = _f64(q), _f64(l), _f64(u)
    Pp, Pi, Px = P.indptr.astype(np.int32), P.indices.astype(np.int32), _f64(P.data)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), _f64(A.data)
    r = Result()
    out = dict(x=np.zeros(n), y=np.zeros(m), z=np.zeros(m), xs=np.zeros(n), zs=np.zeros(m), ys=np.zeros(m),
               active_lo=np.zeros(m, dtype=np.uint8), active_up=np.zeros(m, dtype=np.uint8))
    r.x, r.y, r.z = _dp(out["x"]), _dp(out["y"]), _dp(out["z"])
    r.xs, r.zs, r.ys = _dp(out["xs"]), _dp(out["zs"]), _dp(out["ys"])
    r.active_lo, r.active_up = _bp(out["active_lo"]), _bp(out["active_up"])
    if want_scaled:
        out.update(D=np.zeros(n), E=np.zeros(m), Ps=np.zeros(max(P.nnz, 1)), As=np.zeros(max(A.nnz, 1)), qs=np.zeros(n))
        r.D, r.E, r.Ps, r.As, r.qs = _dp(out["D"]), _dp(out["E"]), _dp(out["Ps"]), _dp(out["As"]), _dp(out["qs"])
    rc = lib().osqp_ref_solve(C.c_int(n), C.c_int(m), _ip(Pp), _ip(Pi), _dp(Px), _dp(q), _ip(Ap), _ip(Ai), _dp(Ax),
                              _dp(l), _dp(u), C.byref(st), C.byref(r))
    if rc:
        raise ValueError("osqp_ref_solve failed with code %d" % rc)
    for k in ("status", "iter", "rho_updates", "status_polish", "n_factor", "obj_val", "pri_res", "dua_res",
              "rho_estimate", "rho_final", "c"):
        out[k] = getattr(r, k)
    return out


def _info_dict(info):
    return {k: getattr(info, k) for k, _ in Info._fields_}


def ctrl_solve(cfg, settings, x0, A=None, B=None, Cc=None, mode=0, x_sched=None, u_prev=None, vel_ref=None,
               curv_ref=None, Cf_new=60.0, lap=1, traj=None, old_steering=(0.0,), old_accel=0.0):
    N = cfg.N
    m = 6 * N + 6 * (N + 1) + cfg.steering_delay
    nz = 6 * (N + 1) + 2 * N
    xP, uP = np.zeros((N + 1, 6)), np.zeros((N, 2))
    alo, aup = np.zeros(m, dtype=np.uint8), np.zeros(m, dtype=np.uint8)
    xs, zs, ys = np.zeros(nz), np.zeros(m), np.zeros(m)
    info = Info()
    x0 = _f64(x0)
    A, B, Cc, x_sched, u_prev, vel_ref, curv_ref, traj = map(_f64, (A, B, Cc, x_sched, u_prev, vel_ref, curv_ref, traj))
    olds = _f64(np.atleast_1d(old_steering))
    rc = lib().lpv_ref_ctrl_solve(C.byref(cfg), C.byref(settings), C.c_int(mode), _dp(x0), _dp(A), _dp(B), _dp(Cc),
                                  _dp(x_sched), _dp(u_prev), _dp(vel_ref), C.c_int(vel_ref.shape[0]), _dp(curv_ref),
                                  C.c_double(Cf_new), C.c_int(lap), _dp(traj), _dp(olds), C.c_double(old_accel),
                                  _dp(xP), _dp(uP), C.byref(info), _bp(alo), _bp(aup), _dp(xs), _dp(zs), _dp(ys))
    if rc:
        raise ValueError("lpv_ref_ctrl_solve failed with code %d" % rc)
    d = _info_dict(info)
    d.update(xPred=xP, uPred=uP, active_lo=alo, active_up=aup, xs=xs, zs=zs, ys=ys)
    return d


def plan_solve(cfg, settings, x0, A=None, B=None, Cc=None, mode=0, x_sched=None, SS=None, u_prev=None, traj=None,
               u_old=(0.0, 0.0), max_ey=0.3, ey_lo=None, ey_hi=None):
    N = cfg.N
    nz = 5 * (N + 1) + 2 * N
    m = 5 * (N + 1) + nz
    xP, uP = np.zeros((N + 1, 5)), np.zeros((N, 2))
    alo, aup = np.zeros(m, dtype=np.uint8), np.zeros(m, dtype=np.uint8)
    xs, zs, ys = np.zeros(nz), np.zeros(m), np.zeros(m)
    info = Info()
    x0 = _f64(x0)
    A, B, Cc, x_sched, SS, u_prev, traj, ey_lo, ey_hi = map(_f64, (A, B, Cc, x_sched, SS, u_prev, traj, ey_lo, ey_hi))
    u_old = _f64(u_old)
    rc = lib().lpv_ref_plan_solve(C.byref(cfg), C.byref(settings), C.c_int(mode), _dp(x0), _dp(A), _dp(B), _dp(Cc),
                                  _dp(x_sched), _dp(SS), _dp(u_prev), _dp(traj), _dp(u_old), C.c_double(max_ey),
                                  _dp(ey_lo), _dp(ey_hi), _dp(xP), _dp(uP), C.byref(info), _bp(alo), _bp(aup),
                                  _dp(xs), _dp(zs), _dp(ys))
    if rc:
        raise ValueError("lpv_ref_plan_solve failed with code %d" % rc)
    d = _info_dict(info)
    d.update(xPred=xP, uPred=uP, active_lo=alo, active_up=aup, xs=xs, zs=zs, ys=ys)
    return d


def ctrl_batch(cfg, settings, x0, u_prev, vel_ref, curv_ref, lap, u_old, Cf_new=60.0, threads=1):
    B, N = x0.shape[0], cfg.N
    x0, u_prev, vel_ref, curv_ref, u_old = map(_f64, (x0, u_prev, vel_ref, curv_ref, u_old))
    lap = np.ascontiguousarray(lap, dtype=np.int32)
    xP, uP = np.zeros((B, N + 1, 6)), np.zeros((B, N, 2))
    status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    solved = lib().lpv_ref_ctrl_batch(C.byref(cfg), C.byref(settings), C.c_int(B), _dp(x0), _dp(u_prev), _dp(vel_ref),
                                      _dp(curv_ref), _ip(lap), _dp(u_old), C.c_double(Cf_new), C.c_int(threads),
                                      _dp(xP), _dp(uP), _ip(status), _ip(iters))
    return dict(xPred=xP, uPred=uP, status=status, iters=iters, solved=solved)


def plan_batch(cfg, settings, x0, SS, u_prev, u_old, max_ey, ey_lo=None, ey_hi=None, threads=1):
    B, N = x0.shape[0], cfg.N
    x0, SS, u_prev, u_old, max_ey, ey_lo, ey_hi = map(_f64, (x0, SS, u_prev, u_old, max_ey, ey_lo, ey_hi))
    xP, uP = np.zeros((B, N + 1, 5)), np.zeros((B, N, 2))
    status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    solved = lib().lpv_ref_plan_batch(C.byref(cfg), C.byref(settings), C.c_int(B), _dp(x0), _dp(SS), _dp(u_prev),
                                      _dp(u_old), _dp(max_ey), _dp(ey_lo), _dp(ey_hi), C.c_int(threads),
                                      _dp(xP), _dp(uP), _ip(status), _ip(iters))
    return dict(xPred=xP, uPred=uP, status=status, iters=iters, solved=solved)


# ------------------------------------------------------------------------------------------------
# closed-loop tick around the controller QP (oracle/loop_ref.c)
class LoopCfg(C.Structure):
    _fields_ = [("sim_dt", C.c_double), ("substeps", C.c_int), ("warmup_ticks", C.c_int), ("swap_ey_epsi", C.c_int),
                ("reserved", C.c_int), ("vel_ref", C.c_double), ("Cf_new", C.c_double), ("half_width", C.c_double),
                ("slack", C.c_double), ("sim_mu", C.c_double)]


def loop_cfg(half_width=0.3, slack=0.45, substeps=7, swap_ey_epsi=1, sim_dt=0.005, warmup_ticks=9, vel_ref=1.0,
             Cf_new=60.0, sim_mu=0.05):
    """Launch values: MAIN_LAUNCH.launch:60,72 (simulator dt, mu), controllerMain.py:77,310,326; Map.halfWidth/.slack of
    the L_shape track (trackInitialization.py:20,53-54)."""
    lc = LoopCfg()
    lc.sim_dt, lc.substeps, lc.warmup_ticks, lc.swap_ey_epsi = sim_dt, substeps, warmup_ticks, swap_ey_epsi
    lc.vel_ref, lc.Cf_new, lc.half_width, lc.slack, lc.sim_mu = vel_ref, Cf_new, half_width, slack, sim_mu
    return lc


def sim_f(state, u, veh=None, mu=0.05, dt=0.005):
    """One Simulator.f step; state = [x y yaw vx vy psiDot ax ay], u = [motor, servo].  Returns the new state."""
    v = Vehicle()
    d = dict(lf=0.125, lr=0.125, m=1.98, Iz=0.03, Cf=60.0, Cr=60.0, mu=0.05)
    d.update(veh or {})
    for k, val in d.items():
        setattr(v, k, float(val))
    st = _f64(state).copy()
    uu = _f64(u)
    lib().loop_ref_sim_f(_dp(st), _dp(uu), C.byref(v), C.c_double(mu), C.c_double(dt))
    return st


def local_position(track, half_width, slack, x, y, psi):
    tr = _f64(track)
    out = np.zeros(3)
    flag = lib().loop_ref_local_position(_dp(tr), C.c_int(tr.shape[0]), C.c_double(half_width), C.c_double(slack),
                                         C.c_double(x), C.c_double(y), C.c_double(psi), _dp(out))
    return out[0], out[1], out[2], flag


def global_position(track, s, ey):
    tr = _f64(track)
    out = np.zeros(3)
    err = lib().loop_ref_global_position(_dp(tr), C.c_int(tr.shape[0]), C.c_double(s), C.c_double(ey), _dp(out))
    if err:
        raise TypeError("no unique track segment holds s")
    return out[0], out[1], out[2]


def loop_state(sim0, N):
    """Fresh closed-loop state for vehicles with simulator states sim0 [B,8] = [x y yaw vx vy psiDot ax ay]."""
    sim = _f64(sim0).copy()
    B = sim.shape[0]
    ctr = np.zeros((B, 8), dtype=np.int32)
    ctr[:, 0] = 1                                  # first_it = 1 (controllerMain.py:87)
    stat = np.zeros((B, 4))
    stat[:, 3] = -1.0
    return dict(sim=sim, cmd=np.zeros((B, 2)), u_pred=np.zeros((B, N, 2)), local=np.zeros((B, 6)), ctr=ctr, stat=stat,
                x_pred=np.zeros((B, N + 1, 6)))


def loop_run(cfg, settings, lc, state, n_ticks, threads=1):
    """Advance `state` (from loop_state) by n_ticks closed-loop ticks in place; returns the number of SOLVED ticks."""
    lib().loop_ref_run.restype = C.c_long
    B = state["sim"].shape[0]
    return lib().loop_ref_run(C.byref(cfg), C.byref(settings), C.byref(lc), C.c_int(B), C.c_int(n_ticks),
                              _dp(state["sim"]), _dp(state["cmd"]), _dp(state["u_pred"]), _dp(state["local"]),
                              _ip(state["ctr"]), _dp(state["stat"]), _dp(state["x_pred"]), C.c_int(threads))


# ------------------------------------------------------------------------------------------------
# planner loop (oracle/loop_ref.c: plan_loop_ref_*)
def plan_guess(x0, N, accel_rate=0.2, dt=0.05, s0=0.0):
    xx = np.zeros((N + 1, 6))
    lib().plan_loop_ref_guess(_dp(_f64(x0)), C.c_int(N), C.c_double(accel_rate), C.c_double(dt), C.c_double(s0), _dp(xx))
    return xx


def plan_loop_state(xstart, N, s0=None):
    xs = _f64(xstart).copy()
    B = xs.shape[0]
    SS = np.zeros((B, N + 1))
    if s0 is not None:
        SS[:, 0] = s0
    return dict(xstart=xs, x_pred=np.zeros((B, N + 1, 5)), u_pred=np.zeros((B, N, 2)), SS=SS,
                ctr=np.zeros((B, 8), dtype=np.int32), stat=np.zeros((B, 4)))


def plan_loop_run(cfg, settings, state, n_ticks, max_ey=0.3, accel_rate=0.2, threads=1):
    lib().plan_loop_ref_run.restype = C.c_long
    B = state["xstart"].shape[0]
    return lib().plan_loop_ref_run(C.byref(cfg), C.byref(settings), C.c_int(B), C.c_int(n_ticks), C.c_double(max_ey),
                                   C.c_double(accel_rate), _dp(state["xstart"]), _dp(state["x_pred"]), _dp(state["u_pred"]),
                                   _dp(state["SS"]), _ip(state["ctr"]), _dp(state["stat"]), C.c_int(threads))


def track_inputs(gstate, s_prev, refs, N, dt, lap=None, index=None):
    """controllerMain.py:196-243 + Body_Frame_Errors for a batch (see loop_ref.h: track_inputs_ref)."""
    g, sp, r = _f64(gstate), _f64(s_prev), _f64(refs)
    B, n_ref = g.shape[0], r.shape[2]
    out = dict(x0=np.zeros((B, 6)), vel_ref=np.zeros((B, N + 1)), curv_ref=np.zeros((B, N)), ex=np.zeros(B))
    f = lib().track_inputs_ref
    f.restype = C.c_double
    for b in range(B):
        out["ex"][b] = f(_dp(g[b]), C.c_int(0 if lap is None else int(lap[b])), C.c_double(sp[b]), _dp(r[b]), C.c_int(n_ref),
                         C.c_int(0 if index is None else int(index[b])), C.c_int(N), C.c_double(dt), _dp(out["x0"][b]),
                         _dp(out["vel_ref"][b]), _dp(out["curv_ref"][b]))
    return out


# ------------------------------------------------------------------------------------------------
# SURVEY 8f row 4 (oracle/aux_ref.c): TS-fuzzy scheduling blend and the polytopic LPV observer
def anfis_abc(sched, A_tab, B_tab, C_tab, bell):
    """ABC_computation_5SV_new (PathFollowingLPVMPC.py:530-602) for a batch: sched [B,5] -> A [B,3], B [B,2], C [B]."""
    sched, A_tab, B_tab, C_tab, bell = map(_f64, (sched, A_tab, B_tab, C_tab, bell))
    n = sched.shape[0]
    A, Bm, Cc = np.zeros((n, 3)), np.zeros((n, 2)), np.zeros(n)
    f = lib().anfis_abc_ref
    f.restype = C.c_double
    for i in range(n):
        Cc[i] = f(_dp(sched[i]), _dp(A_tab), _dp(B_tab), _dp(C_tab), _dp(bell), _dp(A[i]), _dp(Bm[i]))
    return A, Bm, Cc


def observer_step(est, y, u, lim_ls, gains_ls, lim_hs, gains_hs, C_obs, dt, use_est):
    """GS_LPV_Est (stateEstimator.py:349-492) for a batch: est [B,6] -> new est [B,6]."""
    est = _f64(est).copy()
    y, u, lim_ls, gains_ls, lim_hs, gains_hs, C_obs = map(_f64, (y, u, lim_ls, gains_ls, lim_hs, gains_hs, C_obs))
    use = np.broadcast_to(np.asarray(use_est, dtype=np.int32), (est.shape[0],))
    f = lib().observer_step_ref
    f.restype = None
    for i in range(est.shape[0]):
        f(_dp(est[i]), _dp(y[i]), _dp(u[i]), _dp(lim_ls), _dp(gains_ls), _dp(lim_hs), _dp(gains_hs), _dp(C_obs), C.c_double(dt), C.c_int(int(use[i])))
    return est
