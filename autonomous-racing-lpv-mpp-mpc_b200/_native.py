"""ctypes binding of the C-ABI in include/lpvmpc.h (liblpvmpc.so, built in-tree by nvcc).

There is no CPU fallback: if the library is missing it is (re)built with nvcc; if that is impossible, or no
CUDA device is present when a handle is created, an exception is raised.
"""
import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "liblpvmpc.so")
HASH_PATH = os.path.join(_PKG, "liblpvmpc.srchash")
_SRC_DIR = os.path.join(_PKG, "csrc")
_SOURCES = [os.path.join(_SRC_DIR, f) for f in ("lpvmpc.cu", "lpv_qp.cuh", "lpv_h8.cuh", "lpv_h8t.cuh", "lpv_h16t.cuh", "lpv_model.cuh", "lpv_loop.cuh", "lpv_aux.cuh")] + \
           [os.path.join(_ROOT, "include", "lpvmpc.h")]

# No --split-compile: with it ptxas works on the kernels in parallel (45 s instead of 2 min 25 s) but the result is not
# reproducible -- two builds of the same sources differ in register allocation, inlining and size (458 k against 485 k SASS
# lines; the helper-warp H8 kernels came out at 168 or at 255 registers) and the one-QP-per-warp workloads moved by up to
# 13 % from build to build.  Development builds can pass it through LPVMPC_NVCC_EXTRA.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared"]

ABI_VERSION = 1
CONTROLLER, PLANNER = 0, 1
SCHED_GIVEN, SCHED_PREDICT, SCHED_ESTIMATE = 0, 1, 2

STATUS_NAMES = {1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate",
                4: "dual infeasible inaccurate", -2: "maximum iterations reached", -3: "primal infeasible",
                -4: "dual infeasible", -7: "problem non convex", -10: "unsolved",
                -20: "schedule error (Curvature lookup failed)", -21: "data error (l > u)", -22: "off track"}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class Settings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double),
                ("eps_dual_inf", C.c_double), ("delta", C.c_double), ("adaptive_rho_tolerance", C.c_double),
                ("max_iter", C.c_int32), ("check_termination", C.c_int32), ("scaling", C.c_int32),
                ("adaptive_rho", C.c_int32), ("adaptive_rho_interval", C.c_int32), ("polish", C.c_int32),
                ("polish_refine_iter", C.c_int32), ("scaled_termination", C.c_int32)]


class Cfg(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("kind", C.c_int32), ("N", C.c_int32), ("steering_delay", C.c_int32),
                ("dt", C.c_double), ("Q", C.c_double * 36), ("R", C.c_double * 4), ("dR", C.c_double * 2),
                ("L_cf", C.c_double * 5),
                ("lf", C.c_double), ("lr", C.c_double), ("m", C.c_double), ("Iz", C.c_double), ("Cf", C.c_double),
                ("Cr", C.c_double), ("mu", C.c_double), ("max_vel", C.c_double), ("min_vel", C.c_double),
                ("n_track_seg", C.c_int32), ("max_batch", C.c_int32), ("track", c_double_p),
                ("device", C.c_int32), ("variant", C.c_int32), ("settings", Settings)]


class Info(C.Structure):
    _fields_ = [("n", C.c_int32), ("d", C.c_int32), ("N", C.c_int32), ("nz", C.c_int32), ("m", C.c_int32),
                ("variant", C.c_int32), ("workspace_in_smem", C.c_int32), ("smem_bytes_per_qp", C.c_int32),
                ("workspace_bytes", C.c_int64), ("kernel_launches", C.c_int64)]


_IN_D = ["x0", "x_sched", "A", "Bm", "C", "u_prev", "vel_ref", "curv_ref", "SS"]
_OUT_D = ["x_pred", "u_pred"]


class Args(C.Structure):
    _fields_ = ([("sched_mode", C.c_int32), ("x0_from_prediction", C.c_int32), ("lap_all", C.c_int32),
                 ("natural_order", C.c_int32), ("Cf_new", C.c_double)] +
                [(k, C.c_void_p) for k in _IN_D] +
                [("lap", C.c_void_p), ("traj", C.c_void_p), ("u_old", C.c_void_p), ("old_steering", C.c_void_p),
                 ("max_ey", C.c_void_p), ("ey_lo", C.c_void_p), ("ey_hi", C.c_void_p)] +
                [(k, C.c_void_p) for k in _OUT_D] +
                [("status", C.c_void_p), ("iters", C.c_void_p), ("rho_updates", C.c_void_p),
                 ("polish_status", C.c_void_p), ("obj", C.c_void_p), ("pri_res", C.c_void_p), ("dua_res", C.c_void_p),
                 ("active_lo", C.c_void_p), ("active_up", C.c_void_p), ("y", C.c_void_p), ("A_out", C.c_void_p),
                 ("B_out", C.c_void_p), ("states_out", C.c_void_p), ("xs", C.c_void_p), ("zs", C.c_void_p),
                 ("ys", C.c_void_p), ("order_hint", C.c_void_p)])


class LoopCfg(C.Structure):
    _fields_ = [("sim_dt", C.c_double), ("substeps", C.c_int32), ("warmup_ticks", C.c_int32),
                ("swap_ey_epsi", C.c_int32), ("reserved", C.c_int32), ("vel_ref", C.c_double), ("Cf_new", C.c_double),
                ("half_width", C.c_double), ("slack", C.c_double), ("sim_mu", C.c_double)]


class LoopState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("sim", "cmd", "u_pred", "x_pred", "local", "stat", "ctr")]


class PlanLoopState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("x_pred", "u_pred", "SS", "stat", "ctr")]


EXPORTS = ["lpvmpc_abi_version", "lpvmpc_default_settings", "lpvmpc_device_count", "lpvmpc_create", "lpvmpc_destroy",
           "lpvmpc_last_error", "lpvmpc_get_info", "lpvmpc_update_settings", "lpvmpc_schedule_dev",
           "lpvmpc_schedule_host", "lpvmpc_solve_dev", "lpvmpc_solve_host", "lpvmpc_solve_host_view", "lpvmpc_loop_default_cfg", "lpvmpc_loop_init_dev",
           "lpvmpc_loop_init_host", "lpvmpc_loop_run_dev", "lpvmpc_loop_run_host", "lpvmpc_loop_view_dev",
           "lpvmpc_loop_read_host", "lpvmpc_plan_loop_init_host", "lpvmpc_plan_loop_init_dev", "lpvmpc_plan_loop_run_host",
           "lpvmpc_plan_loop_run_dev", "lpvmpc_plan_loop_view_dev", "lpvmpc_plan_loop_read_host", "lpvmpc_plan_refs_setup",
           "lpvmpc_plan_refs_dev", "lpvmpc_plan_refs_host", "lpvmpc_track_inputs_dev", "lpvmpc_track_inputs_host",
           "lpvmpc_anfis_abc_dev", "lpvmpc_anfis_abc_host", "lpvmpc_observer_step_dev", "lpvmpc_observer_step_host"]

_lib = None


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _source_hash():
    """sha256 over the sources and the build flags: what the built library is checked against (mtimes do not survive
    the copy to a GPU box, content does)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for s in _SOURCES:
        with open(s, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile csrc/lpvmpc.cu for sm_100a into liblpvmpc.so (in-tree) when it is missing or was built from other
    sources (content hash kept beside it in liblpvmpc.srchash)."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("liblpvmpc.so is missing/stale and nvcc was not found; there is no CPU fallback")
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:     # ranks of one node may arrive here together
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or is_stale():
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            # LPVMPC_NVCC_EXTRA: extra flags for development builds (e.g. --split-compile=0); not part of the source hash
            import shlex
            cmd = [nvcc] + NVCC_FLAGS + shlex.split(os.environ.get("LPVMPC_NVCC_EXTRA", "")) + \
                  ["-I", os.path.join(_ROOT, "include"), "-o", tmp, os.path.join(_SRC_DIR, "lpvmpc.cu")]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            env = dict(os.environ)
            env.pop("CC", None)
            env.pop("CXX", None)
            subprocess.check_call(cmd, env=env)
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH, "w") as fh:
                fh.write(_source_hash() + "\n")
    return LIB_PATH


def lib():
    """Load the native library (building it if needed).  Raises when it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    build()   # no-op unless the library is missing or its sources changed since it was built
    L = C.CDLL(LIB_PATH)
    L.lpvmpc_abi_version.restype = C.c_int
    L.lpvmpc_device_count.restype = C.c_int
    L.lpvmpc_last_error.restype = C.c_char_p
    L.lpvmpc_last_error.argtypes = [C.c_void_p]
    L.lpvmpc_default_settings.argtypes = [C.POINTER(Settings)]
    L.lpvmpc_create.argtypes = [C.POINTER(Cfg), C.POINTER(C.c_void_p)]
    L.lpvmpc_destroy.argtypes = [C.c_void_p]
    L.lpvmpc_destroy.restype = None
    L.lpvmpc_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    L.lpvmpc_update_settings.argtypes = [C.c_void_p, C.POINTER(Settings)]
    L.lpvmpc_schedule_dev.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Args), C.c_void_p, C.c_void_p]
    L.lpvmpc_schedule_host.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Args), C.c_void_p]
    L.lpvmpc_solve_dev.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Args), C.c_void_p]
    L.lpvmpc_solve_host.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Args)]
    L.lpvmpc_solve_host_view.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Args), C.POINTER(Args)]
    L.lpvmpc_loop_default_cfg.argtypes = [C.POINTER(LoopCfg)]
    L.lpvmpc_loop_default_cfg.restype = None
    L.lpvmpc_loop_init_dev.argtypes = [C.c_void_p, C.c_int32, C.POINTER(LoopCfg), C.c_void_p, C.c_void_p]
    L.lpvmpc_loop_init_host.argtypes = [C.c_void_p, C.c_int32, C.POINTER(LoopCfg), C.c_void_p]
    L.lpvmpc_loop_run_dev.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.lpvmpc_loop_run_host.argtypes = [C.c_void_p, C.c_int32]
    L.lpvmpc_loop_view_dev.argtypes = [C.c_void_p, C.POINTER(LoopState), C.POINTER(C.c_int32)]
    L.lpvmpc_loop_read_host.argtypes = [C.c_void_p, C.POINTER(LoopState)]
    L.lpvmpc_plan_loop_init_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
    L.lpvmpc_plan_loop_init_dev.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    L.lpvmpc_plan_loop_run_host.argtypes = [C.c_void_p, C.c_int32]
    L.lpvmpc_plan_loop_run_dev.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.lpvmpc_plan_loop_view_dev.argtypes = [C.c_void_p, C.POINTER(PlanLoopState), C.POINTER(C.c_int32)]
    L.lpvmpc_plan_loop_read_host.argtypes = [C.c_void_p, C.POINTER(PlanLoopState)]
    L.lpvmpc_plan_refs_setup.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.lpvmpc_plan_refs_dev.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lpvmpc_plan_refs_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lpvmpc_track_inputs_dev.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lpvmpc_track_inputs_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lpvmpc_anfis_abc_dev.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 9
    L.lpvmpc_anfis_abc_host.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
    L.lpvmpc_observer_step_dev.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p, C.c_int32, C.c_void_p]
    L.lpvmpc_observer_step_host.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p, C.c_int32]
    if L.lpvmpc_abi_version() != ABI_VERSION:
        raise RuntimeError("liblpvmpc.so ABI version mismatch; rebuild with _native.build(force=True)")
    _lib = L
    return L


def default_settings(**kw):
    s = Settings()
    lib().lpvmpc_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError("unknown OSQP setting %r" % k)
        setattr(s, k, v)
    return s


class NativeError(RuntimeError):
    pass


def check(rc, handle=None):
    if rc != 0:
        msg = lib().lpvmpc_last_error(handle)
        raise NativeError("lpvmpc error %d: %s" % (rc, msg.decode() if msg else "?"))
