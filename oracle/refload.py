"""TEST INFRASTRUCTURE — headless loader for the *reference's own* Python classes.

Loads ``PathFollowingLPVMPC.py``, ``LPV_MPC_Planner.py``, ``trackInitialization.py`` and
``utilities.py`` straight from the read-only reference checkout (never copied into this repo),
with stub ``rospy`` / ``cvxopt`` / ``osqp`` modules and the single Py2→Py3 source shim the files
need (``print 'x'`` → ``print('x')``).  Used ONLY

  * by ``tests/golden/make_golden.py`` to generate the committed golden vectors, and
  * by ``-m "not gpu"`` tests that cross-check the C restatement when ``/root/reference`` exists.

It cannot travel to the GPU box (``/root/reference`` is absent there); nothing on the product
path imports it.

Reference call sites that the stubs stand in for:
  rospy.get_param            PathFollowingLPVMPC.py:38-48, LPV_MPC_Planner.py:70-82,
                             trackInitialization.py:20,23
  osqp.OSQP().setup/solve    PathFollowingLPVMPC.py:302-323, LPV_MPC_Planner.py:204-215
  cvxopt (dead branch)       PathFollowingLPVMPC.py:13-14,26
"""
import os
import re
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("LPVMPC_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "workspace", "src", "barc", "src")

# MAIN_LAUNCH.launch:5-11,35,40-44
LAUNCH_PARAMS = {
    "lf": 0.125, "lr": 0.125, "m": 1.98, "Iz": 0.03, "Cf": 60.0, "Cr": 60.0, "mu": 0.05,
    "/TrajectoryPlanner/max_vel": 5.0,
    "/TrajectoryPlanner/min_vel": 0.9,
    "/TrajectoryPlanner/halfWidth": 0.2,
    "trackShape": "L_shape",
}

OSQP_CONSTANTS = {
    "OSQP_SOLVED": 1, "OSQP_SOLVED_INACCURATE": 2, "OSQP_MAX_ITER_REACHED": -2,
    "OSQP_PRIMAL_INFEASIBLE": -3, "OSQP_PRIMAL_INFEASIBLE_INACCURATE": 3,
    "OSQP_DUAL_INFEASIBLE": -4, "OSQP_DUAL_INFEASIBLE_INACCURATE": 4,
    "OSQP_NON_CVX": -7, "OSQP_UNSOLVED": -10, "OSQP_INFTY": 1e30,
}


def available():
    return os.path.isdir(REF_SRC)


class CapturedQP(object):
    """What the reference handed to ``OSQP.setup`` (after the wrapper's own normalisation)."""

    def __init__(self, P, q, A, l, u, settings):
        from scipy import sparse
        self.P = sparse.csc_matrix(P)
        self.q = np.asarray(q, dtype=np.float64).copy()
        self.A = sparse.csc_matrix(A)
        self.l = np.asarray(l, dtype=np.float64).copy()
        self.u = np.asarray(u, dtype=np.float64).copy()
        self.settings = dict(settings)


class _Info(object):
    status_val = 1
    status = "solved"


class _Result(object):
    def __init__(self, x, info):
        self.x = x
        self.info = info


class OSQPSeam(object):
    """The stub ``osqp.OSQP``: the one seam where a solver back-end is plugged in.

    ``OSQPSeam.backend`` is a callable ``(CapturedQP) -> (x, status_val)``; by default it returns
    zeros (build-only capture).  ``OSQPSeam.log`` collects every captured QP.
    """
    backend = None
    log = []

    def __init__(self):
        self._qp = None

    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        self._qp = CapturedQP(P, q, A, l, u, settings)
        OSQPSeam.log.append(self._qp)

    def warm_start(self, **kw):
        pass

    def solve(self):
        info = _Info()
        if OSQPSeam.backend is None:
            x = np.zeros(self._qp.q.shape[0])
        else:
            x, info.status_val = OSQPSeam.backend(self._qp)
        return _Result(x, info)

    @staticmethod
    def constant(name):
        return OSQP_CONSTANTS[name]


def _install_stubs(params):
    rospy = types.ModuleType("rospy")
    rospy._params = dict(params)
    rospy.get_param = lambda key, *a: rospy._params[key]
    rospy.ROSInterruptException = Exception
    sys.modules["rospy"] = rospy

    cvxopt = types.ModuleType("cvxopt")
    solvers = types.ModuleType("cvxopt.solvers")
    solvers.options = {}
    solvers.qp = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("cvxopt branch is dead code"))
    cvxopt.solvers = solvers
    cvxopt.spmatrix = lambda *a, **k: None
    cvxopt.matrix = lambda *a, **k: None
    sys.modules["cvxopt"] = cvxopt
    sys.modules["cvxopt.solvers"] = solvers

    osqp = types.ModuleType("osqp")
    osqp.OSQP = OSQPSeam
    sys.modules["osqp"] = osqp
    return rospy


_PRINT_RE = re.compile(r"^(\s*)print (.+)$", re.M)


def _load(name, relpath):
    path = os.path.join(REF_SRC, relpath)
    with open(path, "r") as fh:
        src = fh.read()
    src = _PRINT_RE.sub(r"\1print(\2)", src)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


_cache = {}


def load(params=None):
    """Return a namespace with the reference classes: Map, Curvature, PathFollowingLPV_MPC,
    LPV_MPC_Planner and the two private ``_EstimateABC`` helpers."""
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF_ROOT)
    key = tuple(sorted((params or LAUNCH_PARAMS).items()))
    if key in _cache:
        return _cache[key]
    rospy = _install_stubs(params or LAUNCH_PARAMS)
    util = _load("utilities", "Utilities/utilities.py")
    track = _load("trackInitialization", "Utilities/trackInitialization.py")
    ctrl = _load("PathFollowingLPVMPC", "ControllerObject/PathFollowingLPVMPC.py")
    plan = _load("LPV_MPC_Planner", "PlannerObject/LPV_MPC_Planner.py")
    ns = types.SimpleNamespace(
        rospy=rospy, utilities=util, trackInitialization=track, ctrl_mod=ctrl, plan_mod=plan,
        Map=track.Map, Curvature=util.Curvature,
        PathFollowingLPV_MPC=ctrl.PathFollowingLPV_MPC, LPV_MPC_Planner=plan.LPV_MPC_Planner,
        OSQPSeam=OSQPSeam)
    _cache[key] = ns
    return ns
