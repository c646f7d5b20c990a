"""CPU baseline BASELINE.md §3 promised: the reference's REAL per-instance loop — for every QP, sequentially, the
reference's own Python `LPVPrediction` + `solve` (`_buildMatEqConst`, `_buildMatCost`, `_buildMatIneqConst`, the sparse
conversions of `osqp_solve_qp`; loaded headless from /root/reference by oracle/refload.py) with the QP solved at the
`osqp` stub seam by the oracle's OSQP restatement (upstream `osqp` is an absent PyPI dependency).

Runs in the build container only (the GPU box has no /root/reference); writes profiles/r3_cpu_reference_python.json,
which bench.py quotes next to the live C-port baselines.  One core (the reference loop is single-threaded Python).

    python tools/ref_python_baseline.py [n_qps]
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import refload  # noqa: E402
import lpvmpc_b200 as lp  # noqa: E402  (workload generator only: no solver code is imported)


def main(n=256):
    W = lp.workloads
    ns = refload.load()
    m = ns.Map()
    N = 8
    w = W.controller_batch(4096, N, seed=0)
    ctl = ns.PathFollowingLPV_MPC(W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], N, np.ones(N + 1), W.CTRL_DT, m, "OSQP", 0, 0)
    st = oracle.default_settings(polish=1)
    t_solve = [0.0]
    stat = []

    def backend(qp):
        t0 = time.perf_counter()
        r = oracle.osqp_solve(qp.P, qp.q, qp.A, qp.l, qp.u, settings=st)
        t_solve[0] += time.perf_counter() - t0
        stat.append(r["status"])
        return r["x"], r["status"]

    ns.OSQPSeam.backend = staticmethod(backend)
    ns.OSQPSeam.log = []
    t_pred = t_total = 0.0
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            for b in range(n):
                ctl.OldSteering = [float(w["u_old"][b, 0])]
                ctl.OldAccelera = [float(w["u_old"][b, 1])]
                t0 = time.perf_counter()
                states, A_L, B_L, C_L = ctl.LPVPrediction(w["x0"][b], w["u_prev"][b], w["vel_ref"][b], w["curv_ref"][b], 60.0, 1)
                t1 = time.perf_counter()
                ctl.solve(w["x0"][b], states, w["u_prev"][b], False, w["vel_ref"][b], A_L, B_L, C_L, 10)
                t2 = time.perf_counter()
                t_pred += t1 - t0
                t_total += t2 - t0
                ns.OSQPSeam.log.clear()
    finally:
        ns.OSQPSeam.backend = None
    out = {
        "what": "reference per-instance loop on the first %d QPs of ctrl4096 (seed 0): reference Python LPVPrediction + solve() "
                "(QP build + sparse conversions) + oracle OSQP restatement at the osqp seam; one core; build container" % n,
        "qps": n, "ms_per_qp": 1e3 * t_total / n, "qp_per_s_1core": n / t_total,
        "ms_lpv_prediction": 1e3 * t_pred / n, "ms_build_and_conversions": 1e3 * (t_total - t_pred - t_solve[0]) / n,
        "ms_osqp_restatement_incl_ctypes": 1e3 * t_solve[0] / n, "solved_fraction": float(np.mean(np.array(stat) == 1)),
        "host": {"cores": os.cpu_count(), "python": sys.version.split()[0], "numpy": np.__version__},
    }
    path = os.path.join(ROOT, "profiles", "r3_cpu_reference_python.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
