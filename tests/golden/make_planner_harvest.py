"""Generate autonomous-racing-lpv-mpp-mpc_b200/data/planner_harvest.npz from the REFERENCE'S OWN planner loop.

    python tests/golden/make_planner_harvest.py        (build container only: needs /root/reference)

The reference's Testing-mode planner loop (plannerMain.py:145-146, 152-176, 201-211; start [1, 0, 0, 0, 0] at s = 0,
HW = 0.2) is run headless for 300 ticks exactly as tests/golden/make_golden_planloop.py runs it (the reference's
LPV_MPC_Planner class, Curvature and predicted_vectors_generation; the QP solved at the osqp stub seam by the oracle's
OSQP restatement).  After every tick the inputs of the NEXT tick's QP are harvested:

    x0 = xPred[1] (5)    SS (N+1, after the arc-length integration and SS[0] = SS[1])    uPred (N, 2)

workloads.planner_batch(source="harvest") draws from these tuples and perturbs them.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refload  # noqa: E402
import make_golden_planloop as mg  # noqa: E402

OUT = os.path.join(ROOT, "autonomous-racing-lpv-mpp-mpc_b200", "data", "planner_harvest.npz")


def main(ticks=300):
    ns = refload.load()
    guess = mg.load_guess()
    m = ns.Map()
    # A tick that ends infeasible makes the reference print QUIT... and carry on with garbage, after which Curvature()
    # raises (vx has grown to 2.6 m/s in a corner by then): the harvest is the longest prefix of the run that completes.
    def run(n):
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                return mg.planner_loop(ns, guess, m, [1.0, 0.0, 0.0, 0.0, 0.0], ticks=n, N=40, dt=1.0 / 20.0, HW=0.2)
        except TypeError:
            return None
    r = run(ticks)
    if r is None:
        lo, hi = 1, ticks          # run(lo) completes, run(hi) does not
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if run(mid) is None:
                hi = mid
            else:
                lo = mid
        r = run(lo)
        ticks = lo
    ok = np.isin(r["status"][:, 0], (1, 2, -2))
    x0 = r["xpred"][:, 1, :]
    keep = ok & np.isfinite(x0).all(axis=1)
    np.savez_compressed(OUT, x0=x0[keep], SS=r["SS"][keep], u_pred=r["upred"][keep], tick=np.nonzero(keep)[0].astype(np.int32),
                        half_width=np.float64(0.2), N=np.int32(40), dt=np.float64(1.0 / 20.0))
    print("ticks", ticks, "kept", int(keep.sum()), "vx range", x0[keep, 0].min(), x0[keep, 0].max(), "s range", r["SS"][keep, 0].min(), r["SS"][keep, 0].max(),
          "statuses", sorted(set(r["status"][:, 0].tolist())), "iters p50/max", int(np.median(r["status"][:, 1])), int(r["status"][:, 1].max()))
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
