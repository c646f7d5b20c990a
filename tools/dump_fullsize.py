"""Dumps the device's results on the full BASELINE batches (configs[1], [2] (nominal and harvested), [4]) into
gpurun_out/fullsize_<tag>.npz for offline comparison with the oracle's two linear-system back-ends (polish decisions)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lpvmpc_b200 as lp  # noqa: E402

W = lp.workloads
TRACK = lp.Map("L_shape").PointAndTangent
out = {}
for name, N, B, seed, steer in (("ctrl4096", 8, 4096, 0, 1.0), ("ctrl1024N100", 100, 1024, 3, 0.2)):
    w = W.controller_batch(B, N, seed=seed, steer_scale=steer)
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=TRACK, max_batch=B, **W.CTRL_TT)
    r = s.solve(w["x0"], **{k: w[k] for k in ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")})
    for k in ("status", "iters", "polish_status", "rho_updates", "pri_res", "dua_res", "obj", "u_pred"):
        out["%s_%s" % (name, k)] = np.asarray(r[k])
    out["%s_x_pred" % name] = np.asarray(r["x_pred"]).astype(np.float64)
    s.close()
for name, gen in (("plan16384nominal", lambda: W.planner_batch(16384, 40, seed=1)), ("plan16384", lambda: W.planner_batch_harvest(16384, 40, seed=1))):
    w = gen()
    s = lp.BatchSolver("planner", 40, W.PLAN_DT, track=TRACK, max_batch=16384, **W.PLAN)
    r = s.solve(w["x0"], **{k: w[k] for k in ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")})
    for k in ("status", "iters", "polish_status", "rho_updates", "pri_res", "dua_res", "obj", "u_pred"):
        out["%s_%s" % (name, k)] = np.asarray(r[k])
    out["%s_x1" % name] = np.asarray(r["x_pred"])[:, :4, :]
    s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r3"
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fullsize_%s.npz" % tag), **out)
print("dumped", {k: v.shape for k, v in out.items() if k.endswith("status")})
