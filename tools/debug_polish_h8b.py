import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import lpvmpc_b200 as lp
import oracle
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
N, B = 8, 254
w = W.controller_batch(B, N, seed=0)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
res = {}
for variant in (3, 5):
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, variant=variant, **W.CTRL_TT, polish=1)
    res[variant, 1] = s.solve(w["x0"], extra_outputs=("active_lo", "active_up", "y"), **{k: w[k] for k in keys})
    s.close()
    s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, variant=variant, **W.CTRL_TT, polish=0)
    res[variant, 0] = s.solve(w["x0"], extra_outputs=("xs", "zs", "ys"), **{k: w[k] for k in keys})
    s.close()
for b in (160, 166, 171, 235):
    a3 = res[3, 1].active_lo[b].astype(int) + 2 * res[3, 1].active_up[b].astype(int)
    a5 = res[5, 1].active_lo[b].astype(int) + 2 * res[5, 1].active_up[b].astype(int)
    d = np.nonzero(a3 != a5)[0]
    print("b", b, "rows that differ", d, "a3", a3[d], "a5", a5[d])
    for k in ("xs", "zs", "ys"):
        e = np.abs(res[3, 0][k][b] - res[5, 0][k][b])
        print("   ", k, "max diff %.3e at %d" % (e.max(), e.argmax()), "| at differing rows:", res[3, 0][k][b][d] if k != "xs" else "", res[5, 0][k][b][d] if k != "xs" else "")
