"""Fixed-iteration parity of one kernel variant against the oracle, per vector (xs, zs, ys) and iteration count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import lpvmpc_b200 as lp
import oracle
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
N, B = 8, 48
w = W.controller_batch(B, N, seed=11)
cfg = oracle.make_cfg("controller", N, W.CTRL_DT, W.CTRL_TT["Q"], W.CTRL_TT["R"], W.CTRL_TT["dR"], track)
for variant in [int(v) for v in sys.argv[1:]] or [3, 5]:
    for iters in (25, 50, 100, 200, 400):
        fixed = dict(max_iter=iters, check_termination=0, adaptive_rho=0, polish=0)
        s = lp.BatchSolver("controller", N, W.CTRL_DT, track=track, max_batch=B, variant=variant, **W.CTRL_TT, **fixed)
        r = s.solve(w["x0"], extra_outputs=("xs", "zs", "ys"), **{k: w[k] for k in ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")})
        st = oracle.default_settings(**fixed)
        worst = {k: (0.0, -1, 0.0) for k in ("xs", "zs", "ys")}
        for b in range(B):
            o = oracle.ctrl_solve(cfg, st, w["x0"][b], mode=1, x_sched=w["x0"][b], u_prev=w["u_prev"][b], vel_ref=w["vel_ref"][b],
                                  curv_ref=w["curv_ref"][b], lap=int(w["lap"][b]), old_steering=[w["u_old"][b, 0]], old_accel=float(w["u_old"][b, 1]))
            for k in worst:
                d = np.abs(r[k][b] - o[k]); i = int(d.argmax())
                e = d.max() / max(np.abs(o[k]).max(), 1e-300)
                if e > worst[k][0]: worst[k] = (e, b, i, float(np.abs(o[k]).max()))
        print("variant", variant, "iters", iters, {k: ("%.2e" % v[0], v[1:]) for k, v in worst.items()}, flush=True)
        s.close()
