"""Controller variants on sub-batches of the ctrl65536 distribution: the QPs with the most / the fewest ADMM iterations."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lpvmpc_b200 as lp
W = lp.workloads
track = lp.Map("L_shape").PointAndTangent
dev = torch.device("cuda", 0)
keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
B = 65536
w = W.controller_batch(B, 8, seed=0)
s = lp.BatchSolver("controller", 8, W.CTRL_DT, track=track, max_batch=B, variant=6, **W.CTRL_TT)
r = s.solve(w["x0"], **{k: w[k] for k in keys}); s.close()
it = np.asarray(r.iters)
order = np.argsort(it, kind="stable")
sets = {"fewest": order[:16384], "most": order[-16384:], "random": np.random.default_rng(0).permutation(B)[:16384]}
for name, idx in sets.items():
    tin = {k: torch.as_tensor(w[k][idx]).to(dev) for k in keys}; tx0 = torch.as_tensor(w["x0"][idx]).to(dev)
    for variant in (6, 8):
        for nat_order in ("key", "true_iters", "const"):
            s = lp.BatchSolver("controller", 8, W.CTRL_DT, track=track, max_batch=len(idx), variant=variant, **W.CTRL_TT)
            kw = {} if nat_order == "key" else dict(order_hint=torch.as_tensor(it[idx].astype(np.int32) if nat_order == "true_iters" else np.full(len(idx), 100, np.int32)).to(dev))
            for _ in range(2): s.solve(tx0, **tin, **kw)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for a, b in ev:
                a.record(); s.solve(tx0, **tin, **kw); b.record()
            torch.cuda.synchronize()
            t = np.array([a.elapsed_time(b) for a, b in ev])
            print(json.dumps({"set": name, "iters_mean": float(it[idx].mean()), "variant": variant, "order": nat_order, "kernel_ms_p50": round(float(np.percentile(t, 50)), 3)}))
            s.close()
