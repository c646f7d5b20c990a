"""Cycles per phase of the helper-warp H8 kernels (planner N = 40, controller N = 100), from a -DLPV_H8_PHASE_TIMING build:
   nvcc <NVCC_FLAGS> -DLPV_H8_PHASE_TIMING -I include -o tools/ab/liblpvmpc_timing.so <pkg>/csrc/lpvmpc.cu   (tools/build_timing_lib.sh)
usage (GPU box): python tools/h8_phase_timing.py tools/ab/liblpvmpc_timing.so"""
import os, sys, shutil, json, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
alt = sys.argv[1]
pkg = os.path.join(ROOT, "autonomous-racing-lpv-mpp-mpc_b200")
main = os.path.join(pkg, "liblpvmpc.so")
shutil.copy(main, "/tmp/main_keep.so")
shutil.copy(alt, main)
try:
    import torch
    import lpvmpc_b200 as lp
    W = lp.workloads
    L = lp._native.lib()
    L.lpvmpc_debug_phase_cycles.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    names = ["setup", "factor", "reproject", "snapshot", "fwd sweep", "bwd: wait for the helpers", "sync_yd", "update_info", "check_termination",
             "loop glue", "final checks / objective / iterate out", "polish", "outputs", "bwd: middle stage + go", "bwd: chain", ""]
    track = lp.Map("L_shape").PointAndTangent
    dev = torch.device("cuda", 0)
    for wl in ("ctrl1024N100", "plan16384"):
        if wl == "plan16384":
            B = 16384; w = W.planner_batch_harvest(B, 40, seed=1)
            s = lp.BatchSolver("planner", 40, W.PLAN_DT, track=track, max_batch=B, **W.PLAN)
            keys = ("SS", "u_prev", "u_old", "max_ey", "ey_lo", "ey_hi")
        else:
            B = 1024; w = W.controller_batch(B, 100, seed=3, track=lp.Map("L_shape"), steer_scale=0.2)
            s = lp.BatchSolver("controller", 100, W.CTRL_DT, track=track, max_batch=B, **W.CTRL_TT)
            keys = ("u_prev", "vel_ref", "curv_ref", "lap", "u_old")
        tin = {k: torch.as_tensor(w[k]).to(dev) for k in keys if k in w}
        tx0 = torch.as_tensor(w["x0"]).to(dev)
        r = s.solve(tx0, **tin); torch.cuda.synchronize()
        out = (C.c_ulonglong * 16)()
        L.lpvmpc_debug_phase_cycles(out, 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = s.solve(tx0, **tin); b.record(); torch.cuda.synchronize()
        L.lpvmpc_debug_phase_cycles(out, 1)
        cyc = np.array(list(out), dtype=np.float64)
        tot = cyc.sum()
        print(json.dumps({"workload": wl, "ms": round(a.elapsed_time(b), 3), "iters_mean": float(r.iters.double().mean()),
                          "cycles_per_qp": round(tot / B), "share": {n: round(float(c / tot), 4) for n, c in zip(names, cyc) if n}}))
        s.close()
finally:
    shutil.copy("/tmp/main_keep.so", main)
