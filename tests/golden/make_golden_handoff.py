"""Generate tests/golden/handoff.npz from the REFERENCE'S OWN Python (read from /root/reference, never copied).

Run in the build container only:

    python tests/golden/make_golden_handoff.py

The trajectory-tracking branch of the controller node's main loop — controllerMain.py:200-243, from the yaw unwrapping
``GlobalState[5] = GlobalState[5]-2*np.pi*LapNumber`` to ``SS = LocalState[4]`` — is cut out of the reference file as
text, dedented and exec'd once per tick together with the reference's ``Body_Frame_Errors`` (controllerMain.py:495-506)
and ``wrap`` (Utilities/trackInitialization.py:413-421).  ``planning_data`` is a stand-in for the My_Planning message
filled from the recorded planner references of tests/golden/planner_refs.npz (themselves reference output).

Recorded per tick: the inputs (GlobalState, LapNumber, SS before, message), the window index BEFORE the tick, and the
reference's results (LocalState, Xerror, x_ref/y_ref/yaw_ref/vel_ref/curv_ref, SS after) for max_window = 0 (the value
in the file) and, with that one literal patched, max_window = 3.
"""
import os
import re
import sys
import textwrap
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_pieces():
    path = os.path.join(refload.REF_SRC, "controllerMain.py")
    src = open(path).read()
    bfe = re.search(r"^def Body_Frame_Errors \(.*?^    return s, ex, ey, epsi", src, re.S | re.M).group(0)
    blk = re.search(r"^ +GlobalState\[5\] = GlobalState\[5\]-2\*np\.pi\*LapNumber.*?^ +SS = LocalState\[4\] *$", src, re.S | re.M).group(0)
    tpath = os.path.join(refload.REF_SRC, "Utilities", "trackInitialization.py")
    wrap = re.search(r"^def wrap\(angle\):.*?^    return w_angle", open(tpath).read(), re.S | re.M).group(0)
    g = {"np": np}
    exec(compile(wrap, tpath, "exec"), g)
    exec(compile(bfe, path, "exec"), g)
    return g, textwrap.dedent(blk), path


def run(max_window, n_ticks, N, seed):
    g, blk, path = load_pieces()
    assert "max_window = 0" in blk
    if max_window != 0:
        blk = blk.replace("max_window = 0", "max_window = %d" % max_window)
    code = compile(blk, path, "exec")
    refs = np.load(os.path.join(OUT, "planner_refs.npz"))["refs"]           # [50,5,61]
    rng = np.random.default_rng(seed)
    ns = dict(g)
    ns.update(N=N, dt=1.0 / 30.0, index=0, SS=0.3, PlannerCounter=0, PSI_Planner=np.zeros((1, 4)), References=np.zeros(4),
              LocalState=np.zeros(8), planning_data=types.SimpleNamespace())
    rec = {k: [] for k in ("gstate", "lap", "s_prev", "msg", "index_before", "local", "ex", "x_ref", "y_ref", "yaw_ref", "vel_ref",
                           "curv_ref", "s_after", "fresh")}
    for t in range(n_ticks):
        m = refs[(3 * t + seed) % refs.shape[0]]
        pd = ns["planning_data"]
        pd.x_d, pd.y_d, pd.psi_d, pd.vx_d, pd.curv_d = (m[i].copy() for i in range(5))
        lap = 1 + (t // 17)
        # a vehicle near the head of the message, yaw carrying the laps the odometry has accumulated
        off = rng.normal(0, 0.05, 2)
        psi = m[2, 0] + rng.normal(0, 0.1) + 2 * np.pi * lap + (2 * np.pi if t % 11 == 5 else 0.0)
        G = np.array([rng.uniform(0.6, 2.5), rng.normal(0, 0.1), rng.normal(0, 0.5), m[0, 0] + off[0], m[1, 0] + off[1], psi])
        ns["GlobalState"] = G.copy()
        ns["LapNumber"] = lap
        ns["LocalState"][0:3] = G[0:3]          # controllerMain.py:190-192 LocalState[:] = estimatorData.CurrentState ... (vx, vy, wz)
        rec["gstate"].append(G.copy()); rec["lap"].append(lap); rec["s_prev"].append(ns["SS"]); rec["msg"].append(m.copy())
        rec["index_before"].append(ns["index"])
        before = ns.get("x_ref")
        exec(code, ns)
        rec["fresh"].append(0 if ns["x_ref"] is before else 1)
        rec["local"].append(ns["LocalState"][0:6].copy()); rec["ex"].append(ns["Xerror"])
        for k in ("x_ref", "y_ref", "yaw_ref", "vel_ref", "curv_ref"):
            rec[k].append(np.array(ns[k], dtype=np.float64))
        rec["s_after"].append(ns["SS"])
    return {k: np.array(v) for k, v in rec.items()}


def main():
    out = {}
    for tag, mw in (("w0", 0), ("w3", 3)):
        for k, v in run(mw, 40, 8, 5 + mw).items():
            out["%s_%s" % (tag, k)] = v
    np.savez_compressed(os.path.join(OUT, "handoff.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
