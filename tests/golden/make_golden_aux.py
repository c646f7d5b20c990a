"""Generate tests/golden/aux.npz from the REFERENCE'S OWN Python (read from /root/reference, never copied).

Run in the build container only:

    python tests/golden/make_golden_aux.py

SURVEY.md 8f row 4:
  * ``ABC_computation_5SV_new`` (ControllerObject/PathFollowingLPVMPC.py:530-602): the 32-vertex TS-fuzzy (ANFIS) blend of
    vertex models with generalised-bell memberships.  The function is cut out of the reference file as text and exec'd.
    Its vertex tables (A 32x3, B 32x2, C 32) and bell parameters (10x3) are NOT in the reference repository (they come
    from .mat files that were never committed): the vectors use seeded random tables of the right shapes.
  * the polytopic LPV observer (stateEstimator.py:349-492): ``GS_LPV_Est`` + ``Continuous_AB_Comp`` + ``L_Gain_Comp`` cut
    out of the class body as text and exec'd as methods of a bare object whose attributes are what ``__init__`` would
    have set (C_obs as stateEstimator.py:242-246; the gain tables Llmi (6x5x16) and SchedVars_Limits (6x2) of the two
    polytopes are absent .mat files: seeded random tables of those shapes, limits ordered lo < hi).  ``rospy`` is a stub
    whose clock the script drives (the observer switches from measurements to its own estimate as scheduling variables
    once curr_time > 0.02 s).
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


def anfis_vectors(seed=7, B=64):
    path = os.path.join(refload.REF_SRC, "ControllerObject", "PathFollowingLPVMPC.py")
    src = open(path).read()
    fn = re.search(r"^def ABC_computation_5SV_new\(.*?^    return Anew, Bnew, Cnew", src, re.S | re.M).group(0)
    g = {"np": np}
    exec(compile(fn, path, "exec"), g)
    f = g["ABC_computation_5SV_new"]
    rng = np.random.default_rng(seed)
    A_tab, B_tab, C_tab = rng.normal(0, 1, (32, 3)), rng.normal(0, 1, (32, 2)), rng.normal(0, 1, 32)
    # bell parameters [a (width), b (slope), c (centre)] per membership function: two per scheduling variable
    centres = np.array([0.5, 3.0, -0.3, 0.3, -2.0, 2.0, -0.25, 0.25, -1.0, 2.0])
    bell = np.stack([rng.uniform(0.5, 2.0, 10), rng.uniform(1.0, 3.0, 10), centres], axis=1)
    sched = np.stack([rng.uniform(0.3, 3.5, B), rng.uniform(-0.4, 0.4, B), rng.uniform(-2.5, 2.5, B), rng.uniform(-0.25, 0.25, B),
                      rng.uniform(-1.0, 2.0, B)], axis=1)
    Ao, Bo, Co = np.zeros((B, 3)), np.zeros((B, 2)), np.zeros(B)
    for i in range(B):
        a, b, c = f(sched[i, 0], sched[i, 1], sched[i, 2], sched[i, 3], sched[i, 4], A_tab, B_tab, C_tab, bell)
        Ao[i], Bo[i], Co[i] = a, b, c
    return dict(an_A_tab=A_tab, an_B_tab=B_tab, an_C_tab=C_tab, an_bell=bell, an_sched=sched, an_A=Ao, an_B=Bo, an_C=Co)


def observer_vectors(seed=11, B=24, ticks=12):
    path = os.path.join(os.path.dirname(refload.REF_SRC.rstrip("/")), "src", "stateEstimator.py")
    if not os.path.exists(path):
        path = os.path.join(refload.REF_SRC, "stateEstimator.py")
    src = open(path).read()
    pieces = []
    for name, end in (("GS_LPV_Est", r"self\.index \+= 1"), ("Continuous_AB_Comp", r"\[ 0\.,\s+0\.,\s+1\.,\s+0\.,\s+0\.,\s+0\.\]\]\); # \[theta\]"),
                      ("L_Gain_Comp", r"self\.L_gain = result")):
        m = re.search(r"^    def %s\(self.*?%s[^\n]*$" % (name, end), src, re.S | re.M)
        assert m, name
        pieces.append(textwrap.dedent(m.group(0)))
    clock = types.SimpleNamespace(t=0.0)
    rospy = types.SimpleNamespace(get_rostime=lambda: types.SimpleNamespace(to_sec=lambda: clock.t))
    g = {"np": np, "dot": np.dot, "rospy": rospy}
    for p in pieces:
        exec(compile(p, path, "exec"), g)
    rng = np.random.default_rng(seed)

    def tables():
        lim = np.zeros((6, 2))
        lim[:, 0] = [0.1, -0.4, -3.0, -0.3, -10.0, -3.5]
        lim[:, 1] = [1.1, 0.4, 3.0, 0.3, 10.0, 3.5]
        return lim, rng.normal(0, 0.5, (6, 5, 16))
    lim_ls, gains_ls = tables()
    lim_hs, gains_hs = tables()
    lim_hs[0] = [1.0, 4.0]
    C_obs = np.zeros((5, 6))
    for i, j in enumerate((0, 2, 3, 4, 5)):
        C_obs[i, j] = 1.0
    dt = 0.005
    est0 = np.stack([rng.uniform(0.3, 3.0, B), rng.normal(0, 0.05, B), rng.normal(0, 0.3, B), rng.normal(0, 1, B), rng.normal(0, 1, B),
                     rng.uniform(-3, 3, B)], axis=1)
    ys, us, outs, warm = [], [], [], []
    est = est0.copy()
    for t in range(ticks):
        y = np.stack([est[:, 0] + rng.normal(0, 0.05, B), est[:, 2] + rng.normal(0, 0.05, B), est[:, 3] + rng.normal(0, 0.02, B),
                      est[:, 4] + rng.normal(0, 0.02, B), est[:, 5] + rng.normal(0, 0.02, B)], axis=1)
        u = np.stack([rng.uniform(-0.25, 0.25, B), rng.uniform(-1, 2, B)], axis=1)
        new = np.zeros_like(est)
        for b in range(B):
            o = types.SimpleNamespace(t0=0.0, prev_time=0.0, dt=dt, states_est=est[b].copy(), C_obs=C_obs, n_states=6, n_meas=5,
                                      L_gain=np.zeros((6, 5)), Est_Gains_LS=gains_ls, SchedVars_Limits_LS=lim_ls, Est_Gains_HS=gains_hs,
                                      SchedVars_Limits_HS=lim_hs, index=0)
            clock.t = 0.01 if t < 2 else 0.05 + dt * t    # the first two ticks schedule on the measurements (curr_time <= 0.02)
            g["GS_LPV_Est"](o, o.states_est, y[b], u[b], lambda *a: g["Continuous_AB_Comp"](o, *a), lambda *a: g["L_Gain_Comp"](o, *a))
            new[b] = o.states_est
        ys.append(y); us.append(u); outs.append(new.copy()); warm.append(0 if t < 2 else 1)
        est = new
    return dict(ob_lim_ls=lim_ls, ob_gains_ls=gains_ls, ob_lim_hs=lim_hs, ob_gains_hs=gains_hs, ob_C=C_obs, ob_dt=dt, ob_est0=est0,
                ob_y=np.array(ys), ob_u=np.array(us), ob_est=np.array(outs), ob_warm=np.array(warm, dtype=np.int32))


if __name__ == "__main__":
    d = {}
    d.update(anfis_vectors())
    d.update(observer_vectors())
    np.savez_compressed(os.path.join(OUT, "aux.npz"), **d)
    print("wrote", os.path.join(OUT, "aux.npz"), {k: v.shape for k, v in d.items() if hasattr(v, "shape")})
