"""Synthetic workloads of BASELINE.json's configs: seeded numpy generators and the reference tunings.

Pure host-side data generation (no solver code); used by bench.py and the tests.
"""
import numpy as np

from .track import Map, curvature

# controllerMain.py:139-141 (path tracking) / :146-148 (trajectory tracking)
CTRL_PT = dict(Q=np.diag([100.0, 1.0, 1.0, 20.0, 0.0, 900.0]), R=0.5 * 0.5 * np.diag([1.0, 1.0]),
               dR=1.5 * 25 * np.array([1.3, 1.0]))
CTRL_TT = dict(Q=np.diag([400.0, 1.0, 1.0, 20.0, 0.0, 1100.0]), R=0.0 * np.diag([1.0, 1.0]), dR=np.array([100.0, 45.0]))
# plannerMain.py:96-99
PLAN = dict(Q=-np.diag([-0.000000000000088, -9.703658572659423, -0.5, 0.000000000213635, -0.153591566469547]),
            L_cf=-np.array([1.00702414775175, 0.187661946033823, -0.0, 0.0, -0.0329493219494661]),
            R=np.diag([0.8, 0.0]), dR=np.array([6.0, 6.0]))
CTRL_DT = 1.0 / 30.0   # controllerMain.py:43-44
PLAN_DT = 1.0 / 20.0   # MAIN_LAUNCH.launch:43 (Frecuency 20)


def track_kappa(s, pt):
    """Vectorised Curvature() for s >= 0 (no failure cases)."""
    L = pt[-1, 3] + pt[-1, 4]
    s = np.asarray(s, dtype=np.float64)
    sw = np.where(s > L, np.mod(s, L), s)
    idx = np.clip(np.searchsorted(pt[:, 3], sw, side="right") - 1, 0, pt.shape[0] - 1)
    return pt[idx, 5]


def controller_batch(B, N=8, seed=0, track=None, steer_scale=1.0):
    """Config 2 / 5 inputs: randomised x0 and scheduling vectors, LapNumber=1 (kappa from curv_ref)."""
    pt = (track if track is not None else Map("L_shape")).PointAndTangent
    rng = np.random.default_rng(seed)
    x0 = np.stack([rng.uniform(0.5, 3.0, B), rng.uniform(-0.2, 0.2, B), rng.uniform(-1, 1, B),
                   rng.uniform(-0.2, 0.2, B), rng.uniform(0, 19.2, B), rng.uniform(-0.2, 0.2, B)], axis=1)
    d0 = rng.uniform(-0.2, 0.2, (B, 1))
    steer = np.clip(d0 + 0.01 * np.cumsum(rng.standard_normal((B, N)), axis=1), -0.249, 0.249) * steer_scale
    acc = np.repeat(rng.uniform(-0.5, 1.5, (B, 1)), N, axis=1)
    u_prev = np.stack([steer, acc], axis=2)
    v0 = rng.uniform(0.8, 3.0, (B, 1))
    vel_ref = np.minimum(v0 + 0.05 * np.arange(N + 1)[None, :], 5.0)
    s_path = x0[:, 4:5] + np.cumsum(vel_ref[:, :N] * CTRL_DT, axis=1)
    curv_ref = track_kappa(s_path, pt)
    return dict(x0=x0, u_prev=u_prev, vel_ref=vel_ref, curv_ref=curv_ref, lap=np.ones(B, dtype=np.int32),
                u_old=u_prev[:, 0, :].copy())


_HARVEST = None


def planner_harvest():
    """The (x0 = xPred[1], SS, uPred) tuples recorded from the reference's own Testing-mode planner loop
    (plannerMain.py:145-146, 152-176, 201-211; HW = 0.2; 216 ticks until the loop turns infeasible at vx = 2.6 m/s):
    data/planner_harvest.npz, written by tests/golden/make_planner_harvest.py."""
    global _HARVEST
    if _HARVEST is None:
        import os
        with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "planner_harvest.npz")) as z:
            _HARVEST = {k: np.array(z[k]) for k in z.files}
    return _HARVEST


def planner_batch_harvest(B, N=40, seed=1, obstacle_share=0.25):
    """Config 3 inputs exactly as SURVEY.md 8d words them: (i) tuples harvested from the reference's planner loop,
    (ii) tuple index drawn uniformly, N(0, sigma) with sigma = [.05, .01, .05, .01, .01] added to [vx, vy, wz, ey, epsi],
    max_ey = HW = 0.2 for 75 % of the problems and for 25 % an 'obstacle': stages [k0, k0 + 8), k0 ~ U{10..30}, get
    ey in [0.05, HW]; uOld = 0 (what the reference effectively does).  A share of these QPs is genuinely
    infeasible (the unperturbed tuples already sit on the velocity / yaw-rate boxes in the corners)."""
    h = planner_harvest()
    if N != int(h["N"]):
        raise ValueError("the harvest was recorded at N = %d" % int(h["N"]))
    hw = float(h["half_width"])
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, h["x0"].shape[0], B)
    x0 = h["x0"][idx] + rng.standard_normal((B, 5)) * np.array([0.05, 0.01, 0.05, 0.01, 0.01])[None, :]
    max_ey = np.full(B, hw)
    ey_lo = np.repeat(-max_ey[:, None], N + 1, axis=1)
    ey_hi = np.repeat(max_ey[:, None], N + 1, axis=1)
    obs = rng.uniform(0, 1, B) < obstacle_share
    k0 = rng.integers(10, 31, B)
    for b in np.nonzero(obs)[0]:
        ey_lo[b, k0[b]:k0[b] + 8] = 0.05
    return dict(x0=x0, SS=h["SS"][idx].copy(), u_prev=h["u_pred"][idx].copy(), u_old=np.zeros((B, 2)), max_ey=max_ey, ey_lo=ey_lo, ey_hi=ey_hi,
                tuple_index=idx.astype(np.int32))


def planner_batch(B, N=40, seed=1, track=None, half_width=0.3, obstacle_share=0.25):
    """Planner inputs near a nominal straight/cornering roll-out (any horizon; used by the parity tests): per-problem
    max_ey and, for a share of problems, an 'obstacle' = tightened lateral interval on 8 consecutive stages.
    BASELINE configs[2] itself is `planner_batch_harvest`."""
    pt = (track if track is not None else Map("L_shape")).PointAndTangent
    rng = np.random.default_rng(seed)
    vx = rng.uniform(1.0, 2.4, B)
    x0 = np.stack([vx, rng.normal(0, 0.01, B), rng.normal(0, 0.05, B), rng.normal(0, 0.01, B), rng.normal(0, 0.01, B)], axis=1)
    s0 = rng.uniform(0, 19.0, (B, 1))
    acc = rng.uniform(0.0, 0.5, (B, 1))
    v_path = vx[:, None] + acc * PLAN_DT * np.arange(N + 1)[None, :]
    SS = s0 + np.concatenate([np.zeros((B, 1)), np.cumsum(v_path[:, :N] * PLAN_DT, axis=1)], axis=1)
    kap = track_kappa(SS[:, :N], pt)
    steer = np.clip(0.25 * kap + 0.005 * rng.standard_normal((B, N)), -0.249, 0.249)  # ~ L * kappa feed-forward
    u_prev = np.stack([steer, np.repeat(acc, N, axis=1)], axis=2)
    max_ey = np.full(B, half_width)
    ey_lo = np.repeat(-max_ey[:, None], N + 1, axis=1)
    ey_hi = np.repeat(max_ey[:, None], N + 1, axis=1)
    obs = rng.uniform(0, 1, B) < obstacle_share
    k0 = rng.integers(10, max(11, N - 9), B)
    for b in np.nonzero(obs)[0]:
        ey_lo[b, k0[b]:k0[b] + 8] = 0.05
    return dict(x0=x0, SS=SS, u_prev=u_prev, u_old=np.zeros((B, 2)), max_ey=max_ey, ey_lo=ey_lo, ey_hi=ey_hi)
