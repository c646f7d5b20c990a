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


def planner_batch(B, N=40, seed=1, track=None, half_width=0.3, obstacle_share=0.25):
    """Config 3 inputs: planner states near a nominal straight/cornering roll-out, per-problem max_ey and,
    for a share of problems, an 'obstacle' = tightened lateral interval on 8 consecutive stages."""
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
