"""BASELINE configs[0]: one LPV-MPC path-following controller stepped closed loop, ROS stubbed, through the drop-in
``PathFollowingLPV_MPC`` class (GPU) with the CPU oracle solving the same tick beside it.

The loop restates controllerMain.py:177-454 (lap 0, path-tracking tune :139-141): 9 warm-up ticks that linearise
around the hard-coded guess of ``predicted_vectors_generation`` (:310-320, :510-553) through ``_EstimateABC``, then
``LPVPrediction`` + ``solve`` every tick (:326-331), the command taken from ``uPred[0]`` (:381-383) and pushed into
``OldSteering`` / ``OldAccelera`` (:289-298).

The plant is the reference simulator's bicycle model with the linear tyres ``Fy = 60 alpha``
(vehicleSimulator.py:164-199, 5 ms Euler sub-steps), integrated in curvilinear coordinates so that the test does
not need ``Map.getLocalPosition`` (not on the hot path).  True-state feedback (the estimator needs
gain tables that are not in the reference repository).

Tolerance: every tick's control and predicted states within 1e-6 of the oracle's (both polish to ~1e-10), status and
iteration count identical.
"""
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

lp = pytest.importorskip("lpvmpc_b200")
W = lp.workloads

VEH = dict(lf=0.125, lr=0.125, m=1.98, Iz=0.03, Cf=60.0, Cr=60.0, mu=0.05)  # MAIN_LAUNCH.launch:5-11


def guess_vectors(local, N):
    """predicted_vectors_generation (controllerMain.py:510-553), first N rows."""
    dv = [0.05, 0.2, 0.4, 0.6, 0.7, 0.8, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9]
    ds = [0, 0.01, 0.02, 0.04, 0.07, 0.1, 0.14, 0.18, 0.23, 0.55, 0.66, 0.77, 0.89, 1.00, 1.19, 1.39, 1.59, 1.79, 1.89, 1.999]
    ua = [0.0, 0.3, 0.5, 0.7, 0.8, 0.9, 0.9, 0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.30, 0.22, 0.18, 0.14, 0.1, 0.1, 0.1]
    xx = np.array([[local[0] + dv[i], local[1], local[2], 0.0001, local[4] + ds[i], 0.0001] for i in range(20)])
    uu = np.array([[0.0, ua[i]] for i in range(20)])
    return xx[:N], uu[:N]


class Plant(object):
    """vehicleSimulator.py:164-199 in curvilinear coordinates; state [vx vy wz epsi s ey], input [delta a]."""

    def __init__(self, track):
        self.track = track
        self.vx, self.vy, self.wz, self.epsi, self.s, self.ey = 0.2, 0.0, 0.0, 0.0, 0.01, 0.0  # :140-142
        self.ax = self.ay = 0.0
        self.dt = 0.005

    def state(self):
        return np.array([max(self.vx, 0.01), self.vy, self.wz, self.epsi, self.s, self.ey])  # controllerMain.py:183-184

    def step(self, delta, a, T):
        for _ in range(int(round(T / self.dt))):
            aF = aR = 0.0
            if abs(self.vx) > 0.2:
                aF = delta - math.atan((self.vy + VEH["lf"] * self.wz) / abs(self.vx))
                aR = math.atan((-self.vy + VEH["lr"] * self.wz) / abs(self.vx))
            FyF, FyR = 60.0 * aF, 60.0 * aR
            kap = oracle.curvature(self.s, self.track)
            sdot = (self.vx * math.cos(self.epsi) - self.vy * math.sin(self.epsi)) / (1.0 - self.ey * kap)
            vx, vy, wz = self.vx, self.vy, self.wz
            self.s += self.dt * sdot
            self.ey += self.dt * (vx * math.sin(self.epsi) + vy * math.cos(self.epsi))
            self.epsi += self.dt * (wz - kap * sdot)
            self.vx += self.dt * (self.ax + wz * vy)
            self.vy += self.dt * (self.ay - wz * vx)
            self.ax = a - VEH["mu"] * vx - FyF / VEH["m"] * math.sin(delta)
            self.ay = 1.0 / VEH["m"] * (FyF * math.cos(delta) + FyR)
            self.wz += self.dt * (1.0 / VEH["Iz"] * (VEH["lf"] * FyF * math.cos(delta) - VEH["lr"] * FyR))
            self.vx = abs(self.vx)


def _drive(max_ticks, until_s=None):
    """Runs the loop for `max_ticks` ticks (or until the plant's arc length passes `until_s`); returns the plant, the
    number of ticks driven and the worst deviations from the oracle."""
    N, dt = 8, 1.0 / 30.0
    track_map = lp.Map("L_shape")
    track = track_map.PointAndTangent
    Q = np.diag([100.0, 1.0, 1.0, 20.0, 0.0, 900.0])
    R = 0.25 * np.eye(2)
    dR = 37.5 * np.array([1.3, 1.0])
    ctl = lp.PathFollowingLPV_MPC(Q, R, dR, N, np.ones(N + 1), dt, track_map, "OSQP", 0, 0, params=dict(VEH, **{"/TrajectoryPlanner/max_vel": 5.0}))
    cfg = oracle.make_cfg("controller", N, dt, Q, R, dR, track)
    st = oracle.default_settings(polish=1)
    plant = Plant(track)
    cmd = np.zeros(2)
    first_it = 1
    u_pred_o = None
    worst_u = worst_x = 0.0
    ey_max = 0.0
    tick = 0
    for tick in range(max_ticks):
        if until_s is not None and plant.s >= until_s:
            break
        x = plant.state()
        ctl.OldSteering.append(float(cmd[0])); ctl.OldAccelera.append(float(cmd[1]))
        ctl.OldSteering.pop(0); ctl.OldAccelera.pop(0)
        old = ([float(cmd[0])], float(cmd[1]))
        if first_it < 10:
            xx, uu = guess_vectors(x, N)
            ctl.solve(x, xx, uu, False, np.ones(N), 0, 0, 0, first_it)
            o = oracle.ctrl_solve(cfg, st, x, mode=2, traj=xx, u_prev=uu, vel_ref=np.ones(N), old_steering=old[0], old_accel=old[1])
            first_it += 1
        else:
            vel_ref, curv_ref = np.ones(N + 1), np.zeros(N)
            states, A_L, B_L, C_L = ctl.LPVPrediction(x, ctl.uPred, vel_ref, curv_ref, 60.0, 0)
            so, Ao, Bo, Co, err = oracle.ctrl_predict(cfg, x, u_pred_o, vel_ref, curv_ref, 60.0, 0)
            assert err == 0
            np.testing.assert_allclose(states, so, rtol=0, atol=1e-7)
            ctl.solve(states[0, :], states, ctl.uPred, False, vel_ref, A_L, B_L, C_L, first_it)
            o = oracle.ctrl_solve(cfg, st, so[0, :], A=Ao, B=Bo, Cc=Co, mode=0, vel_ref=vel_ref, old_steering=old[0], old_accel=old[1])
        assert ctl.status_val == o["status"], (tick, ctl.status_val, o["status"])
        assert int(ctl.info["iters"]) == o["iter"], (tick, ctl.info["iters"], o["iter"])
        assert ctl.feasible == 1, tick
        worst_u = max(worst_u, np.abs(ctl.uPred - o["uPred"]).max())
        worst_x = max(worst_x, np.abs(ctl.xPred - o["xPred"]).max())
        assert ctl.xPred.shape == (N + 1, 6) and ctl.uPred.shape == (N, 2) and ctl.LinPoints.shape == (N + 1, 6)
        u_pred_o = o["uPred"]
        cmd = ctl.uPred[0, :].copy()           # controllerMain.py:381-383 (delays are 0)
        plant.step(cmd[0], cmd[1], dt)
        ey_max = max(ey_max, abs(plant.ey))
    return plant, tick, worst_u, worst_x, ey_max, track_map


def test_controller_closed_loop_through_the_dropin_class():
    plant, ticks, worst_u, worst_x, ey_max, _ = _drive(75)
    assert worst_u < 1e-6 and worst_x < 1e-6, (worst_u, worst_x)
    # the car accelerates towards the 1 m/s reference and stays on the centre line
    assert 0.5 < plant.vx < 1.2 and abs(plant.ey) < 0.1 and plant.s > 0.5


def test_controller_one_full_lap_through_the_dropin_class():
    """BASELINE configs[0] as stated: the single controller stepped closed loop for ONE LAP of the L-shaped track
    (19.23 m at the 1 m/s path-tracking reference: 552 ticks in the survey's probe of the reference loop),
    every tick solved by the GPU through the drop-in class and by the oracle beside it: status and iteration count
    identical on every tick, controls and predicted states within 1e-6."""
    track_len = float(lp.Map("L_shape").TrackLength)
    plant, ticks, worst_u, worst_x, ey_max, _ = _drive(700, until_s=track_len)
    assert plant.s >= track_len, (plant.s, ticks)       # the lap was completed ...
    assert 540 <= ticks <= 620, ticks                    # ... in about the reference's 552 ticks
    assert worst_u < 1e-6 and worst_x < 1e-6, (worst_u, worst_x)
    assert ey_max < 0.1 and 0.9 < plant.vx < 1.1, (ey_max, plant.vx)
