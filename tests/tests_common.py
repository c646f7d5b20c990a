"""Shared constants / generators for the tests (reference tunings; SURVEY.md 8d synthetic configs)."""
import numpy as np

# controllerMain.py:139-141 / :146-148
CTRL_PT = (np.diag([100.0, 1.0, 1.0, 20.0, 0.0, 900.0]), 0.25 * np.eye(2), 1.5 * 25 * np.array([1.3, 1.0]))
CTRL_TT = (np.diag([400.0, 1.0, 1.0, 20.0, 0.0, 1100.0]), 0.0 * np.eye(2), np.array([100.0, 45.0]))
# plannerMain.py:96-99
PLAN_Q = -np.diag([-0.000000000000088, -9.703658572659423, -0.5, 0.000000000213635, -0.153591566469547])
PLAN_L = -np.array([1.00702414775175, 0.187661946033823, -0.0, 0.0, -0.0329493219494661])
PLAN_R = np.diag([0.8, 0.0])
PLAN_DR = np.array([6.0, 6.0])
