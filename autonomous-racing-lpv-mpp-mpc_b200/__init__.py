"""B200-native batched LPV-MPC QP solver: drop-in for the schedule -> build -> OSQP hot path of
euge2838/Autonomous-Racing-LPV-MPP-MPC (include/lpvmpc.h)."""
from . import _native
from ._native import (SCHED_ESTIMATE, SCHED_GIVEN, SCHED_PREDICT, STATUS_NAMES, NativeError, build,
                      default_settings)
from .controller import PathFollowingLPV_MPC
from .estimation import anfis_abc, observer_step
from .fleet import ClosedLoopFleet, PlannerFleet, fleet_start
from .handoff import ReferenceWindow, track_inputs
from .planner import LPV_MPC_Planner
from .solver import BatchResult, BatchSolver
from .track import Map, curvature

__all__ = ["BatchSolver", "BatchResult", "ClosedLoopFleet", "PlannerFleet", "fleet_start", "ReferenceWindow", "track_inputs", "anfis_abc", "observer_step", "PathFollowingLPV_MPC", "LPV_MPC_Planner", "Map", "curvature", "build",
           "default_settings", "NativeError", "STATUS_NAMES", "SCHED_GIVEN", "SCHED_PREDICT", "SCHED_ESTIMATE",
           "_native"]
import importlib
workloads = importlib.import_module(__name__ + '.workloads')
sharding = importlib.import_module(__name__ + '.sharding')
