"""Mirror of `MPC/HumanoidMPCVariants/HumanoidMPCCustomLCBF.py`.

The variant keeps the CoM at least `distance_from_obstacles` away from every obstacle by shifting each LDCBF row:
h(x) = eta^T (x - c) - delta >= 0 (reference `:30-31`).  In this package delta is simply a per-scenario input of the
CUDA step (`delta` of `ldcbf_mpc_qp_f64`), so the subclass only has to record it before the base constructor runs.
A margin of 1e-6 is also the recommended setting for plain closed-loop runs with the exact solver.
"""
import numpy as np

from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC


class HumanoidMPCCustomLCBF(HumanoidMPC):
    def __init__(self, goal, obstacles, N_horizon=3, N_mpc_timesteps=100, sampling_time=1e-3,
                 init_state=np.array([0, 0, 0, 0, 0]), start_with_right_foot: bool = True, verbosity: int = 1,
                 distance_from_obstacles: float = 0.0):
        if distance_from_obstacles < 0.0:
            raise AssertionError("distance_from_obstacles must be non-negative")
        self.distance_from_obstacles = float(distance_from_obstacles)      # read by HumanoidMPC._solve / _run_fused
        HumanoidMPC.__init__(self, goal, obstacles, N_horizon=N_horizon, N_mpc_timesteps=N_mpc_timesteps,
                             sampling_time=sampling_time, init_state=init_state,
                             start_with_right_foot=start_with_right_foot, verbosity=verbosity)
