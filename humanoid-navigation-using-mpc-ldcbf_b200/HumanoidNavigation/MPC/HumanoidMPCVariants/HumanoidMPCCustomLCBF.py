"""Mirror of `MPC/HumanoidMPCVariants/HumanoidMPCCustomLCBF.py`: LDCBF with a safety margin,
h(x) = eta^T (x - c) - delta (reference :30-31).  delta is a per-scenario input of the CUDA step."""
import numpy as np

from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC


class HumanoidMPCCustomLCBF(HumanoidMPC):
    def __init__(self, goal, obstacles, N_horizon=3, N_mpc_timesteps=100, sampling_time=1e-3,
                 init_state=np.array([0, 0, 0, 0, 0]), start_with_right_foot: bool = True, verbosity: int = 1,
                 distance_from_obstacles: float = 0.0):
        assert distance_from_obstacles >= 0.0, "distance_from_obstacles must be non-negative"
        self.distance_from_obstacles = distance_from_obstacles
        super().__init__(goal, obstacles, N_horizon, N_mpc_timesteps, sampling_time, init_state,
                         start_with_right_foot, verbosity)
