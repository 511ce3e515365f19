"""Mirror of `MPC/HumanoidMPCVariants/HumanoidMPCUnknownEnvironment.py`: the obstacle list of each step is
inferred from a LiDAR scan (reference :30-68).  Ray casting is the K4 kernel, the half-planes of the inferred hulls
the K1 kernel; clustering / hulls are host code (see RangeFinder)."""
import numpy as np
from scipy.spatial import ConvexHull

from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
from HumanoidNavigation.RangeFinder.range_finder_wth_polygons_dbscan import range_finder
from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils


class HumanoidMPCUnknownEnvironment(HumanoidMPC):
    def __init__(self, goal, obstacles, N_horizon=3, N_mpc_timesteps=100, sampling_time=1e-3,
                 init_state=np.array([0, 0, 0, 0, 0]), start_with_right_foot: bool = True, verbosity: int = 1,
                 lidar_range: float = 3.0, lidar_resolution: int = 360, noisy: bool = True):
        self.lidar_range = lidar_range
        self.lidar_resolution = lidar_resolution
        self.noisy = noisy          # the reference always adds unseeded noise (:44-50); False makes runs reproducible
        super().__init__(goal, obstacles, N_horizon=N_horizon, N_mpc_timesteps=N_mpc_timesteps,
                         sampling_time=sampling_time, init_state=init_state,
                         start_with_right_foot=start_with_right_foot, verbosity=verbosity)

    def _get_list_c_and_eta(self, x_k: float, y_k: float):
        pos = np.array([x_k, y_k])
        lidar_readings, _, inferred = range_finder(lidar_position=pos, obstacles=[ch.points for ch in self.obstacles],
                                                   lidar_range=self.lidar_range, resolution=self.lidar_resolution,
                                                   noisy=self.noisy)
        hulls = [ConvexHull(o) for o in inferred]
        self.list_inferred_obstacles.append(hulls)
        self.list_lidar_readings.append(lidar_readings)
        c, eta = ObstaclesUtils.closest_points_and_normals(pos, hulls)
        return [ci.reshape(2, 1) for ci in c], [ei.reshape(2, 1) for ei in eta]
