"""Mirror of `report_simulations/Scenario.py`: the named maps of the reference's report, as a data catalogue.

`Scenario.load_scenario(scenario, start, goal, ...) -> (start, goal, obstacles)` with the reference's arguments
(`:27-52`); obstacles are `scipy.spatial.ConvexHull` objects like the reference returns.  Fixed maps are vertex
tables (reference `:103-231`), random maps go through the seeded generator of `Utils/obstacles.py`.
"""
from enum import Enum

import numpy as np
from scipy.spatial import ConvexHull

from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils
from HumanoidNavigation.Utils.obstacles import generate_obstacles, set_seed

_OUTER = [[[-1, -0.5], [3.5, -0.5], [-1, -1], [3.5, -1]]]          # low wall shared by both mazes
_LOWER_RIGHT = [[3.5, -1], [3.5, 0], [9, -1], [7, 2.5], [9, 2.5]]

FIXED_MAPS = {
    "HORIZONTAL_WALL": [[[1, -10], [1, 10], [3, 10], [3, -10]]],
    "VERTICAL_SLALOM": [[[1, -1], [1, 10], [2, 10], [2, -1]], [[3, 1], [3, -10], [4, -10], [4, 1]]],
    "FEW_OBSTACLES": [[[3, 2], [5, 4], [2, 2], [2, 4]], [[4, 1], [5, 0.5], [7, 3], [6, 2.5]]],
    "EMPTY": [],
    "MAIN_PAPER": [[[2.0, 7.5], [1.5, 7.0], [1.8, 6.5]],
                   [[4.0, 6.5], [4.3, 6.8], [4.7, 6.5], [4.5, 6.2], [4.1, 6.2]],
                   [[7.0, 7.0], [7.5, 7.5], [8.0, 7.0], [7.5, 6.5]],
                   [[6.0, 2.5], [6.5, 2.0], [7.0, 2.5]],
                   [[1.5, 3.0], [1.8, 3.3], [2.2, 3.0], [2.0, 2.6], [1.6, 2.6]],
                   [[2.5, 3.5], [2.8, 3.8], [3.2, 3.5], [3.0, 3.1], [2.6, 3.1]]],
    "MAZE_1": _OUTER + [[[-0.5, -0.5], [-0.5, 6], [-1, -0.5], [-1, 6]],
                        [[8.5, 2.5], [9, 2.5], [8.5, 8.5], [9, 8.5]],
                        [[3.5, 8.5], [9, 8.5], [3.5, 9], [9, 9]],
                        [[1, 1.5], [2.5, 2.5], [3.5, 3.5], [3, 5], [1, 4], [7, 4], [7, 4.5]],
                        [[5, 6.5], [8.5, 6.5], [5, 6], [8.5, 6]],
                        [[-1, 6], [3.5, 6], [-1, 9], [3.5, 9]],
                        _LOWER_RIGHT],
    "MAZE_2": _OUTER + [[[-0.5, -0.5], [-0.5, 8.5], [-1, -0.5], [-1, 8.5]],
                        [[8.5, 2.5], [9, 2.5], [8.5, 7], [9, 7]],
                        [[-1, 8.5], [5, 8.5], [-1, 9], [5, 9]],
                        [[-0.5, 2.5], [1, 2.5], [-0.5, 4.5], [1, 4.5]],
                        [[1, 2.5], [3.5, 3.5], [3, 5], [1, 4], [6, 3.5], [6, 4]],
                        [[-0.5, 6.5], [3.5, 6.5], [-0.5, 5.5], [3.5, 6]],
                        [[5, 7], [9, 7], [5, 9], [9, 9]],
                        _LOWER_RIGHT],
}
DEFAULT_ENDPOINTS = {"MAZE_1": ((0.5, 0.5), (7.5, 7.5)), "MAZE_2": ((0.5, 0.5), (0.5, 7.5))}
CIRCLES = ((10, 0.5, (5.5, -1.2)), (20, 1, (4, 2)), (25, 1.2, (1.7, 0)))


class Scenario(Enum):
    CROWDED = 0
    CROWDED_START = 1
    CROWDED_END = 2
    START_CLOSE_TO_OBSTACLE = 3
    END_CLOSE_TO_OBSTACLE = 4
    HORIZONTAL_WALL = 5
    VERTICAL_SLALOM = 6
    EMPTY = 7
    FEW_OBSTACLES = 8
    CIRCLE_OBSTACLES = 9
    MAIN_PAPER = 10
    BASE = 11
    MAZE_1 = 12
    MAZE_2 = 13

    @staticmethod
    def load_scenario(scenario, start, goal, num_max_obstacles=5, min_distance=2.0, delta=1.0, range_x=None,
                      range_y=None, seed: int = None):
        if seed is not None:
            ObstaclesUtils.set_random_seed(seed)
            set_seed(seed)
        name = scenario.name
        if name in ("CROWDED", "CROWDED_START", "CROWDED_END"):
            d = min_distance
            if name == "CROWDED":                                        # box between start and goal, shrunk by d
                xs = (start[0] + d, goal[0] - d)
                ys = (start[1] + d, goal[1] - d)
            else:                                                        # box of half-width d around one endpoint
                c = start if name == "CROWDED_START" else goal
                xs = (c[0] - d, c[0] + d)
                ys = (c[1] + d, c[1] - d)
            obstacles = generate_obstacles(start=start, goal=goal, num_obstacles=num_max_obstacles,
                                           x_range=(min(xs), max(xs)) if range_x is None else range_x,
                                           y_range=(min(ys), max(ys)) if range_y is None else range_y, delta=delta)
        elif name == "BASE":
            obstacles = generate_obstacles(start=start, goal=goal, num_obstacles=5, x_range=(0, 5), y_range=(0, 5),
                                           delta=delta)
        elif name in ("START_CLOSE_TO_OBSTACLE", "END_CLOSE_TO_OBSTACLE"):
            x = (start if name.startswith("START") else goal)[0]
            obstacles = [ConvexHull(np.array([[x + 0.1, -3], [x + 0.1, 3], [x + 0.3, 3], [x + 0.3, -3]]))]
        elif name == "CIRCLE_OBSTACLES":
            obstacles = [ObstaclesUtils.generate_circle_like_polygon(n, r, c) for n, r, c in CIRCLES]
        else:
            if name in DEFAULT_ENDPOINTS:
                start = DEFAULT_ENDPOINTS[name][0] if start is None else start
                goal = DEFAULT_ENDPOINTS[name][1] if goal is None else goal
            if name == "MAIN_PAPER":
                start, goal = (0, 0), (10, 10)
            obstacles = [ConvexHull(np.array(v)) for v in FIXED_MAPS[name]]
        return start, goal, obstacles
