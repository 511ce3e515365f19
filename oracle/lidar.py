"""Oracle: 2-D LiDAR ray casting with (obstacle, edge) hit indices.  TEST INFRASTRUCTURE ONLY.

Instrumented restatement of
* `/root/reference/HumanoidNavigation/RangeFinder/range_finder_wth_polygons_dbscan.py:26-63`
  (`compute_lidar_readings`) and `:13-23` (`get_closest_point`),
* `/root/reference/HumanoidNavigation/Utils/obstacles.py:95-139`
  (`line_polygon_intersection` / `compute_intersection`).

Every operation is a scalar IEEE fp64 operation in the reference's order (no FMA contraction except inside
numpy's 2-element dot of `np.linalg.norm`, `model.dot2`); the
comparisons keep the reference's strictness (`denom == 0`, `0 <= ua <= 1`, `curr < distance`,
`distance < min_distance`) and first-wins tie order (edges inside an obstacle, then obstacles).
Pinned bit-for-bit against the reference's readings in `tests/test_oracle_golden.py`.
"""
import math

import numpy as np

from .model import dot2


def ray_table(lidar_range, resolution=360):
    """(resolution, 2) table of lidar_range * (cos, sin)(i * 2*pi/resolution), host libm as in `:28-37`."""
    step = 2 * math.pi / resolution
    return np.array([[lidar_range * math.cos(i * step), lidar_range * math.sin(i * step)]
                     for i in range(resolution)])


def cast(position, obstacles, lidar_range, resolution=360, rays=None):
    """Returns hit_obs[R] (int32, -1 = no hit), hit_edge[R] (int32), hit_xy[R,2] (NaN = no hit).

    obstacles: list of (n,2) arrays; edges are (i, i+1 mod n) over the rows as given
    (`obstacles.py:127-134` with an ndarray polygon).
    """
    px, py = float(position[0]), float(position[1])
    rays = ray_table(lidar_range, resolution) if rays is None else rays
    R = len(rays)
    hit_obs = np.full(R, -1, dtype=np.int32)
    hit_edge = np.full(R, -1, dtype=np.int32)
    hit_xy = np.full((R, 2), np.nan)
    obs = [np.asarray(o, dtype=np.float64) for o in obstacles]
    for r in range(R):
        b1x, b1y = px + float(rays[r][0]), py + float(rays[r][1])      # ray end (`:39`)
        a1x, a1y = px, py
        min_distance = lidar_range
        for oi, o in enumerate(obs):
            n = len(o)
            distance = lidar_range                                      # get_closest_point `:15`
            best = None
            for e in range(n):
                a2x, a2y = float(o[e][0]), float(o[e][1])
                b2x, b2y = float(o[(e + 1) % n][0]), float(o[(e + 1) % n][1])
                denom = (b2y - a2y) * (b1x - a1x) - (b2x - a2x) * (b1y - a1y)
                if denom == 0:
                    continue
                ua = ((b2x - a2x) * (a1y - a2y) - (b2y - a2y) * (a1x - a2x)) / denom
                ub = ((b1x - a1x) * (a1y - a2y) - (b1y - a1y) * (a1x - a2x)) / denom
                if 0 <= ua <= 1 and 0 <= ub <= 1:
                    x = a1x + ua * (b1x - a1x)
                    y = a1y + ua * (b1y - a1y)
                    dx, dy = x - px, y - py
                    curr = math.sqrt(dot2(dx, dy, dx, dy))                # np.linalg.norm of a 2-vector
                    if curr < distance:
                        distance = curr
                        best = (x, y, e)
            if best is None:
                continue
            if distance <= lidar_range and distance < min_distance:     # `:57`
                min_distance = distance
                hit_obs[r], hit_edge[r] = oi, best[2]
                hit_xy[r] = best[0], best[1]
    return hit_obs, hit_edge, hit_xy
