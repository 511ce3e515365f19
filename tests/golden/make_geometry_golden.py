"""Generate geometry / LiDAR golden vectors by running the REFERENCE's own functions.

Run in the BUILD container only (imports /root/reference with a stub matplotlib, SURVEY.md Appendix C.3;
writes tests/golden/geometry_golden.npz and tests/golden/lidar_golden.npz):

    python tests/golden/make_geometry_golden.py

Reference functions executed, unmodified:
* `HumanoidNavigation/Utils/ObstaclesUtils.py:21-36` generate_circle_like_polygon, `:60-109`
  get_closest_point_and_normal_vector_from_obs
* `HumanoidNavigation/report_simulations/Scenario.py:27-233` load_scenario (CIRCLE_OBSTACLES, CROWDED seed 10
  with the arguments of `simulation_1.py:201-218`, MAIN_PAPER)
* `HumanoidNavigation/RangeFinder/range_finder_wth_polygons_dbscan.py:26-63` compute_lidar_readings
"""
import os
import sys
import tempfile
import textwrap

import numpy as np

STUB = textwrap.dedent('''
    import sys, types
    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Any()
        def __getattr__(self, n): return _Any()
    def _mk(name):
        m = types.ModuleType(name); m.__getattr__ = lambda n: _Any(); sys.modules[name] = m; return m
    for n in ("pyplot", "patches", "animation", "transforms", "collections", "lines"):
        setattr(sys.modules[__name__], n, _mk("matplotlib." + n))
    pm = types.ModuleType("matplotlib.path")
    class Path:
        def __init__(self, v):
            import numpy as np
            self.v = np.asarray(v, dtype=float)
        def contains_point(self, p):
            v = self.v; n = len(v); x, y = float(p[0]), float(p[1]); inside = False
            for i in range(n):
                x1, y1 = v[i]; x2, y2 = v[(i + 1) % n]
                if (y1 > y) != (y2 > y) and x < (x2 - x1) * (y - y1) / (y2 - y1) + x1: inside = not inside
            return inside
    pm.Path = Path
    sys.modules["matplotlib.path"] = pm
    path = pm
''')


def main():
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "matplotlib"))
    open(os.path.join(tmp, "matplotlib", "__init__.py"), "w").write(STUB)
    sys.path[:0] = [tmp, "/root/reference"]
    from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils
    from HumanoidNavigation.report_simulations.Scenario import Scenario
    from HumanoidNavigation.RangeFinder.range_finder_wth_polygons_dbscan import compute_lidar_readings

    out = os.path.dirname(os.path.abspath(__file__))
    rng = np.random.default_rng(20261018)

    maps = {}
    _, _, maps["circles"] = Scenario.load_scenario(Scenario.CIRCLE_OBSTACLES, start=(0, 3), goal=(6, -3))
    ObstaclesUtils.set_random_seed(10)
    from HumanoidNavigation.Utils.obstacles import set_seed
    set_seed(10)
    _, _, maps["crowded10"] = Scenario.load_scenario(Scenario.CROWDED, (0, 0), (4, 3.5), 20,
                                                     range_x=(-1, 6), range_y=(-1, 6))
    _, _, maps["main_paper"] = Scenario.load_scenario(Scenario.MAIN_PAPER, (0, 0), (10, 10))

    geo = {}
    lid = {}
    for name, hulls in maps.items():
        n_obs = len(hulls)
        geo[f"{name}/n_obs"] = np.int64(n_obs)
        allpts = np.concatenate([h.points for h in hulls])
        lo, hi = allpts.min(0) - 1.5, allpts.max(0) + 1.5
        queries = rng.uniform(lo, hi, size=(200, 2))
        # a few queries strictly inside obstacles (normal flips), a few far away
        inside = np.array([h.points[h.vertices].mean(0) + rng.normal(0, 0.02, 2) for h in hulls])
        queries = np.concatenate([queries, inside, np.array([[0.0, 3.0], [0.0, 0.0], [50.0, -40.0]])])
        geo[f"{name}/queries"] = queries
        C = np.zeros((len(queries), n_obs, 2))
        E = np.zeros((len(queries), n_obs, 2))
        for oi, h in enumerate(hulls):
            geo[f"{name}/obs{oi}/points"] = h.points
            geo[f"{name}/obs{oi}/vertices"] = h.vertices.astype(np.int32)
            geo[f"{name}/obs{oi}/simplices"] = h.simplices.astype(np.int32)
            for qi, x in enumerate(queries):
                c, eta = ObstaclesUtils.get_closest_point_and_normal_vector_from_obs(x, h, True)
                C[qi, oi] = c[:, 0]
                E[qi, oi] = eta[:, 0]
        geo[f"{name}/c"] = C
        geo[f"{name}/eta"] = E

        # LiDAR: the unknown-environment variant casts against ConvexHull.points
        # (HumanoidMPCUnknownEnvironment.py:44-50)
        obstacles = [h.points for h in hulls]
        for rng_name, lidar_range in (("r15", 1.5), ("r30", 3.0)):
            positions = rng.uniform(allpts.min(0) - 0.5, allpts.max(0) + 0.5, size=(12, 2))
            positions[0] = (0.0, 0.0)
            reads = np.full((len(positions), 360, 2), np.nan)
            for pi, pos in enumerate(positions):
                pts = compute_lidar_readings(np.array(pos), obstacles, lidar_range, 360)
                for ri, p in enumerate(pts):
                    if p is not None:
                        reads[pi, ri] = p
            lid[f"{name}/{rng_name}/range"] = np.float64(lidar_range)
            lid[f"{name}/{rng_name}/positions"] = positions
            lid[f"{name}/{rng_name}/readings"] = reads
            print(name, rng_name, "hits", int(np.isfinite(reads[..., 0]).sum()), "of", reads.shape[0] * 360)
    np.savez_compressed(os.path.join(out, "geometry_golden.npz"), **geo)
    np.savez_compressed(os.path.join(out, "lidar_golden.npz"), **lid)
    print("wrote", len(geo), "geometry arrays,", len(lid), "lidar arrays")


if __name__ == "__main__":
    main()
