"""Generate occupancy-grid golden vectors by running the REFERENCE's own `_build_occupancy_grid`.

Run in the BUILD container only (imports /root/reference with stub `matplotlib`, `casadi` and `rrtplanner` modules:
none of them is touched by the function executed; writes tests/golden/occupancy_golden.npz):

    python tests/golden/make_occupancy_golden.py

Reference function executed, unmodified:
* `HumanoidNavigation/MPC/HumanoidMPCVariants/HumanoidMPCWithRRT.py:21-88` _build_occupancy_grid (called unbound on a
  bare object carrying `.obstacles` and `.goal`), followed by the two library calls of `:103-108`
  (`scipy.ndimage.distance_transform_edt`, `np.exp`): the fixture keeps a checksum and samples of the clearance map.
Maps: the wall of `report_simulations/simulation_rrt.py:18-24`, CIRCLE_OBSTACLES, MAIN_PAPER, CROWDED seed 10.
"""
import os
import sys
import tempfile
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_geometry_golden import STUB          # noqa: E402


class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n): return _Any()


def _stub_module(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda n: _Any()
    sys.modules[name] = m
    return m


def main():
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "matplotlib"))
    open(os.path.join(tmp, "matplotlib", "__init__.py"), "w").write(STUB)
    sys.path[:0] = [tmp, "/root/reference"]
    for name in ("casadi", "rrtplanner"):
        _stub_module(name)
    import math
    cs = sys.modules["casadi"]                  # HumanoidMpc.py:21,34-48 evaluate these at import time
    cs.pi, cs.cosh, cs.sinh = math.pi, math.cosh, math.sinh
    cs.horzcat = lambda *a: np.array(a, dtype=float)[None, :]
    cs.vertcat = lambda *a: np.vstack(a)
    from scipy.ndimage import distance_transform_edt
    from scipy.spatial import ConvexHull
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCWithRRT import HumanoidMPCWithRRT
    from HumanoidNavigation.report_simulations.Scenario import Scenario
    from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils

    maps = {}
    maps["wall"] = ((5.0, 0.0), [ConvexHull(np.array([[2, -3], [2, 3], [3, -3], [3, 3]]))])
    _, _, circ = Scenario.load_scenario(Scenario.CIRCLE_OBSTACLES, (0, 3), (6, -3))
    maps["circles"] = ((6.0, -3.0), circ)
    _, _, mp = Scenario.load_scenario(Scenario.MAIN_PAPER, (0, 0), (10, 10))
    maps["main_paper"] = ((10.0, 10.0), mp)
    ObstaclesUtils.set_random_seed(10)
    try:
        from HumanoidNavigation.Utils.obstacles import set_seed
        set_seed(10)
    except Exception:
        pass
    _, _, cr = Scenario.load_scenario(Scenario.CROWDED, (0, 0), (4, 3.5), 20, range_x=(-1, 6), range_y=(-1, 6))
    maps["crowded10"] = ((4.0, 3.5), cr)

    out = {}
    for name, (goal, obstacles) in maps.items():
        fake = types.SimpleNamespace(obstacles=obstacles, goal=goal)
        og, fwd, inv = HumanoidMPCWithRRT._build_occupancy_grid(fake, 250)
        dist = distance_transform_edt(1 - og)
        out[f"{name}/goal"] = np.array(goal)
        out[f"{name}/n_obs"] = np.array(len(obstacles))
        for o, h in enumerate(obstacles):
            out[f"{name}/obs{o}/ring"] = h.points[h.vertices]
        out[f"{name}/og_packed"] = np.packbits(og.astype(np.uint8), axis=None)
        out[f"{name}/og_shape"] = np.array(og.shape)
        # clearance map = two library calls on og (:103-108); kept as a checksum and a few samples, not in full
        cost = np.exp(-dist)
        out[f"{name}/dist_sum"] = np.array(dist.sum())
        out[f"{name}/dist_max"] = np.array(dist.max())
        idx = np.array([[0, 0], [10, 20], [125, 137], [250, og.shape[1] - 1], [60, 200]])
        out[f"{name}/sample_idx"] = idx
        out[f"{name}/sample_dist"] = dist[idx[:, 0], idx[:, 1]]
        out[f"{name}/sample_cost"] = cost[idx[:, 0], idx[:, 1]]
        probe = np.array([[0.0, 0.0], list(goal), [1.234, -0.777], [2.5, 2.5]])
        out[f"{name}/probe_xy"] = probe
        out[f"{name}/probe_cells"] = np.array([fwd(x, y) for x, y in probe])
        out[f"{name}/probe_back"] = np.array([inv(10, 20), inv(125, 137), inv(250, og.shape[1] - 1)])
        print(name, og.shape, int(og.sum()), "occupied cells; max clearance", dist.max())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "occupancy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
