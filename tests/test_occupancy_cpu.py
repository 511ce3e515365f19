"""f3 on the CPU: the occupancy/clearance oracle against outputs of the reference's own `_build_occupancy_grid`
(tests/golden/occupancy_golden.npz, made by tests/golden/make_occupancy_golden.py), and the host RRT* of the mirror
(property tests: the reference's planner is a third-party package that is not available, so its tree is unpinned)."""
import os

import numpy as np
import pytest

from oracle import occupancy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAPS = ("wall", "circles", "main_paper", "crowded10")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "occupancy_golden.npz"))


def golden_map(g, name):
    rings = [g[f"{name}/obs{o}/ring"] for o in range(int(g[f"{name}/n_obs"]))]
    shape = tuple(g[f"{name}/og_shape"])
    og = np.unpackbits(g[f"{name}/og_packed"])[:shape[0] * shape[1]].reshape(shape)
    return g[f"{name}/goal"], rings, og


@pytest.mark.parametrize("name", MAPS)
def test_oracle_grid_equals_reference(gold, name):
    goal, rings, ref = golden_map(gold, name)
    og, frame = occupancy.occupancy_grid(goal, rings)
    assert og.shape == ref.shape and np.array_equal(og, ref)
    dist, cost = occupancy.clearance(og)
    idx = gold[f"{name}/sample_idx"]
    assert dist.sum() == float(gold[f"{name}/dist_sum"]) and dist.max() == float(gold[f"{name}/dist_max"])
    assert np.array_equal(dist[idx[:, 0], idx[:, 1]], gold[f"{name}/sample_dist"])
    assert np.array_equal(cost[idx[:, 0], idx[:, 1]], gold[f"{name}/sample_cost"])
    cells = np.array([occupancy.to_grid(frame, x, y) for x, y in gold[f"{name}/probe_xy"]])
    assert np.array_equal(cells, gold[f"{name}/probe_cells"])
    back = np.array([occupancy.to_world(frame, 10, 20), occupancy.to_world(frame, 125, 137),
                     occupancy.to_world(frame, 250, og.shape[1] - 1)])
    assert np.array_equal(back, gold[f"{name}/probe_back"])


def test_distance_transform_is_exact_brute_force(gold):
    goal, rings, og = golden_map(gold, "crowded10")
    dist, _ = occupancy.clearance(og)
    occ = np.argwhere(og == 1)
    rng = np.random.default_rng(0)
    for i, j in rng.integers(0, [og.shape[0], og.shape[1]], size=(200, 2)):
        d2 = ((occ - (i, j)) ** 2).sum(1).min()
        assert dist[i, j] == np.sqrt(float(d2))


def test_host_rrt_star_properties(gold):
    from HumanoidNavigation.MPC.HumanoidMPCVariants.rrt_star import RRTStar
    goal, rings, og = golden_map(gold, "wall")
    _, frame = occupancy.occupancy_grid(goal, rings)
    _, cost = occupancy.clearance(og)
    start, end = occupancy.to_grid(frame, 0, 0), occupancy.to_grid(frame, goal[0], goal[1])
    planner = RRTStar(og, cost, n=400, r_rewire=80, seed=1)
    path = planner.plan(start, end)
    assert path is not None and tuple(path[0]) == tuple(start) and tuple(path[-1]) == tuple(end)
    for a, b in zip(path, path[1:]):
        assert planner.collision_free(a, b)
        assert not og[b[0], b[1]]
    for a, b in zip(path[:-2], path[1:-1]):                       # tree edges (the last one is the goal connection)
        assert np.hypot(*(a - b)) <= 80 + 1
    again = RRTStar(og, cost, n=400, r_rewire=80, seed=1).plan(start, end)
    assert len(again) == len(path) and all(tuple(p) == tuple(q) for p, q in zip(path, again))
    # the straight line start -> goal crosses the wall: the path must leave it
    assert not planner.collision_free(start, end) and len(path) > 2
