"""f3 on the GPU: `ldcbf_clearance_grid_f64` against the oracle and against the reference's own grids (goldens).
Occupancy and distances are compared bit for bit; the cost exp(-d) within 2 ulp of numpy's."""
import os

import numpy as np
import pytest
import torch

from oracle import occupancy
from tests.test_occupancy_cpu import MAPS, golden_map

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import ldcbf_b200
    assert torch.cuda.is_available()
    ldcbf_b200.lib()
    return ldcbf_b200


def cu(a, dt=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()


def run(L, goals, ring_lists, h_cap=None):
    from ldcbf_b200 import scenarios
    verts, nverts, nobs = scenarios.pack_rings(ring_lists)
    r = L.clearance_grid(cu(np.asarray(goals, dtype=np.float64)), cu(verts), cu(nverts, torch.int32),
                         cu(nobs, torch.int32), h_cap=h_cap)
    return {k: v.cpu().numpy() for k, v in r.items()}


def test_golden_maps_bit_exact(L):
    g = np.load(os.path.join(ROOT, "tests", "golden", "occupancy_golden.npz"))
    goals, rings, refs = zip(*(golden_map(g, name) for name in MAPS))
    out = run(L, goals, list(rings), h_cap=300)
    for b, name in enumerate(MAPS):
        og_ref = refs[b]
        H = og_ref.shape[1] - 1
        frame = occupancy.grid_frame(goals[b], rings[b])
        assert tuple(out["meta"][b, :5]) == frame and out["meta"][b, 5] == og_ref.sum()
        assert np.array_equal(out["og"][b, :, :H + 1], og_ref) and not out["og"][b, :, H + 1:].any()
        dist, cost = occupancy.clearance(og_ref)
        assert np.array_equal(out["dist"][b, :, :H + 1], dist)
        assert np.ascontiguousarray(out["dist"][b, :, :H + 1]).sum() == float(g[f"{name}/dist_sum"])   # checksum of the reference run
        rel = np.abs(out["cost"][b, :, :H + 1] - cost) / cost
        assert rel.max() <= 2 * np.finfo(np.float64).eps


def test_random_maps_match_oracle(L):
    from ldcbf_b200 import scenarios
    c5 = scenarios.config5(8, 16, seed=2, pool=8)
    c3 = scenarios.config3(8, seed=4, pool=8)
    goals = [(10.0, 10.0)] * 8 + [(4.0, 3.5)] * 8
    rings = [c5["rings"][i] for i in range(8)] + [c3["rings"][i] for i in range(8)]
    out = run(L, goals, rings, h_cap=320)
    for b in range(16):
        og, frame = occupancy.occupancy_grid(goals[b], rings[b])
        H = frame[4]
        assert np.array_equal(out["og"][b, :, :H + 1], og)
        dist, _ = occupancy.clearance(og)
        assert np.array_equal(out["dist"][b, :, :H + 1], dist)


def test_height_cap_is_reported(L):
    wall = np.array([[2.0, -3.0], [3.0, -3.0], [3.0, 3.0], [2.0, 3.0]])
    with pytest.raises(ValueError):
        run(L, [(5.0, 0.0)], [[wall]], h_cap=100)


def test_mirror_plans_and_walks_around_the_wall(L):
    """HumanoidMPCWithRRT without sub_goals: GPU map + clearance, host RRT*, one rollout launch over the way-points."""
    from scipy.spatial import ConvexHull
    from HumanoidNavigation.MPC.HumanoidMpc import conf
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCWithRRT import HumanoidMPCWithRRT
    wall_pts = np.array([[2, -3], [2, 3], [3, -3], [3, 3]], dtype=float)
    m = HumanoidMPCWithRRT(goal=(5, 0), obstacles=[ConvexHull(wall_pts)], N_horizon=3, N_mpc_timesteps=300,
                           sampling_time=conf["DELTA_T"], init_state=(0, 0, 0, 0, 0), verbosity=0)
    og, fwd, inv = m._build_occupancy_grid(250)
    ref_og, frame = occupancy.occupancy_grid((5.0, 0.0), [wall_pts[ConvexHull(wall_pts).vertices]])
    assert np.array_equal(og, ref_og)
    assert np.array_equal(fwd(1.234, -0.777), occupancy.to_grid(frame, 1.234, -0.777))
    sub_goals = m.plan_sub_goals(n=400)
    assert np.allclose(sub_goals[-1], (5, 0), atol=0.05)
    # way-points are free cells and consecutive ones see each other on the reference's grid
    from HumanoidNavigation.MPC.HumanoidMPCVariants.rrt_star import RRTStar
    chk = RRTStar(ref_og, np.ones_like(ref_og, dtype=np.float64))
    cells = [fwd(0, 0)] + [fwd(x, y) for x, y in sub_goals]
    assert all(chk.collision_free(a, b) for a, b in zip(cells, cells[1:]))
    # the closed loop over the planned way-points runs in one rollout launch; whether every leg is feasible for the
    # walking constraints is the reference formulation's business (a failed leg ends like :419-429 and the next starts)
    X, U, _ = m.run_simulation(path_to_gif=None, make_fast_plot=False, fill_animator=False)
    assert X.shape[0] == 5 and U.shape[0] == 3 and X.shape[1] >= U.shape[1] + 1
    assert np.array_equal(m.sub_goals, m.plan_sub_goals())
    inside = (X[0] > 2) & (X[0] < 3) & (X[2] > -3) & (X[2] < 3)
    assert not inside.any()
