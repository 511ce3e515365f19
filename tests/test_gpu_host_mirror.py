"""GPU tests of the reference-shaped host API (HumanoidNavigation.MPC...) — the calls a user of the reference makes."""
import numpy as np
import pytest
import torch

from oracle import lidar as olidar, model, mpc
from tests import helpers

pytestmark = pytest.mark.gpu


def _hulls():
    from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils
    return [ObstaclesUtils.generate_circle_like_polygon(10, 0.5, (5.5, -1.2)),
            ObstaclesUtils.generate_circle_like_polygon(20, 1, (4, 2)),
            ObstaclesUtils.generate_circle_like_polygon(25, 1.2, (1.7, 0))]


def test_basic_simulation_matches_oracle_closed_loop():
    """Config 1 (simulation_1.py:80-102): init (0,0,3,0,0), goal (6,-3), CIRCLE_OBSTACLES, N=3, T=0.4."""
    from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC, conf
    hulls = _hulls()
    m = HumanoidMPC(N_horizon=3, N_mpc_timesteps=300, sampling_time=conf['DELTA_T'], goal=(6, -3),
                    init_state=(0, 0, 3, 0, 0), obstacles=hulls, verbosity=0)
    X, U, anim = m.run_simulation(path_to_gif=None, make_fast_plot=False, plot_animation=False, fill_animator=False)
    rings = [h.points[h.vertices] for h in hulls]
    Xo, Uo = mpc.run_simulation((6, -3), rings, (0, 0, 3, 0, 0), 3, 300, 0.4)
    assert X.shape == Xo.shape and U.shape == Uo.shape and X.shape[0] == 5 and U.shape[0] == 3
    assert 80 <= U.shape[1] <= 92                      # the reference's own run takes 86 steps
    np.testing.assert_allclose(X, Xo, atol=1e-6)
    np.testing.assert_allclose(U, Uo, atol=1e-6)
    # stepwise path (hooks in the loop) gives the same trajectory as the fused rollout kernel
    Xs, Us = m._run_stepwise()
    np.testing.assert_allclose(Xs, X, atol=1e-9)
    np.testing.assert_allclose(Us, U, atol=1e-9)


def test_delta_variant_and_substeps():
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCCustomLCBF import HumanoidMPCCustomLCBF
    hulls = _hulls()
    rings = [h.points[h.vertices] for h in hulls]
    m = HumanoidMPCCustomLCBF(N_horizon=3, N_mpc_timesteps=300, sampling_time=0.4, goal=(6, -3),
                              init_state=(0, 0, 3, 0, 0), obstacles=hulls, verbosity=0, distance_from_obstacles=0.3)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False)
    Xo, Uo = mpc.run_simulation((6, -3), rings, (0, 0, 3, 0, 0), 3, 300, 0.4, delta=0.3)
    assert X.shape == Xo.shape
    np.testing.assert_allclose(X, Xo, atol=1e-6)
    # every visited CoM keeps the margin (LDCBF rows within 1e-6)
    from oracle import halfplane
    for k in range(X.shape[1]):
        c, eta = halfplane.half_planes(X[[0, 2], k], rings)
        assert min(eta[o] @ (X[[0, 2], k] - c[o]) for o in range(3)) >= 0.3 - 1e-6
    # sampling_time = 0.1 -> mpc_step = 4 (bounds_tuning.py): heading advances every sub-step, CoM every 4th
    from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
    m = HumanoidMPC(N_horizon=3, N_mpc_timesteps=40, sampling_time=0.1, goal=(6, -3), init_state=(0, 0, 3, 0, 0),
                    obstacles=hulls, verbosity=0)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False)
    Xo, Uo = mpc.run_simulation((6, -3), rings, (0, 0, 3, 0, 0), 3, 40, 0.1)
    assert X.shape == Xo.shape
    np.testing.assert_allclose(X, Xo, atol=1e-6)
    np.testing.assert_allclose(U, Uo, atol=1e-6)


def test_unknown_environment_variant_runs_and_lidar_matches():
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCUnknownEnvironment import HumanoidMPCUnknownEnvironment
    from HumanoidNavigation.RangeFinder.range_finder_wth_polygons_dbscan import compute_lidar_readings
    from scipy.spatial import ConvexHull
    geo = helpers.load_geo()
    pts = helpers.map_points(geo, "crowded10")
    hulls = [ConvexHull(p) for p in pts]
    reads = compute_lidar_readings(np.array([0.0, 0.0]), [h.points for h in hulls], 1.5, 360)
    _, _, xy = olidar.cast(np.array([0.0, 0.0]), [h.points for h in hulls], 1.5, 360)
    for r, o in zip(reads, xy):
        assert (r is None and np.isnan(o[0])) or (r[0] == o[0] and r[1] == o[1])
    m = HumanoidMPCUnknownEnvironment(N_horizon=3, N_mpc_timesteps=60, sampling_time=0.4, goal=(4, 3.5),
                                      init_state=(0, 0, 0, 0, np.pi / 2), obstacles=hulls, verbosity=0,
                                      lidar_range=1.5, noisy=False)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False)
    assert X.shape[1] >= 20 and len(m.list_lidar_readings) >= X.shape[1] - 1
    assert np.hypot(X[0, -1] - 4, X[2, -1] - 3.5) < np.hypot(4, 3.5) - 1.0      # made progress toward the goal


def test_subgoal_sequencing_matches_oracle():
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCWithRRT import HumanoidMPCWithRRT
    from scipy.spatial import ConvexHull
    wall = ConvexHull(np.array([[2, -3], [2, 3], [3, -3], [3, 3.0]]))
    subs = [(1.0, 2.0), (2.5, 3.8), (4.0, 2.0), (5.0, 0.0)]
    m = HumanoidMPCWithRRT(goal=(5, 0), obstacles=[wall], N_horizon=3, N_mpc_timesteps=120, sampling_time=0.4, verbosity=0)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False, sub_goals=subs)
    Xo, Uo = mpc.run_subgoals(subs, [wall.points[wall.vertices]], 3, 120, 0.4)
    assert X.shape == Xo.shape and U.shape == Uo.shape
    np.testing.assert_allclose(X, Xo, atol=1e-6)
    np.testing.assert_allclose(U, Uo, atol=1e-6)
    assert np.hypot(X[0, -1] - 5, X[2, -1]) < 0.3
