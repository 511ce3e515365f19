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


def check_run(X, U, goal, rings, sampling_time, delta=0.0, right_first=True, N_simul=300, prefix_atol=1e-6):
    """Per-step parity of a closed-loop run on identical inputs.

    The closed loop amplifies perturbations by about 2x per step (a 1e-11 difference reaches 0.1 m after 30 steps,
    measured), so whole trajectories of two exact solvers legitimately part ways; parity is therefore asserted
    (a) transition by transition, feeding the oracle the states the GPU run visited, and (b) on the first steps of
    the oracle's own closed loop.  Returns the oracle's verdict on the final state."""
    from HumanoidNavigation.MPC.HumanoidMpc import INTERIOR_MARGIN
    delta = delta + INTERIOR_MARGIN          # the clearance the mirror adds in closed loop (interior-point behaviour)
    conf = model.default_conf()
    sub = int(conf["DELTA_T"] / sampling_time) or 1
    K = U.shape[1]
    assert X.shape == (5, K + 1) and U.shape[0] == 3
    s_v = model.foot_parity(K // sub + 8, right_first)
    u_prev = None
    for k in range(K):
        if k % sub == 0:
            r = mpc.mpc_step(X[:, k], goal, rings, s_v[k // sub:k // sub + 4], sampling_time=sampling_time, delta=delta)
            assert r["status"] == 0, k
            assert np.abs(r["x_next"] - X[:, k + 1]).max() <= 1e-4, (k, np.abs(r["x_next"] - X[:, k + 1]).max())
            assert np.abs(r["U"][0] - U[:2, k]).max() <= 1e-4 and abs(r["omega"][0] - U[2, k]) <= 1e-12
            u_prev = U[:2, k]
        else:
            th, om = model.heading_schedule(X[:4, k], X[4, k], goal, 3, sampling_time, conf)
            assert np.array_equal(X[:4, k + 1], X[:4, k]) and abs(X[4, k + 1] - th[1]) <= 1e-12
            assert np.array_equal(U[:2, k], u_prev) and abs(U[2, k] - om[0]) <= 1e-12
    Xo, Uo = mpc.run_simulation(goal, rings, tuple(X[:, 0]), 3, N_simul, sampling_time, right_first, delta=delta)
    n = min(8, K, Uo.shape[1])
    np.testing.assert_allclose(X[:, :n + 1], Xo[:, :n + 1], atol=prefix_atol)
    np.testing.assert_allclose(U[:, :n], Uo[:, :n], atol=prefix_atol)
    # why did the run end?  Either the stop rule fired (previous objective < 0.05) or the next QP is infeasible.
    if K % sub == 0:
        last = mpc.mpc_step(X[:, K], goal, rings, s_v[K // sub:K // sub + 4], sampling_time=sampling_time, delta=delta)
        return last["status"]
    return 0


def test_basic_simulation_matches_oracle_closed_loop():
    """Config 1 (simulation_1.py:80-102): init (0,0,3,0,0), goal (6,-3), CIRCLE_OBSTACLES, N=3, T=0.4."""
    from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC, conf
    hulls = _hulls()
    m = HumanoidMPC(N_horizon=3, N_mpc_timesteps=300, sampling_time=conf['DELTA_T'], goal=(6, -3),
                    init_state=(0, 0, 3, 0, 0), obstacles=hulls, verbosity=0)
    X, U, anim = m.run_simulation(path_to_gif=None, make_fast_plot=False, plot_animation=False, fill_animator=False)
    rings = [h.points[h.vertices] for h in hulls]
    assert U.shape[1] >= 20
    end = check_run(X, U, (6, -3), rings, 0.4)
    # with the mirror's interior margin the basic simulation reaches the goal like the reference's (86 steps there)
    assert m.last_status == 0 and end == 0 and np.hypot(X[0, -1] - 6, X[2, -1] + 3) < 0.3 and 80 <= U.shape[1] <= 92
    # the bare QP (no margin) may instead end because the NEXT step's problem has no solution — infeasible (2) or,
    # with the CoM exactly on an obstacle edge, a degenerate normal (3, `ObstaclesUtils.py:104`); which one is decided
    # by the last bits, the oracle must agree about that final state
    import HumanoidNavigation.MPC.HumanoidMpc as hm
    keep, hm.INTERIOR_MARGIN = hm.INTERIOR_MARGIN, 0.0
    try:
        X0, U0, _ = m.run_simulation(path_to_gif=None, make_fast_plot=False, plot_animation=False, fill_animator=False)
        end0 = check_run(X0, U0, (6, -3), rings, 0.4)
        assert (m.last_status == 0 and np.hypot(X0[0, -1] - 6, X0[2, -1] + 3) < 0.3) or \
            (m.last_status in (2, 3) and end0 == m.last_status)
    finally:
        hm.INTERIOR_MARGIN = keep
    # stepwise path (subclass hooks in the loop) visits the same first states as the fused rollout kernel
    Xs, Us = m._run_stepwise()
    n = min(10, Xs.shape[1], X.shape[1])
    np.testing.assert_allclose(Xs[:, :n], X[:, :n], atol=1e-9)
    check_run(Xs, Us, (6, -3), rings, 0.4)


def test_delta_variant_and_substeps():
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCCustomLCBF import HumanoidMPCCustomLCBF
    hulls = _hulls()
    rings = [h.points[h.vertices] for h in hulls]
    m = HumanoidMPCCustomLCBF(N_horizon=3, N_mpc_timesteps=300, sampling_time=0.4, goal=(6, -3),
                              init_state=(0, 0, 3, 0, 0), obstacles=hulls, verbosity=0, distance_from_obstacles=0.3)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False)
    check_run(X, U, (6, -3), rings, 0.4, delta=0.3)
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
    assert U.shape[1] >= 40
    check_run(X, U, (6, -3), rings, 0.1, N_simul=40)


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
    ring = wall.points[wall.vertices]
    subs = [(1.0, 2.0), (2.5, 3.8), (4.0, 2.0), (5.0, 0.0)]
    m = HumanoidMPCWithRRT(goal=(5, 0), obstacles=[wall], N_horizon=3, N_mpc_timesteps=120, sampling_time=0.4, verbosity=0)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False, sub_goals=subs)
    # the concatenation repeats the junction state (HumanoidMPCWithRRT.py:178-181): split the runs back apart
    cuts = [k for k in range(X.shape[1] - 1) if np.array_equal(X[:, k], X[:, k + 1])]
    assert len(cuts) == len(subs) - 1 and X.shape[1] == U.shape[1] + len(subs)
    s0 = 0
    for g, sg in enumerate(subs):
        e = cuts[g] if g < len(cuts) else X.shape[1] - 1
        Xg, Ug = X[:, s0:e + 1], U[:, s0 - g:e - g]
        assert Xg.shape[1] == Ug.shape[1] + 1
        end = check_run(Xg, Ug, sg, [ring], 0.4, N_simul=120)     # parity restarts with the right foot per run
        # a run ends because its stop rule fired (then it is near its sub-goal) or because the next QP is
        # infeasible (the reference breaks, HumanoidMpc.py:419-429, and the next sub-goal run starts from there)
        assert end in (0, 2)
        exhausted = Ug.shape[1] == 120 - 1      # all num_inputs steps used (pressed against the wall): :458 drops a column
        if end == 0 and not exhausted:
            assert ((Xg[[0, 2], -1] - np.array(sg)) ** 2).sum() < 0.2
        s0 = e + 1


def test_rrt_variant_on_maze2_and_single_subgoal_equals_base_class():
    """MAZE_2 has 9 obstacles (simulation_maze.py runs HumanoidMPCWithRRT on it with N_horizon = 2): one more than the
    register-resident 8, so the fused rollout streams the ninth.  And a run with a single sub-goal is exactly the base
    class's run to that goal (same LDCBF margin, same start)."""
    from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCWithRRT import HumanoidMPCWithRRT
    from HumanoidNavigation.report_simulations.Scenario import FIXED_MAPS
    from scipy.spatial import ConvexHull
    hulls = [ConvexHull(np.array(v, dtype=float)) for v in FIXED_MAPS["MAZE_2"]]
    assert len(hulls) == 9
    rings = [h.points[h.vertices] for h in hulls]
    subs = [(2.0, 1.2), (4.5, 1.5), (7.0, 2.0)]
    m = HumanoidMPCWithRRT(goal=subs[-1], obstacles=hulls, N_horizon=2, N_mpc_timesteps=60, sampling_time=0.4, verbosity=0)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False, sub_goals=subs)
    assert U.shape[1] >= 10 and X.shape[1] == U.shape[1] + len(subs)
    # first run, transition by transition (horizon 2)
    cut = next(k for k in range(X.shape[1] - 1) if np.array_equal(X[:, k], X[:, k + 1]))
    s_v = model.foot_parity(cut + 8, True)
    from HumanoidNavigation.MPC.HumanoidMpc import INTERIOR_MARGIN
    for k in range(cut):
        r = mpc.mpc_step(X[:, k], subs[0], rings, s_v[k:k + 3], N=2, sampling_time=0.4, delta=INTERIOR_MARGIN)
        assert r["status"] == 0 and np.abs(r["x_next"] - X[:, k + 1]).max() <= 1e-4, k
    # one sub-goal == the base class
    wall = ConvexHull(np.array([[2, -3], [2, 3], [3, -3], [3, 3.0]]))
    a = HumanoidMPCWithRRT(goal=(1.0, 2.0), obstacles=[wall], N_horizon=3, N_mpc_timesteps=60, sampling_time=0.4, verbosity=0)
    Xa, Ua, _ = a.run_simulation(None, make_fast_plot=False, fill_animator=False, sub_goals=[(1.0, 2.0)])
    b = HumanoidMPC(goal=(1.0, 2.0), obstacles=[wall], N_horizon=3, N_mpc_timesteps=60, sampling_time=0.4,
                    init_state=(0, 0, 0, 0, 0), verbosity=0)
    Xb, Ub, _ = b.run_simulation(None, make_fast_plot=False, fill_animator=False)
    assert np.array_equal(Xa, Xb) and np.array_equal(Ua, Ub)


def test_exhausted_subgoal_run_restarts_from_its_last_returned_state():
    """A sub-goal run that uses up its num_inputs steps returns X_pred[:, :num_inputs] (HumanoidMpc.py:458) and the next
    run starts from that array's last column (HumanoidMPCWithRRT.py:178): the oracle's run_subgoals does the same."""
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCWithRRT import HumanoidMPCWithRRT
    from HumanoidNavigation.MPC.HumanoidMpc import INTERIOR_MARGIN
    subs = [(3.0, 0.5), (3.0, -2.0)]
    m = HumanoidMPCWithRRT(goal=subs[-1], obstacles=[], N_horizon=3, N_mpc_timesteps=6, sampling_time=0.4, verbosity=0)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False, sub_goals=subs)
    # first run: 6 steps used up -> 6 columns (the 7th is dropped), 5 inputs; second run starts from column 5
    assert np.array_equal(X[:, 5], X[:, 6])
    Xo, Uo = mpc.run_subgoals(subs, [], 3, 6, 0.4, True)
    assert Xo.shape == X.shape and Uo.shape == U.shape
    np.testing.assert_allclose(X, Xo, atol=1e-6)
    np.testing.assert_allclose(U, Uo, atol=1e-6)


def test_known_crowded_map_runs_fused_with_twenty_obstacles():
    """The reference's CROWDED map (20 obstacles) as a known map: 8 half-planes register-resident, 12 streamed, the
    whole loop in one launch."""
    from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCCustomLCBF import HumanoidMPCCustomLCBF
    from scipy.spatial import ConvexHull
    geo = helpers.load_geo()
    hulls = [ConvexHull(p) for p in helpers.map_points(geo, "crowded10")]
    rings = [h.points[h.vertices] for h in hulls]
    m = HumanoidMPCCustomLCBF(N_horizon=3, N_mpc_timesteps=80, sampling_time=0.4, goal=(4, 3.5), init_state=(0, 0, 0, 0, np.pi / 2),
                              obstacles=hulls, verbosity=0, distance_from_obstacles=1e-6)
    X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False)
    assert U.shape[1] >= 10
    check_run(X, U, (4, 3.5), rings, 0.4, delta=1e-6, N_simul=80)


def test_bounds_tuning_grid_in_one_launch():
    """A slice of the reference's hyper-parameter grid (bounds_tuning.py:17-26): every combination is one scenario of a
    single rollout launch; each run is checked against the oracle's closed loop under the same mutated `conf`."""
    from HumanoidNavigation.report_simulations.bounds_tuning import bounds_tuning, hyperparameter_grid
    grid = hyperparameter_grid()
    assert grid.shape == (26880, 4)
    sub = grid[::2689][:10]
    (best, best_res, score, steps) = bounds_tuning(grid=sub, N_mpc_timesteps=40, return_all=True)
    assert len(score) == len(sub) and (steps >= 0).all() and (steps > 0).sum() >= 8     # a combination may be infeasible at step 0
    # oracle for two of them: first 6 MPC steps (24 loop iterations) agree
    import ldcbf_b200
    for i in (0, 7):
        conf = model.default_conf()
        conf["V_MAX"], conf["ALPHA"], conf["OMEGA_MAX"], conf["OMEGA_MIN"] = [sub[i][0], sub[i][1]], sub[i][2], sub[i][3], -sub[i][3]
        Xo, Uo = mpc.run_simulation((5, 5), [], (0, 0, 0, 0, 0), 3, 40, 0.1, conf=conf)
        from HumanoidNavigation.MPC import HumanoidMpc
        saved = dict(HumanoidMpc.conf)
        try:
            HumanoidMpc.conf.update(ALPHA=sub[i][2], V_MAX=[sub[i][0], sub[i][1]], OMEGA_MAX=sub[i][3], OMEGA_MIN=-sub[i][3])
            m = HumanoidMpc.HumanoidMPC(N_horizon=3, N_mpc_timesteps=40, sampling_time=0.1, goal=(5, 5),
                                        init_state=(0, 0, 0, 0, 0), obstacles=[], verbosity=0)
            X, U, _ = m.run_simulation(None, make_fast_plot=False, fill_animator=False)
        finally:
            HumanoidMpc.conf.clear(); HumanoidMpc.conf.update(saved)
        n = min(24, X.shape[1], Xo.shape[1])
        np.testing.assert_allclose(X[:, :n], Xo[:, :n], atol=1e-6)
