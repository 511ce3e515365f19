"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

Tolerances are the ones BASELINE.json states: footsteps / CoM within 1e-4 m, objective within 1e-6 relative,
every LDCBF row >= -1e-6; LiDAR hits and half-planes are compared bit for bit.
"""
import numpy as np
import pytest
import torch

from oracle import halfplane, lidar, model, mpc
from tests import helpers

pytestmark = pytest.mark.gpu

TOL_M = 1e-4
TOL_OBJ = 1e-6
TOL_CBF = 1e-6


@pytest.fixture(scope="module")
def L():
    import ldcbf_b200
    assert torch.cuda.is_available()
    ldcbf_b200.lib()
    return ldcbf_b200


def cu(a, dt=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()


@pytest.mark.parametrize("name", ("circles", "crowded10", "main_paper"))
def test_halfplanes_match_oracle_and_reference(L, name):
    from ldcbf_b200 import scenarios
    geo = helpers.load_geo()
    rings = helpers.map_rings(geo, name)
    Q = geo[f"{name}/queries"]
    verts, nverts, nobs = scenarios.pack_rings([rings] * len(Q))
    ce = L.half_planes(cu(Q), cu(verts), cu(nverts, torch.int32), cu(nobs, torch.int32)).cpu().numpy()
    for qi, x in enumerate(Q):
        c, eta = halfplane.half_planes(x, rings)
        assert np.array_equal(ce[qi, :, :2], c), (name, qi)          # default mode: bit-equal to the oracle
        assert np.array_equal(ce[qi, :, 2:], eta), (name, qi)
    # and within 1e-12 of the reference's own output (its edge order is ConvexHull.simplices, ours the vertex ring)
    np.testing.assert_allclose(ce[:, :, :2], geo[f"{name}/c"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ce[:, :, 2:], geo[f"{name}/eta"], rtol=0, atol=1e-12)


def test_fast_geometry_flag_within_1e12(L):
    """LDCBF_FLAG_FAST_GEOMETRY: same half-planes to 1e-12 through the full step entry point."""
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import FLAG_FAST_GEOMETRY
    sc = scenarios.config2(512, seed=9)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    args = (cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8), cu(sc["verts"]),
            cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32))
    a = {k: v.cpu().numpy() for k, v in L.mpc_step(L.default_params(0.4), *args).items()}
    b = {k: v.cpu().numpy() for k, v in L.mpc_step(L.default_params(0.4, flags=FLAG_FAST_GEOMETRY), *args).items()}
    np.testing.assert_allclose(a["c_eta"], b["c_eta"], rtol=0, atol=1e-12)
    assert np.array_equal(a["status"], b["status"])
    ok = a["status"] == 0
    np.testing.assert_allclose(a["U"][ok], b["U"][ok], rtol=0, atol=1e-8)


def _run_steps(L, states, goals, foots, rings_list, deltas, sampling_time=0.4, N=3):
    from ldcbf_b200 import scenarios
    verts, nverts, nobs = scenarios.pack_rings(rings_list)
    prm = L.default_params(sampling_time)
    out = L.mpc_step(prm, cu(states[:, :4]), cu(states[:, 4]), cu(goals), cu(foots, torch.int8), cu(verts),
                     cu(nverts, torch.int32), cu(nobs, torch.int32), delta=cu(deltas))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}, nobs


def _compare(out, ref, nobs, foots, deltas):
    conf = model.default_conf()
    B = len(ref)
    n_ok = 0
    for b, r in enumerate(ref):
        assert out["status"][b] == r["status"], (b, out["status"][b], r["status"])
        np.testing.assert_allclose(out["theta"][b], r["theta"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(out["omega"][b], r["omega"], rtol=0, atol=1e-12)
        if r["status"] != 0:
            assert np.all(np.isnan(out["U"][b]))
            continue
        n_ok += 1
        assert np.abs(out["U"][b] - r["U"]).max() <= TOL_M, (b, np.abs(out["U"][b] - r["U"]).max())
        assert np.abs(out["X"][b] - r["X"]).max() <= TOL_M
        assert abs(out["obj"][b] - r["obj"]) <= TOL_OBJ * max(1.0, abs(r["obj"])), (b, out["obj"][b], r["obj"])
        # an exact solver can only be at least as good as the NNLS oracle
        assert out["obj"][b] <= r["obj"] * (1 + 1e-9) + 1e-9
    ok = out["status"] == 0
    worst, cbf = helpers.check_constraints(out["U"][ok], out["X"][ok], out["theta"][ok], out["omega"][ok],
                                           foots[ok].astype(float), out["c_eta"][ok], nobs[ok], deltas[ok], conf)
    assert worst <= 1e-8 and cbf <= TOL_CBF, (worst, cbf)
    return n_ok


def test_step0_known_answer(L):
    geo = helpers.load_geo()
    rings = helpers.map_rings(geo, "circles")
    out, _ = _run_steps(L, np.array([[0, 0, 3, 0, 0.]]), np.array([[6., -3]]), np.array([[1, -1, 1, -1]]), [rings],
                        np.array([0.0]))
    assert out["status"][0] == 0
    np.testing.assert_allclose(out["U"][0], [[-0.033515786476, 3.074353039892], [0.041962248849, 2.79245422353],
                                             [0.164198575344, 2.937686235951]], atol=1e-9)
    assert abs(out["obj"][0] - 279.52722776098) < 1e-9 * 279.5
    np.testing.assert_allclose(out["X"][0, 1], [0.02992878541529, 0.1687237253599, 2.933604536551, -0.3743048635156],
                               atol=1e-9)


def test_mpc_step_matches_oracle_on_reference_trajectories(L):
    """Identical per-step inputs taken from the reference's own IPOPT runs (delta = 0 and delta = 0.3)."""
    rings, states, goals, foots, deltas = helpers.golden_step_inputs()
    out, nobs = _run_steps(L, states, goals, foots, [rings] * len(states), deltas)
    ref = helpers.oracle_steps(states, goals, foots, [rings] * len(states), deltas)
    assert _compare(out, ref, nobs, foots, deltas) == len(states)
    # distance to the reference's own (tol=1e-5) IPOPT next states: report-level bound, SURVEY.md §8c
    nxt = np.concatenate([np.load(f"{helpers.G}/{f}")["X"][:, 1:].T for f in ("circles_traj.npz", "circles_delta_traj.npz")])
    d = np.abs(out["X"][:, 1][:, [0, 2]] - nxt[:, [0, 2]]).max(1)
    assert d[0] < 1e-6 and np.median(d) < 5e-4 and d.max() < 5e-3


def test_mpc_step_matches_oracle_on_config2_scenarios(L):
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(256, seed=3)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    deltas = np.where(np.arange(256) % 2 == 0, 0.0, 0.2)
    out, nobs = _run_steps(L, sc["state"], sc["goal"], foots, sc["rings"], deltas)
    ref = helpers.oracle_steps(sc["state"], sc["goal"], foots, sc["rings"], deltas)
    assert _compare(out, ref, nobs, foots, deltas) > 200


@pytest.mark.parametrize("N", (1, 2, 4))
def test_other_horizons_match_oracle(L, N):
    """N_horizon is a constructor argument of the reference (maze runs use N=2, simulation_maze.py:33)."""
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(96, seed=20 + N)
    foots = scenarios.foot_window(sc["right_first"], 0, N)
    verts, nverts, nobs = sc["verts"], sc["nverts"], sc["nobs"]
    prm = L.default_params(0.4)
    out = L.mpc_step(prm, cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8), cu(verts),
                     cu(nverts, torch.int32), cu(nobs, torch.int32))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    n_ok = 0
    for b in range(96):
        r = mpc.mpc_step(sc["state"][b], sc["goal"][b], sc["rings"][b], [int(v) for v in foots[b]], N=N, sampling_time=0.4)
        assert out["status"][b] == r["status"]
        if r["status"] == 0:
            n_ok += 1
            assert np.abs(out["U"][b] - r["U"]).max() <= TOL_M and np.abs(out["X"][b] - r["X"]).max() <= TOL_M
            assert abs(out["obj"][b] - r["obj"]) <= TOL_OBJ * abs(r["obj"])
    assert n_ok > 80


def test_many_obstacles_and_per_scenario_limits(L):
    """8 obstacles (the MO = 8 instantiation), ragged obstacle counts incl. zero, and per-scenario ALPHA / V_MAX /
    OMEGA overrides (bounds_tuning.py:22-26 mutates these between runs)."""
    from ldcbf_b200 import scenarios
    rs = np.random.default_rng(5)
    B = 64
    rings_all, states, goals = [], [], []
    for b in range(B):
        n = int(rs.integers(0, 9))
        rings = []
        for i in range(n):
            c = np.array([1.5 + 1.2 * (i % 4), -1.5 + 2.2 * (i // 4)]) + rs.uniform(-0.2, 0.2, 2)
            rings.append(scenarios.circle_ring(int(rs.integers(4, 12)), 0.3, c))
        rings_all.append(rings)
        states.append([-0.5, 0, rs.uniform(-2, 2), 0, rs.uniform(-1, 1)])
        goals.append([7.0, rs.uniform(-2, 2)])
    states, goals = np.array(states), np.array(goals)
    foots = scenarios.foot_window(np.ones(B, bool), 0, 3)
    verts, nverts, nobs = scenarios.pack_rings(rings_all, 8, 11)
    limits = np.full((B, 6), np.nan)          # (ALPHA, V_MAX[0], V_MAX[1], OMEGA_MAX, OMEGA_MIN, reserved)
    limits[::2, :5] = np.column_stack((rs.uniform(1.0, 4.0, B // 2), rs.uniform(0.5, 0.9, B // 2),
                                       rs.uniform(0.25, 0.45, B // 2), rs.uniform(0.2, 0.6, B // 2),
                                       -rs.uniform(0.2, 0.6, B // 2)))
    prm = L.default_params(0.4)
    out = L.mpc_step(prm, cu(states[:, :4]), cu(states[:, 4]), cu(goals), cu(foots, torch.int8), cu(verts),
                     cu(nverts, torch.int32), cu(nobs, torch.int32), limits=cu(limits))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    n_ok = 0
    for b in range(B):
        conf = model.default_conf()
        if b % 2 == 0:
            conf["ALPHA"], conf["V_MAX"], conf["OMEGA_MAX"], conf["OMEGA_MIN"] = (limits[b, 0], [limits[b, 1], limits[b, 2]],
                                                                                  limits[b, 3], limits[b, 4])
        r = mpc.mpc_step(states[b], goals[b], rings_all[b], [int(v) for v in foots[b]], sampling_time=0.4, conf=conf)
        assert out["status"][b] == r["status"], b
        if r["status"] == 0:
            n_ok += 1
            assert np.abs(out["U"][b] - r["U"]).max() <= TOL_M, b
            assert abs(out["obj"][b] - r["obj"]) <= TOL_OBJ * abs(r["obj"])
    assert n_ok > 40


def test_twenty_obstacle_known_map(L):
    """CROWDED map as a known map (20 obstacles): 8 in registers, 12 streamed from c_eta."""
    geo = helpers.load_geo()
    rings = helpers.map_rings(geo, "crowded10")
    rs = np.random.default_rng(8)
    B = 128
    pos = rs.uniform((-0.8, -0.8), (5.0, 4.5), (B, 2))
    states = np.column_stack((pos[:, 0], rs.uniform(-0.2, 0.2, B), pos[:, 1], rs.uniform(-0.2, 0.2, B), rs.uniform(-2, 2, B)))
    goals = np.tile([4.0, 3.5], (B, 1))
    foots = np.tile([1, -1, 1, -1], (B, 1)).astype(np.int8)
    deltas = np.where(np.arange(B) % 2 == 0, 0.0, 0.05)
    out, nobs = _run_steps(L, states, goals, foots, [rings] * B, deltas)
    ref = helpers.oracle_steps(states, goals, foots, [rings] * B, deltas)
    assert _compare(out, ref, nobs, foots, deltas) > 60


def test_packed_step_and_host_path_equal_the_plain_step(L):
    """ldcbf_mpc_step_packed_f64 (one state row in, one result row out) and BatchedHumanoidMPC.step_host give the same
    numbers as the plain step."""
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(300, seed=13)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    verts, nverts, nobs = cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32)
    prm = L.default_params(0.4)
    a = L.mpc_step(prm, cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8), verts, nverts, nobs)
    state6 = np.column_stack((sc["state"], foots[:, 0].astype(np.float64)))
    b = L.mpc_step_packed(prm, cu(state6), cu(sc["goal"]), verts, nverts, nobs)
    nxt = b["next"].cpu().numpy()
    st = a["status"].cpu().numpy()
    ok = st == 0
    assert np.array_equal(nxt[:, 9].astype(np.int32), st)
    assert np.array_equal(nxt[ok, :4], a["X"][:, 1].cpu().numpy()[ok])
    assert np.array_equal(nxt[ok, 4], a["theta"][:, 1].cpu().numpy()[ok])
    assert np.array_equal(nxt[ok, 5:7], a["U"][:, 0].cpu().numpy()[ok])
    assert np.array_equal(nxt[ok, 7], a["omega"][:, 0].cpu().numpy()[ok])
    assert np.array_equal(nxt[ok, 8], a["obj"].cpu().numpy()[ok])
    assert np.array_equal(b["iters"].cpu().numpy(), a["iters"].cpu().numpy())
    eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4)
    pinned = torch.as_tensor(state6).pin_memory()
    h = eng.step_host(pinned).numpy()
    assert np.array_equal(h[ok], nxt[ok]) and np.array_equal(h[:, 9], nxt[:, 9])
    # second call with the same buffer replays the captured CUDA graph: it must read the buffer's CURRENT content
    # (a loop that advances its state in place) ...
    adv = np.where(ok[:, None], np.column_stack((nxt[:, :5], -state6[:, 5])), state6)
    pinned.copy_(torch.as_tensor(adv))
    h2 = eng.step_host(pinned).numpy().copy()
    c = L.mpc_step_packed(prm, cu(adv), cu(sc["goal"]), verts, nverts, nobs)["next"].cpu().numpy()
    assert np.array_equal(np.nan_to_num(h2, nan=-1e300), np.nan_to_num(c, nan=-1e300))
    # ... and a different buffer takes the plain path again with the same result
    h3 = eng.step_host(torch.as_tensor(adv).pin_memory()).numpy()
    assert np.array_equal(np.nan_to_num(h3, nan=-1e300), np.nan_to_num(c, nan=-1e300))


def test_infeasible_and_degenerate_status(L):
    geo = helpers.load_geo()
    rings = helpers.map_rings(geo, "circles")
    inside = rings[2].mean(0)
    on_edge = 0.5 * (rings[1][3] + rings[1][4])
    states = np.array([[inside[0], 0, inside[1], 0, 0.0],          # inside an obstacle: k=0 row violated
                       [0, 0, 3, 0, 0.0],                          # fine
                       [2.95, 0, 0.0, 0, 0.0]])                    # 0.05 from obstacle 3 with delta 0.3
    goals = np.tile([6.0, -3.0], (3, 1))
    foots = np.tile([1, -1, 1, -1], (3, 1))
    out, _ = _run_steps(L, states, goals, foots, [rings] * 3, np.array([0.0, 0.0, 0.3]))
    assert list(out["status"]) == [2, 0, 2]
    assert np.isnan(out["U"][0]).all() and np.isfinite(out["U"][1]).all()
    ref = helpers.oracle_steps(states, goals, foots, [rings] * 3, [0.0, 0.0, 0.3])
    assert [r["status"] for r in ref] == [2, 0, 2]
    # CoM exactly on an edge -> ||x - c|| = 0 -> NaN normal (ObstaclesUtils.py:104) -> status 3
    ring = np.array([[1.0, -1.0], [2.0, -1.0], [2.0, 1.0], [1.0, 1.0]])
    out, _ = _run_steps(L, np.array([[1.0, 0, 0.0, 0, 0.0]]), np.array([[6.0, 0.0]]), foots[:1], [[ring]], np.array([0.0]))
    assert out["status"][0] == 3


@pytest.mark.parametrize("name", ("circles", "crowded10", "main_paper"))
@pytest.mark.parametrize("rn", ("r15", "r30"))
def test_lidar_bit_exact(L, name, rn):
    from ldcbf_b200 import scenarios
    geo = helpers.load_geo()
    lid = np.load(f"{helpers.G}/lidar_golden.npz")
    obstacles = helpers.map_points(geo, name)           # ConvexHull.points order, as the reference casts
    positions = lid[f"{name}/{rn}/positions"]
    rng = float(lid[f"{name}/{rn}/range"])
    verts, nverts, nobs = scenarios.pack_rings([obstacles] * len(positions))
    ho, he, xy = L.lidar_cast(cu(positions), cu(verts), cu(nverts, torch.int32), cu(nobs, torch.int32), rng, 360)
    ho, he, xy = ho.cpu().numpy(), he.cpu().numpy(), xy.cpu().numpy()
    ref = lid[f"{name}/{rn}/readings"]
    assert np.array_equal(np.isnan(xy), np.isnan(ref))
    assert np.array_equal(xy[~np.isnan(xy)], ref[~np.isnan(ref)])        # bit-equal to the REFERENCE's readings
    for pi, pos in enumerate(positions):
        o_ho, o_he, o_xy = lidar.cast(pos, obstacles, rng, 360)
        assert np.array_equal(ho[pi], o_ho) and np.array_equal(he[pi], o_he)   # hit indices bit-exact vs oracle


def test_rollout_matches_stepwise_and_oracle(L):
    """Closed loop in one launch == repeated single steps; every transition agrees with the oracle on identical inputs."""
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(64, seed=5)
    B, T = 64, 150
    verts, nverts, nobs = cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32)
    eng = L.BatchedHumanoidMPC(sc["goal"], verts, nverts, nobs, N_horizon=3, sampling_time=0.4)
    state = cu(sc["state"])
    r = eng.rollout(state, cu(sc["right_first"].astype(np.int8), torch.int8), T)
    torch.cuda.synchronize()
    tX, tU, steps = r["traj_X"].cpu().numpy(), r["traj_U"].cpu().numpy(), r["steps"].cpu().numpy()
    status = r["status"].cpu().numpy()
    # total_solves counts attempted solves: every executed step plus one failed attempt per stopped scenario
    assert int(r["total_solves"].item()) == int(steps.sum()) + int((status != 0).sum())
    assert (steps[status == 0] > 15).all()
    reached = np.hypot(tX[np.arange(B), steps, 0] - sc["goal"][:, 0], tX[np.arange(B), steps, 2] - sc["goal"][:, 1])
    assert np.median(reached[status == 0]) < 0.3
    # oracle on identical per-step inputs (sample of transitions)
    rs = np.random.default_rng(0)
    for b in rs.choice(B, 12, replace=False):
        s_v = model.foot_parity(T + 8, bool(sc["right_first"][b]))
        if steps[b] == 0:
            continue
        for k in rs.choice(steps[b], min(6, steps[b]), replace=False):
            o = mpc.mpc_step(tX[b, k], sc["goal"][b], sc["rings"][b], s_v[k:k + 4], sampling_time=0.4)
            assert o["status"] == 0
            assert np.abs(o["x_next"] - tX[b, k + 1]).max() <= TOL_M
            assert np.abs(o["U"][0] - tU[b, k, :2]).max() <= TOL_M and abs(o["omega"][0] - tU[b, k, 2]) < 1e-12


def test_full_size_properties_config2(L):
    """BASELINE config 2 at full size (B = 4096): size-independent properties of every returned solution."""
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(4096, seed=0)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    deltas = np.zeros(4096)
    out, nobs = _run_steps(L, sc["state"], sc["goal"], foots, sc["rings"], deltas)
    ok = out["status"] == 0
    assert ok.mean() > 0.95
    worst, cbf = helpers.check_constraints(out["U"][ok], out["X"][ok], out["theta"][ok], out["omega"][ok],
                                           foots[ok].astype(float), out["c_eta"][ok], nobs[ok], deltas[ok],
                                           model.default_conf())
    assert worst <= 1e-8 and cbf <= TOL_CBF
    # objective equals the reference's cost evaluated on the returned states
    J = ((out["X"][ok][:, :, [0, 2]] - sc["goal"][ok][:, None, :]) ** 2).sum((1, 2))
    np.testing.assert_allclose(out["obj"][ok], J, rtol=1e-12)
    assert out["iters"][ok].max() <= 100
