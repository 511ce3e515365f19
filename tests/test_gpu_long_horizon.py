"""GPU parity of the long-horizon solver (config 5 of BASELINE.json: horizon 10-40, 8-64 obstacles) against
`oracle/qp_pspace.py`, the Lawson-Hanson restatement that stays well-posed where the footstep-space Hessian of
`oracle/qp.py` is singular.  Tolerances as in test_gpu_parity.py; measured agreement is ~1e-9.
"""
import numpy as np
import pytest
import torch

from oracle import model, mpc, qp_pspace

pytestmark = pytest.mark.gpu

TOL_M = 1e-4
TOL_OBJ = 1e-6
TOL_ROW = 1e-6


@pytest.fixture(scope="module")
def L():
    import ldcbf_b200
    assert torch.cuda.is_available()
    ldcbf_b200.lib()
    return ldcbf_b200


def cu(a, dt=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()


def run(L, sc, N, delta=None):
    from ldcbf_b200 import scenarios
    foots = scenarios.foot_window(sc["right_first"], 0, N)
    out = L.mpc_step(L.default_params(0.4), cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]),
                     cu(foots, torch.int8), cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32),
                     delta=None if delta is None else cu(np.full(len(foots), delta)))
    return foots, {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("N,n_obs,B,ncheck", [(10, 8, 96, 32), (20, 16, 64, 12), (40, 32, 48, 6), (40, 64, 48, 4)])
def test_long_horizon_step_matches_pspace_oracle(L, N, n_obs, B, ncheck):
    from ldcbf_b200 import scenarios
    conf = model.default_conf()
    sc = scenarios.config5(B, n_obs, seed=N)
    foots, out = run(L, sc, N)
    assert np.all(out["status"] != 1), "iteration cap hit"
    assert (out["status"] == 0).sum() >= B // 2
    for b in range(ncheck):
        rings = sc["rings"][sc["map_index"][b]]
        ref = qp_pspace.mpc_step(sc["state"][b], sc["goal"][b], rings, foots[b], N, 0.4, conf)
        assert ref["status"] == out["status"][b], (b, ref["status"], out["status"][b])
        assert np.abs(ref["theta"] - out["theta"][b]).max() < 1e-12 and np.abs(ref["omega"] - out["omega"][b]).max() < 1e-12
        if ref["status"] != 0:
            assert np.all(np.isnan(out["U"][b])) and np.isnan(out["obj"][b])
            continue
        assert np.abs(ref["X"] - out["X"][b]).max() < TOL_M
        assert np.abs(ref["U"] - out["U"][b]).max() < TOL_M
        assert abs(ref["obj"] - out["obj"][b]) <= TOL_OBJ * ref["obj"]
    # every solved scenario of the batch: all rows of the QP hold at the returned point, and the LIP recursion holds
    ch, sob, gt = qp_pspace.lip_scalars(conf)
    for b in np.flatnonzero(out["status"] == 0)[:24]:
        rings = sc["rings"][sc["map_index"][b]]
        x0 = sc["state"][b, :4]
        q = qp_pspace.assemble(x0, out["theta"][b], out["omega"][b], foots[b], out["c_eta"][b, :, :2],
                               out["c_eta"][b, :, 2:], sc["goal"][b], conf)
        w = out["X"][b][1:, [0, 2]].reshape(-1)
        scale = np.maximum(1.0, np.linalg.norm(q["G"], axis=1))
        assert ((q["G"] @ w - q["h"]) / scale).max() < TOL_ROW
        X, U = out["X"][b], out["U"][b]
        A, Bm = model.lip_matrices(conf)
        for k in range(N):
            assert np.allclose(A @ X[k] + Bm @ U[k], X[k + 1], atol=1e-8)


def test_first_long_horizon_equals_footstep_oracle(L):
    """N = 5 is the first horizon served by the block-per-scenario kernel; the reference's own footstep-space
    formulation (oracle/qp.py) is still well conditioned there (cond 1.4e7)."""
    from ldcbf_b200 import scenarios
    conf = model.default_conf()
    sc = scenarios.config2(64, seed=5)
    foots, out = run(L, sc, 5)
    n = 0
    for b in range(32):
        ref = mpc.mpc_step(sc["state"][b], sc["goal"][b], sc["rings"][b], foots[b], 5, 0.4, conf)
        assert ref["status"] == out["status"][b]
        if ref["status"] == 0:
            n += 1
            assert np.abs(ref["U"] - out["U"][b]).max() < TOL_M
            assert np.abs(ref["X"] - out["X"][b]).max() < TOL_M
            assert abs(ref["obj"] - out["obj"][b]) <= TOL_OBJ * ref["obj"]
    assert n >= 16


def test_long_horizon_margin_and_infeasible_start(L):
    """delta-variant rows (HumanoidMPCCustomLCBF.py:30-31) and the constant k = 0 row: a start closer to an obstacle
    than delta is reported infeasible, like the oracle."""
    from ldcbf_b200 import scenarios
    conf = model.default_conf()
    sc = scenarios.config5(48, 8, seed=3)
    foots, out = run(L, sc, 10, delta=0.2)          # starts are >= 0.15 from the octagons: some violate delta = 0.2
    seen = set()
    for b in range(24):
        rings = sc["rings"][sc["map_index"][b]]
        ref = qp_pspace.mpc_step(sc["state"][b], sc["goal"][b], rings, foots[b], 10, 0.4, conf, delta=0.2)
        assert ref["status"] == out["status"][b]
        seen.add(int(ref["status"]))
        if ref["status"] == 0:
            assert np.abs(ref["X"] - out["X"][b]).max() < TOL_M
    assert 0 in seen


def test_long_horizon_packed_step_equals_plain(L):
    from ldcbf_b200 import scenarios
    sc = scenarios.config5(40, 8, seed=1)
    foots, out = run(L, sc, 12)
    state6 = np.column_stack((sc["state"], foots[:, 0].astype(np.float64)))
    pk = L.mpc_step_packed(L.default_params(0.4), cu(state6), cu(sc["goal"]), cu(sc["verts"]),
                           cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32), N=12)
    nxt = pk["next"].cpu().numpy()
    ok = out["status"] == 0
    assert np.array_equal(nxt[:, 9].astype(np.int32), out["status"])
    assert np.array_equal(nxt[ok, :4], out["X"][ok, 1]) and np.array_equal(nxt[ok, 5:7], out["U"][ok, 0])
    assert np.array_equal(nxt[ok, 8], out["obj"][ok])


def test_mirror_runs_a_long_horizon_closed_loop(L):
    """HumanoidMPC(N_horizon=10) walks the circles map; every step equals the oracle's step on the same state."""
    from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
    from scipy.spatial import ConvexHull
    from ldcbf_b200 import scenarios
    conf = model.default_conf()
    rings = scenarios.circle_rings()
    m = HumanoidMPC(goal=(6, -3), obstacles=[ConvexHull(r) for r in rings], N_horizon=10, N_mpc_timesteps=12,
                    sampling_time=conf["DELTA_T"], init_state=(0, 0, 3, 0, 0), verbosity=0)
    X, U, _ = m.run_simulation(path_to_gif=None, make_fast_plot=False, fill_animator=False)
    assert X.shape[1] >= 8
    s_v = model.foot_parity(40, True)
    for k in range(U.shape[1]):
        ref = qp_pspace.mpc_step(X[:, k], (6, -3), rings, s_v[k:k + 11], 10, 0.4, conf)
        assert ref["status"] == 0
        assert np.abs(ref["x_next"] - X[:, k + 1]).max() < TOL_M
        assert np.abs(ref["U"][0] - U[:2, k]).max() < TOL_M
