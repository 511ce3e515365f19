"""GPU parity of the LiDAR post-processing (f1) and of the batched unknown-environment step."""
import numpy as np
import pytest
import torch

from oracle import lidar as olidar, model, mpc, range_finder
from tests import helpers

pytestmark = pytest.mark.gpu


def cu(a, dt=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()


def _on_boundary(p, ring, tol=1e-9):
    a, b = ring, np.roll(ring, -1, axis=0)
    ab = b - a
    t = np.clip(((p - a) * ab).sum(1) / (ab * ab).sum(1), 0, 1)
    return np.min(np.hypot(*(a + t[:, None] * ab - p).T)) <= tol


def same_polygon(A, B, tol=1e-9):
    """Same convex polygon up to vertices that are collinear to rounding (Qhull and the monotone chain may differ there)."""
    return all(_on_boundary(p, B, tol) for p in A) and all(_on_boundary(p, A, tol) for p in B)


@pytest.mark.parametrize("sigma", (0.0, 0.01))
def test_clusters_and_hulls_match_sklearn_and_qhull(sigma):
    import ldcbf_b200 as L
    from sklearn.cluster import DBSCAN
    lid = np.load(f"{helpers.G}/lidar_golden.npz")
    scans = np.concatenate([lid[f"{m}/{r}/readings"] for m in ("circles", "crowded10", "main_paper") for r in ("r15", "r30")])
    rs = np.random.default_rng(1)
    noise = rs.normal(0, sigma, scans.shape) if sigma else None
    out = L.lidar_clusters(cu(scans), noise=None if noise is None else cu(noise), max_hulls=12, max_hull_verts=192)
    labels, verts, nverts, nobs = (out[k].cpu().numpy() for k in ("labels", "verts", "nverts", "nobs"))
    assert out["overflow"].sum().item() == 0
    for b, reads in enumerate(scans):
        valid = ~np.isnan(reads[:, 0])
        pts = reads[valid] + (noise[b][valid] if noise is not None else 0.0)
        if len(pts) == 0:
            assert (labels[b] == -1).all() and nobs[b] == 0
            continue
        ref = DBSCAN(eps=0.3, min_samples=3).fit(pts).labels_
        assert np.array_equal(labels[b][valid], ref), b               # cluster labels: exact, sklearn's numbering
        assert (labels[b][~valid] == -1).all()
        rings = range_finder.local_obstacles(np.where(valid[:, None], reads + (noise[b] if noise is not None else 0.0), np.nan))
        assert nobs[b] == len(rings), (b, nobs[b], len(rings))
        for o, ring in enumerate(rings):
            assert same_polygon(verts[b, o, :nverts[b, o]], ring), (b, o)


def test_batched_unknown_env_step_matches_oracle():
    """K4 -> f1 -> K1 -> K2+K3 on the device vs the oracle chain (lidar.cast -> sklearn DBSCAN -> Qhull -> QP)."""
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios
    geo = helpers.load_geo()
    pts = helpers.map_points(geo, "crowded10")
    rs = np.random.default_rng(2)
    B = 48
    pos = rs.uniform((-0.5, -0.5), (4.5, 4.0), (B, 2))
    state = np.column_stack((pos[:, 0], rs.uniform(-0.2, 0.2, B), pos[:, 1], rs.uniform(-0.2, 0.2, B), rs.uniform(-1, 2, B)))
    goal = np.tile([4.0, 3.5], (B, 1))
    foots = scenarios.foot_window(np.ones(B, bool), 0, 3)
    noise = rs.normal(0, 0.01, (B, 360, 2))
    verts, nverts, nobs = scenarios.pack_rings([pts] * B)
    eng = L.BatchedUnknownEnvMPC(goal, verts, nverts, nobs, lidar_range=1.5, sampling_time=0.4, max_hull_verts=64)
    out = eng.step(cu(state[:, :4]), cu(state[:, 4]), cu(foots, torch.int8), noise=cu(noise))
    assert out["sensed"]["overflow"].sum().item() == 0
    U, X, st, obj = (out[k].cpu().numpy() for k in ("U", "X", "status", "obj"))
    n_ok = 0
    for b in range(B):
        c, eta, rings, _ = range_finder.unknown_env_half_planes(pos[b], pts, 1.5, 360, noise=noise[b])
        assert int(out["sensed"]["nobs"][b].item()) == len(rings)
        r = mpc.mpc_step(state[b], goal[b], None, [int(v) for v in foots[b]], sampling_time=0.4, c_eta=(c, eta))
        dmin = min([np.hypot(*(pos[b] - ci)) for ci in c] + [1.0])
        if dmin < 1e-6:
            continue                                   # on an inferred edge: the normal is numerically undefined
        assert st[b] == r["status"], (b, st[b], r["status"])
        if r["status"] == 0:
            n_ok += 1
            assert np.abs(U[b] - r["U"]).max() <= 1e-4 and np.abs(X[b] - r["X"]).max() <= 1e-4
            assert abs(obj[b] - r["obj"]) <= 1e-6 * abs(r["obj"])
    assert n_ok >= 30


def test_unknown_env_rollout_on_device_transition_by_transition():
    """ldcbf_rollout_unknown_f64 (K4 -> f1 -> K1 -> K2+K3 -> advance per step, no host round trip): every recorded
    transition equals the oracle chain (lidar.cast -> DBSCAN -> hulls -> half-planes -> QP) on the state the device
    visited, the loop-control outputs are consistent, and the stepwise batched step reproduces the first steps."""
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import END_NAMES
    B, T = 48, 40
    c3 = scenarios.config3(B, seed=0)
    state0 = np.zeros((B, 5))
    rs = np.random.default_rng(6)
    state0[:, 0], state0[:, 2], state0[:, 4] = rs.uniform(-0.2, 0.2, B), rs.uniform(-0.2, 0.2, B), np.pi / 2
    noise = rs.normal(0, 0.01, (B, 360, 2))
    eng = L.BatchedUnknownEnvMPC(c3["goal"], c3["verts"], c3["nverts"], c3["nobs"], lidar_range=1.5, sampling_time=0.4,
                                 N_horizon=3, delta=np.full(B, 1e-6))
    state = cu(state0)
    r = eng.rollout(state, cu(np.ones(B, np.int8), torch.int8), T, noise=cu(noise))
    tX, tU, steps, status, end = (r[k].cpu().numpy() for k in ("traj_X", "traj_U", "steps", "status", "end_code"))
    assert np.array_equal(tX[:, 0], state0)
    assert np.array_equal(state.cpu().numpy(), tX[np.arange(B), steps])          # final state = last recorded row
    assert int(r["total_solves"].item()) == int(steps.sum()) + int((status != 0).sum())
    assert steps.max() >= 15
    for b in range(B):
        name = END_NAMES[end[b]]
        assert (name == "step_budget") == (steps[b] == T and status[b] == 0) or name == "stop_rule", (b, name, steps[b])
    # the first three steps through the single-step front end
    st = state0.copy()
    for k in range(3):
        foots = scenarios.foot_window(np.ones(B, bool), k, 3)
        o = eng.step(cu(st[:, :4]), cu(st[:, 4]), cu(foots, torch.int8), noise=cu(noise))
        alive = steps > k
        nxt = np.column_stack((o["X"][:, 1].cpu().numpy(), o["theta"][:, 1].cpu().numpy()))
        assert np.array_equal(nxt[alive], tX[alive, k + 1])
        st = np.where(alive[:, None], tX[:, k + 1], st)
    # oracle on sampled transitions
    n = 0
    for b in range(0, B, 2):
        pts = [np.asarray(q) for q in c3["rings"][c3["map_index"][b]]]
        s_v = model.foot_parity(T + 8, True)
        for k in rs.choice(steps[b], min(3, steps[b]), replace=False):
            pos = tX[b, k][[0, 2]]
            c, eta, rings, _ = range_finder.unknown_env_half_planes(pos, pts, 1.5, 360, noise=noise[b])
            if min([np.hypot(*(pos - ci)) for ci in c] + [1.0]) < 1e-5:
                continue
            o = mpc.mpc_step(tX[b, k], c3["goal"][b], None, s_v[k:k + 4], sampling_time=0.4, c_eta=(c, eta), delta=1e-6)
            assert o["status"] == 0, (b, k)
            assert np.abs(o["x_next"] - tX[b, k + 1]).max() <= 1e-4, (b, k)
            assert np.abs(o["U"][0] - tU[b, k, :2]).max() <= 1e-4
            n += 1
    assert n >= 40


def test_unknown_env_overflow_is_reported():
    """More clusters than max_hulls: the step and the rollout raise instead of silently dropping inferred obstacles."""
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios
    c3 = scenarios.config3(32, seed=0)
    foots = scenarios.foot_window(np.ones(32, bool), 0, 3)
    eng = L.BatchedUnknownEnvMPC(c3["goal"], c3["verts"], c3["nverts"], c3["nobs"], lidar_range=3.0, sampling_time=0.4,
                                 N_horizon=3, max_hulls=1)
    with pytest.raises(RuntimeError, match="overflow"):
        eng.step(cu(c3["state"][:, :4]), cu(c3["state"][:, 4]), cu(foots, torch.int8))
    with pytest.raises(RuntimeError, match="overflow"):
        eng.rollout(cu(c3["state"]), cu(np.ones(32, np.int8), torch.int8), 3)
