"""GPU parity of the solver variants behind one ABI: the initial active-set guess (default), the cold start
(LDCBF_FLAG_COLD_START) and the warp-per-scenario kernel (LDCBF_FLAG_COOP_LANES) must return the same optimum — the
oracle's — on identical inputs.  Tolerances as in test_gpu_parity.py (BASELINE.json)."""
import numpy as np
import pytest
import torch

from oracle import mpc
from tests import helpers

pytestmark = pytest.mark.gpu

TOL_M = 1e-4
TOL_OBJ = 1e-6


@pytest.fixture(scope="module")
def L():
    import ldcbf_b200
    assert torch.cuda.is_available()
    ldcbf_b200.lib()
    return ldcbf_b200


def cu(a, dt=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()


def _step(L, sc, foots, flags, N=3, delta=None):
    prm = L.default_params(0.4, flags=flags)
    out = L.mpc_step(prm, cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8),
                     cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32),
                     delta=None if delta is None else cu(delta))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _against_oracle(out, sc, foots, N, delta=None):
    n_ok = 0
    for b in range(len(sc["state"])):
        r = mpc.mpc_step(sc["state"][b], sc["goal"][b], sc["rings"][b], [int(v) for v in foots[b]], N=N, sampling_time=0.4,
                         delta=0.0 if delta is None else float(delta[b]))
        assert out["status"][b] == r["status"], (b, out["status"][b], r["status"])
        if r["status"] == 0:
            n_ok += 1
            assert np.abs(out["U"][b] - r["U"]).max() <= TOL_M and np.abs(out["X"][b] - r["X"]).max() <= TOL_M, b
            assert abs(out["obj"][b] - r["obj"]) <= TOL_OBJ * max(1.0, abs(r["obj"])), b
        else:
            assert np.all(np.isnan(out["U"][b]))
    return n_ok


@pytest.mark.parametrize("N", (1, 2, 3))
def test_coop_lanes_match_oracle(L, N):
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import FLAG_COOP_LANES
    sc = scenarios.config2(160, seed=40 + N)
    foots = scenarios.foot_window(sc["right_first"], 0, N)
    delta = np.where(np.arange(160) % 3 == 0, 0.2, 0.0)
    out = _step(L, sc, foots, FLAG_COOP_LANES, N=N, delta=delta)
    assert _against_oracle(out, sc, foots, N, delta) > 120


def test_coop_lanes_on_reference_trajectories_and_statuses(L):
    """The reference's own per-step inputs (both circle runs) through the warp-per-scenario kernel, plus the
    infeasible (CoM inside the margin) and degenerate (CoM on an edge) statuses."""
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import FLAG_COOP_LANES
    rings, states, goals, foots, deltas = helpers.golden_step_inputs()
    verts, nverts, nobs = scenarios.pack_rings([rings] * len(states))
    sc = dict(state=states, goal=goals, verts=verts, nverts=nverts, nobs=nobs, rings=[rings] * len(states))
    out = _step(L, sc, foots, FLAG_COOP_LANES, delta=deltas)
    assert _against_oracle(out, sc, foots, 3, deltas) == len(states)
    # delta larger than the clearance: the constant k = 0 row is violated -> status 2, NaN outputs
    big = np.full(len(states), 5.0)
    out = _step(L, sc, foots, FLAG_COOP_LANES, delta=big)
    assert np.all(out["status"] == 2) and np.all(np.isnan(out["obj"]))


def test_coop_lanes_eight_obstacles_ragged(L):
    """The MO = 8 instantiation with ragged obstacle counts (including none)."""
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import FLAG_COOP_LANES
    rs = np.random.default_rng(12)
    base = scenarios.config2(64, seed=77)
    rings_list = []
    for b in range(64):
        n = int(rs.integers(0, 9))
        rings = []
        for o in range(n):
            ctr = rs.uniform((1.5, -3.0), (5.0, 3.0))
            ang = np.sort(rs.uniform(0, 2 * np.pi, 6))
            rings.append(ctr + 0.25 * np.column_stack((np.cos(ang), np.sin(ang))))
        rings_list.append(rings)
    verts, nverts, nobs = scenarios.pack_rings(rings_list, max_obs=8)
    sc = dict(state=base["state"], goal=base["goal"], verts=verts, nverts=nverts, nobs=nobs, rings=rings_list)
    foots = scenarios.foot_window(base["right_first"], 0, 3)
    out = _step(L, sc, foots, FLAG_COOP_LANES)
    assert _against_oracle(out, sc, foots, 3) > 30


def test_guess_cold_and_coop_agree_at_full_size(L):
    """Config 2 at its full size (4096): the three variants agree with each other to 1e-5 m (a vertex with multipliers
    ~3e4 sits in this batch, see test_solver_host_build.py), the guess needs fewer than half the iterations."""
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import FLAG_COLD_START, FLAG_COOP_LANES
    sc = scenarios.config2(4096, seed=0)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    guess = _step(L, sc, foots, 0)
    cold = _step(L, sc, foots, FLAG_COLD_START)
    assert np.array_equal(guess["status"], cold["status"])
    ok = cold["status"] == 0
    assert ok.sum() > 4000
    assert np.abs(guess["U"][ok] - cold["U"][ok]).max() <= 1e-5 and np.abs(guess["X"][ok] - cold["X"][ok]).max() <= 1e-5
    assert np.percentile(np.abs(guess["U"][ok] - cold["U"][ok]).max(axis=(1, 2)), 99.9) <= 1e-8
    assert np.abs(guess["obj"][ok] - cold["obj"][ok]).max() <= 1e-7 * np.abs(cold["obj"][ok]).max()
    assert guess["iters"].mean() < 0.5 * cold["iters"].mean()
    for lo in range(0, 4096, 1024):          # the cooperative kernel takes batches of at most 1024
        sub = {k: (v[lo:lo + 1024] if k != "rings" else v[lo:lo + 1024]) for k, v in sc.items() if k in
               ("state", "goal", "verts", "nverts", "nobs", "rings")}
        coop = _step(L, sub, foots[lo:lo + 1024], FLAG_COOP_LANES | FLAG_COLD_START)
        assert np.array_equal(coop["status"], cold["status"][lo:lo + 1024])
        okc = coop["status"] == 0
        assert np.abs(coop["U"][okc] - cold["U"][lo:lo + 1024][okc]).max() <= 1e-5
        assert np.array_equal(coop["theta"], cold["theta"][lo:lo + 1024])


def test_large_batch_prepare_resume_split(L):
    """Batches of 151 552 scenarios and more run as two kernels (prepare + resume with lane refill).  Config 2 tiled
    40 times (163 840 scenarios): every copy of a scenario gets the same answer, equal to the cold-start kernel's and
    to the small-batch kernel's; then the record pool is trimmed."""
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import FLAG_COLD_START
    sc = scenarios.config2(4096, seed=0)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    small = _step(L, sc, foots, 0)
    rep = 40
    big = {k: np.tile(sc[k], (rep,) + (1,) * (sc[k].ndim - 1)) for k in ("state", "goal", "verts", "nverts", "nobs")}
    fbig = np.tile(foots, (rep, 1))
    split = _step(L, big, fbig, 0)
    cold = _step(L, big, fbig, FLAG_COLD_START)
    assert L.lib().ldcbf_trim_workspace() == 0
    B = 4096 * rep
    for name in ("U", "X", "obj", "theta", "omega"):
        a = split[name].reshape((rep, 4096) + split[name].shape[1:])
        assert np.array_equal(np.nan_to_num(a, nan=-1e300), np.nan_to_num(np.broadcast_to(a[0], a.shape), nan=-1e300)), name
    assert np.array_equal(split["status"], cold["status"]) and np.array_equal(split["status"][:4096], small["status"])
    ok = cold["status"] == 0
    assert ok.sum() > 0.97 * B
    d = np.abs(split["U"][ok] - cold["U"][ok]).max(axis=(1, 2))
    assert d.max() <= 1e-5 and np.percentile(d, 99.9) <= 1e-8
    ok0 = small["status"] == 0
    assert np.abs(split["U"][:4096][ok0] - small["U"][ok0]).max() <= 1e-5
    assert np.array_equal(split["theta"][:4096], small["theta"])
    assert split["iters"].mean() < 0.5 * cold["iters"].mean()


def test_batch_size_invariance_and_empty_batch(L):
    """A scenario's answer does not depend on what else is in the batch (prefixes of one batch give bit-identical
    rows, including odd sizes that leave half-filled warps and blocks), and an empty batch is a no-op."""
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(33, seed=8)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    full = _step(L, sc, foots, 0)
    for B in (1, 2, 3, 17):
        sub = {k: v[:B] for k, v in sc.items() if k in ("state", "goal", "verts", "nverts", "nobs")}
        out = _step(L, sub, foots[:B], 0)
        for name in ("U", "X", "obj", "theta", "omega", "c_eta", "status", "iters"):
            assert np.array_equal(np.nan_to_num(out[name], nan=-1e300), np.nan_to_num(full[name][:B], nan=-1e300)), (B, name)
    empty = {k: v[:0] for k, v in sc.items() if k in ("state", "goal", "verts", "nverts", "nobs")}
    out = _step(L, empty, foots[:0], 0)
    assert out["U"].shape == (0, 3, 2) and out["status"].shape == (0,)


def test_device_step_graph_replay_reads_current_inputs(L):
    """BatchedHumanoidMPC.step(graph=True) replays the captured launches on the tensors' CURRENT contents."""
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(64, seed=4)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4)
    x0, th, ft = cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(foots, torch.int8)
    a = {k: v.clone() for k, v in eng.step(x0, th, ft, graph=True).items()}
    ref = _step(L, sc, foots, 0)
    assert np.array_equal(np.nan_to_num(a["U"].cpu().numpy(), nan=-1e300), np.nan_to_num(ref["U"], nan=-1e300))
    ok = ref["status"] == 0
    x0[torch.as_tensor(ok).cuda()] = cu(ref["X"][ok][:, 1])          # advance the state tensors in place
    th[torch.as_tensor(ok).cuda()] = cu(ref["theta"][ok][:, 1])
    ft.copy_(-ft)
    b = eng.step(x0, th, ft, graph=True)
    sc2 = dict(sc, state=np.column_stack((x0.cpu().numpy(), th.cpu().numpy())))
    ref2 = _step(L, sc2, -foots, 0)
    for name in ("U", "X", "obj", "status"):
        assert np.array_equal(np.nan_to_num(b[name].cpu().numpy(), nan=-1e300), np.nan_to_num(ref2[name], nan=-1e300)), name


def test_large_batch_split_other_shapes(L):
    """The prepare / resume pair against the cold-start refill kernel on a shape the other large-batch test does not
    touch: N = 2, eight obstacle slots with ragged counts, per-scenario margins and limit overrides, and a batch that
    is not a multiple of anything (151 552 + 37)."""
    from ldcbf_b200 import scenarios
    from ldcbf_b200.binding import FLAG_COLD_START
    rs = np.random.default_rng(21)
    base = scenarios.config2(2048, seed=31)
    B = 148 * 2 * 128 * 4 + 37
    idx = rs.integers(0, 2048, B)
    state = base["state"][idx] + np.column_stack((rs.uniform(-0.05, 0.05, (B, 1)), np.zeros((B, 3)), rs.uniform(-0.3, 0.3, (B, 1))))
    goal = base["goal"][idx] + rs.uniform(-0.2, 0.2, (B, 2))
    verts = np.zeros((B, 8, base["verts"].shape[2], 2))
    nverts = np.zeros((B, 8), np.int32)
    verts[:, :3], nverts[:, :3] = base["verts"][idx], base["nverts"][idx]
    # obstacles 3..7: copies of the first three shifted far away (never active, but they go through every code path)
    for o in range(3, 8):
        verts[:, o], nverts[:, o] = base["verts"][idx, o % 3] + np.array([30.0 + o, 25.0]), base["nverts"][idx, o % 3]
    nobs = rs.integers(0, 9, B).astype(np.int32)
    delta = np.where(rs.random(B) < 0.5, 0.0, rs.uniform(0.0, 0.2, B))
    limits = np.full((B, 6), np.nan)
    sel = rs.random(B) < 0.3
    limits[sel, 0], limits[sel, 1] = rs.uniform(2.0, 4.0, sel.sum()), rs.uniform(0.4, 0.9, sel.sum())
    foots = scenarios.foot_window(base["right_first"][idx], 0, 2)
    args = (cu(state[:, :4]), cu(state[:, 4]), cu(goal), cu(foots, torch.int8), cu(verts), cu(nverts, torch.int32),
            cu(nobs, torch.int32))
    outs = []
    for flags in (0, FLAG_COLD_START):
        o = L.mpc_step(L.default_params(0.4, flags=flags), *args, delta=cu(delta), limits=cu(limits))
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items()})
    a, c = outs
    assert np.array_equal(a["status"], c["status"]) and np.array_equal(a["theta"], c["theta"])
    ok = c["status"] == 0
    assert 0.5 * B < ok.sum()
    d = np.abs(a["U"][ok] - c["U"][ok]).max(axis=(1, 2))
    assert d.max() <= 1e-5 and np.percentile(d, 99.9) <= 1e-8
    assert np.abs(a["obj"][ok] - c["obj"][ok]).max() <= 1e-6 * np.abs(c["obj"][ok]).max()
    assert np.all(np.isnan(a["U"][~ok]))
    assert L.lib().ldcbf_trim_workspace() == 0


def test_marginally_feasible_states_on_the_gpu():
    """tests/golden/marginal_feasible_states.npz through ldcbf_mpc_qp_f64 (see test_solver_host_build.py): solved, equal to
    the oracle's optimum, in the one-thread, racing and large-batch launch shapes."""
    import os
    import ldcbf_b200 as L
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "marginal_feasible_states.npz"))["rows"]
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    for rep in (1, 1024, 13000):                       # racing kernel, racing kernel (wider blocks), one thread per scenario
        rows = np.tile(g, (rep, 1))
        out = L.mpc_qp(L.default_params(0.4), cu(rows[:, :4]), cu(rows[:, 4]), cu(rows[:, 5:7]), cu(rows[:, 7:11], torch.int8),
                       cu(rows[:, 11:23].reshape(-1, 3, 4)), cu(np.full(len(rows), 3), torch.int32), delta=cu(np.full(len(rows), 1e-6)))
        st, U = out["status"].cpu().numpy(), out["U"].cpu().numpy()
        assert (st == 0).all()
        assert np.abs(U - rows[:, 23:29].reshape(-1, 3, 2)).max() <= 1e-9
