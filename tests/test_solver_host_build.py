"""CPU coverage of the kernel's solver source: csrc/mpc_qp.cuh compiled for the host (tests/cpu_harness) must agree
with the oracle.  This is test infrastructure only — the product library launches the same source on the GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import halfplane, model, mpc
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "cpu_harness")
CSRC = os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200", "csrc")


@pytest.fixture(scope="module")
def host_lib():
    so = os.path.join(HARNESS, "libqp_host.so")
    srcs = [os.path.join(HARNESS, "qp_host.cu"), os.path.join(CSRC, "mpc_qp.cuh"), os.path.join(CSRC, "ldcbf_common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                        "-I", CSRC, "-gencode", "arch=compute_100a,code=sm_100a", srcs[0], "-o", so], check=True)
    import ldcbf_b200  # noqa: F401  (LdcbfParams mirror)
    return ctypes.CDLL(so)


def host_solve(lib, states, goals, foots, c_eta, nobs, deltas, sampling_time=0.4):
    import ldcbf_b200
    from ldcbf_b200.binding import LdcbfParams
    B = len(states)
    prm = LdcbfParams()
    ldcbf_b200.lib().ldcbf_params_default(ctypes.byref(prm))
    prm.sampling_time = sampling_time
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, th, g = f64(states[:, :4]), f64(states[:, 4]), f64(goals)
    ft, ce, no, dl = np.ascontiguousarray(foots, dtype=np.int8), f64(c_eta), np.ascontiguousarray(nobs, dtype=np.int32), f64(deltas)
    out = dict(U=np.zeros((B, 3, 2)), X=np.zeros((B, 4, 4)), theta=np.zeros((B, 4)), omega=np.zeros((B, 3)),
               obj=np.zeros(B), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.qp_host_solve_n3(ctypes.byref(prm), B, ce.shape[1], P(x0), P(th), P(g), P(ft), P(ce), P(no), P(dl),
                              P(out["U"]), P(out["X"]), P(out["theta"]), P(out["omega"]), P(out["obj"]),
                              P(out["status"]), P(out["iters"]))
    assert rc == 0
    return out


def c_eta_of(states, rings_list):
    B = len(states)
    mo = max(len(r) for r in rings_list)
    ce = np.zeros((B, mo, 4))
    for b in range(B):
        c, eta = halfplane.half_planes(states[b][[0, 2]], rings_list[b])
        ce[b, :len(c), :2], ce[b, :len(c), 2:] = c, eta
    return ce, np.array([len(r) for r in rings_list], dtype=np.int32)


def compare(out, ref):
    worst = 0.0
    for b, r in enumerate(ref):
        assert out["status"][b] == r["status"], (b, out["status"][b], r["status"])
        if r["status"] == 0:
            worst = max(worst, np.abs(out["U"][b] - r["U"]).max(), np.abs(out["X"][b] - r["X"]).max())
            assert abs(out["obj"][b] - r["obj"]) <= 1e-7 * max(1.0, abs(r["obj"]))   # BASELINE tolerance is 1e-6
    return worst


def test_host_build_on_reference_trajectories(host_lib):
    rings, states, goals, foots, deltas = helpers.golden_step_inputs()
    ce, nobs = c_eta_of(states, [rings] * len(states))
    out = host_solve(host_lib, states, goals, foots, ce, nobs, deltas)
    ref = helpers.oracle_steps(states, goals, foots, [rings] * len(states), deltas)
    assert compare(out, ref) < 1e-8


def test_host_build_on_closed_loops_with_margin(host_lib):
    """Closed loops driven by the host build itself (delta = 0.3 and 0, incl. states that end infeasible)."""
    geo = helpers.load_geo()
    rings = helpers.map_rings(geo, "circles")
    for delta in (0.3, 0.0):
        state = np.array([0.0, 0, 3, 0, 0])
        s_v = model.foot_parity(400)
        for k in range(120):
            st = state[None, :]
            ce, nobs = c_eta_of(st, [rings])
            out = host_solve(host_lib, st, np.array([[6.0, -3.0]]), np.array([s_v[k:k + 4]]), ce, nobs, np.array([delta]))
            r = mpc.mpc_step(state, (6, -3), rings, s_v[k:k + 4], sampling_time=0.4, delta=delta)
            assert compare(out, [r]) < 1e-8, (delta, k)
            if r["status"] != 0 or r["obj"] < 0.05:
                break
            state = np.concatenate([out["X"][0, 1], [out["theta"][0, 1]]])
        assert k > 20


def test_host_build_on_config2_batch(host_lib):
    from ldcbf_b200 import scenarios
    sc = scenarios.config2(512, seed=0)
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    ce, nobs = c_eta_of(sc["state"], sc["rings"])
    out = host_solve(host_lib, sc["state"], sc["goal"], foots, ce, nobs, np.zeros(512))
    ref = helpers.oracle_steps(sc["state"], sc["goal"], foots, sc["rings"], np.zeros(512))
    # scenario 50 sits on an ill-conditioned vertex (two nearly anti-parallel velocity rows, multipliers ~3e4)
    # where the NNLS oracle itself only reaches a complementarity residual of 5e-5: 1e-6 there, 1e-9 elsewhere
    assert compare(out, ref) < 1e-5
    d = np.array([np.abs(out["U"][b] - r["U"]).max() for b, r in enumerate(ref) if r["status"] == 0])
    assert np.percentile(d, 99) < 1e-9
    assert out["iters"].max() <= 60


def test_host_build_with_streamed_obstacles(host_lib):
    """More than the register-resident obstacles: the CROWDED map (20 obstacles, `Scenario.py:54-69`) as a KNOWN map."""
    geo = helpers.load_geo()
    rings = helpers.map_rings(geo, "crowded10")
    rs = np.random.default_rng(4)
    B = 96
    pos = rs.uniform((-0.8, -0.8), (5.0, 4.5), (B, 2))
    states = np.column_stack((pos[:, 0], rs.uniform(-0.2, 0.2, B), pos[:, 1], rs.uniform(-0.2, 0.2, B), rs.uniform(-2, 2, B)))
    goals = np.tile([4.0, 3.5], (B, 1))
    foots = np.tile([1, -1, 1, -1], (B, 1)).astype(np.int8)
    ce, nobs = c_eta_of(states, [rings] * B)
    assert ce.shape[1] == 20
    out = host_solve(host_lib, states, goals, foots, ce, nobs, np.zeros(B))
    ref = helpers.oracle_steps(states, goals, foots, [rings] * B, np.zeros(B))
    assert compare(out, ref) < 1e-6
    assert sum(r["status"] == 0 for r in ref) > 50
