"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol include/*.h declares, validates
arguments without touching a device, and the Python side refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG, "ldcbf_b200", "libldcbf_b200.so")):
        g.build()
    import ldcbf_b200
    return ldcbf_b200.lib()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ldcbf_mpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ldcbf_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert getattr(lib, n) is not None
    from ldcbf_b200 import binding
    assert sorted(binding.EXPORTS) == names


def test_params_default_match_reference_config(lib):
    import yaml
    import ldcbf_b200
    from oracle import model
    p = ldcbf_b200.default_params()
    conf = model.default_conf()
    mine = yaml.safe_load(open(os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200", "HumanoidNavigation", "config.yml")))
    for key, val in (("DELTA_T", p.delta_t), ("GRAVITY_CONST", p.gravity), ("COM_HEIGHT", p.com_height), ("ALPHA", p.alpha),
                     ("L_MAX_X", p.l_max_x), ("L_MAX_Y", p.l_max_y), ("L_MIN_X", p.l_min_x), ("L_MIN_Y", p.l_min_y)):
        assert conf[key] == val == mine[key]
    assert list(p.v_min) == conf["V_MIN"] == mine["V_MIN"] and list(p.v_max) == conf["V_MAX"] == mine["V_MAX"]
    assert p.omega_max == conf["OMEGA_MAX"] and p.omega_min == conf["OMEGA_MIN"]
    assert p.foot_offset == model.FOOT_LATERAL_OFFSET and p.stop_objective == model.STOP_OBJECTIVE
    assert p.sampling_time == 1e-3          # ctor default, HumanoidMpc.py:50
    assert ldcbf_b200.abi_version() == 2


def test_argument_errors_without_device(lib):
    import ldcbf_b200
    p = ldcbf_b200.default_params()
    # B = 0 is a no-op, null pointers / bad sizes are LDCBF_E_ARG (-1), unsupported shapes LDCBF_E_SHAPE (-2)
    assert lib.ldcbf_halfplanes_f64(0, 3, 24, None, None, None, None, None, None) == 0
    assert lib.ldcbf_halfplanes_f64(4, 3, 24, None, None, None, None, None, None) == -1
    assert lib.ldcbf_halfplanes_f64(4, 0, 24, None, None, None, None, None, None) == -1
    assert lib.ldcbf_mpc_qp_f64(ctypes.byref(p), 2, 3, 3, *([None] * 16)) == -1
    assert lib.ldcbf_mpc_qp_f64(None, 2, 3, 3, *([None] * 16)) == -1
    # horizons 1..4 and 5..48 are served; beyond that LDCBF_E_SHAPE, decided before anything touches the device
    dummy = (ctypes.c_double * 8)()
    ptrs = [ctypes.cast(dummy, ctypes.c_void_p)] * 15
    for n_h, want in ((49, -2), (0, -2), (-3, -2)):
        args = ptrs[:6] + [None, None] + ptrs[:7] + [None]
        assert lib.ldcbf_mpc_qp_f64(ctypes.byref(p), 2, n_h, 3, *args) == want
    assert lib.ldcbf_mpc_step_f64(ctypes.byref(p), 0, 3, 3, 24, *([None] * 19)) == 0
    assert lib.ldcbf_mpc_step_packed_f64(ctypes.byref(p), 4, 3, 3, 24, *([None] * 11)) == -1
    assert lib.ldcbf_lidar_cast_f64(1, 0, None, 1.5, None, 3, 24, None, None, None, None, None, None, None) == -1
    assert lib.ldcbf_rollout_f64(ctypes.byref(p), 1, 3, 0, 1, 10, 3, 24, *([None] * 17)) == -1
    assert lib.ldcbf_workspace_bytes(4096, 3, 3, 24) == 0


def test_no_cpu_fallback():
    import torch
    import ldcbf_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(TypeError):
        ldcbf_b200.half_planes(torch.zeros(1, 2, dtype=torch.float64), torch.zeros(1, 1, 4, 2, dtype=torch.float64),
                               torch.ones(1, 1, dtype=torch.int32), torch.ones(1, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        ldcbf_b200.BatchedHumanoidMPC(np.zeros((1, 2)), np.zeros((1, 1, 4, 2)), np.ones((1, 1)), np.ones(1))
    from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
    with pytest.raises(RuntimeError):
        HumanoidMPC(goal=(1, 1), obstacles=[], sampling_time=0.4)


def test_product_path_never_imports_oracle():
    """The package must not route through the oracle: no file under the package imports `oracle`."""
    pkg = os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)


def test_scenario_generator_and_foot_window():
    from ldcbf_b200 import scenarios
    from oracle import model
    sc = scenarios.config2(32, seed=0)
    assert sc["verts"].shape == (32, 3, 24, 2) and list(sc["nverts"][0]) == [9, 19, 24] and (sc["nobs"] == 3).all()
    for b in range(32):
        for k in (0, 1, 7):
            want = model.foot_parity(64, bool(sc["right_first"][b]))[k:k + 4]
            assert list(scenarios.foot_window(sc["right_first"][b:b + 1], k, 3)[0]) == want
    # same vertices as the reference's generate_circle_like_polygon hull (golden), up to rotation of the ring
    g = np.load(os.path.join(ROOT, "tests", "golden", "geometry_golden.npz"))
    for o, ring in enumerate(scenarios.circle_rings()):
        ref = g[f"circles/obs{o}/points"][g[f"circles/obs{o}/vertices"]]
        assert ring.shape == ref.shape
        j = int(np.argmin(np.abs(ref - ring[0]).sum(1)))
        np.testing.assert_allclose(np.roll(ref, -j, axis=0), ring, atol=1e-15)


def test_crowded_map_reproduced():
    """f4: the mirror's seeded generators rebuild the reference's CROWDED map (seed 10, simulation_1.py:201-218) and
    its fixed maps bit for bit; the goldens were produced by the reference's own Scenario.load_scenario."""
    pytest.importorskip("torch")
    import importlib
    import torch
    if not torch.cuda.is_available():
        # the mirror package imports ldcbf_b200 (loads the .so, no device needed) — fine on CPU
        pass
    from HumanoidNavigation.report_simulations.Scenario import Scenario
    from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils
    from HumanoidNavigation.Utils.obstacles import set_seed
    g = np.load(os.path.join(ROOT, "tests", "golden", "geometry_golden.npz"))
    ObstaclesUtils.set_random_seed(10)
    set_seed(10)
    _, _, obs = Scenario.load_scenario(Scenario.CROWDED, (0, 0), (4, 3.5), 20, range_x=(-1, 6), range_y=(-1, 6))
    assert len(obs) == int(g["crowded10/n_obs"]) == 20
    for o, h in enumerate(obs):
        assert np.array_equal(h.points, g[f"crowded10/obs{o}/points"])
        assert np.array_equal(h.vertices, g[f"crowded10/obs{o}/vertices"])
    for name, key in (("MAIN_PAPER", "main_paper"), ("CIRCLE_OBSTACLES", "circles")):
        _, _, obs = Scenario.load_scenario(getattr(Scenario, name), (0, 3), (6, -3))
        assert len(obs) == int(g[f"{key}/n_obs"])
        for o, h in enumerate(obs):
            assert np.array_equal(h.points, g[f"{key}/obs{o}/points"])
    _, _, maze = Scenario.load_scenario(Scenario.MAZE_1, None, None)
    assert len(maze) == 8
