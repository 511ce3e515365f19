"""Pin the CPU oracle against the reference: its own geometry/LiDAR outputs and its own report trajectories."""
import os

import numpy as np
import pytest

from oracle import halfplane, lidar, model, mpc, qp

G = os.path.join(os.path.dirname(__file__), "golden")
MAPS = ("circles", "crowded10", "main_paper")


@pytest.fixture(scope="module")
def geo():
    return np.load(os.path.join(G, "geometry_golden.npz"))


@pytest.fixture(scope="module")
def lid():
    return np.load(os.path.join(G, "lidar_golden.npz"))


def _hulls(geo, name):
    out = []
    for oi in range(int(geo[f"{name}/n_obs"])):
        out.append((geo[f"{name}/obs{oi}/points"], geo[f"{name}/obs{oi}/vertices"], geo[f"{name}/obs{oi}/simplices"]))
    return out


@pytest.mark.parametrize("name", MAPS)
def test_halfplane_bit_equal_to_reference_on_its_edge_list(geo, name):
    """Same edge list (ConvexHull.simplices order) -> c and eta bit-equal to ObstaclesUtils.py:60-109."""
    Q = geo[f"{name}/queries"]
    C, E = geo[f"{name}/c"], geo[f"{name}/eta"]
    for oi, (pts, verts, simp) in enumerate(_hulls(geo, name)):
        ring = pts[verts]
        edges = [tuple(s) for s in simp]
        for qi, x in enumerate(Q):
            c, eta = halfplane.closest_point_and_normal(x, pts, edges, ring)
            assert np.array_equal(c, C[qi, oi]), (name, oi, qi)
            assert np.array_equal(eta, E[qi, oi]), (name, oi, qi)


@pytest.mark.parametrize("name", MAPS)
def test_halfplane_cyclic_ring_matches_reference(geo, name):
    """The CUDA path's input convention (CCW vertex ring, edges i->i+1) agrees with the reference to 1e-12."""
    Q = geo[f"{name}/queries"]
    rings = [pts[verts] for pts, verts, _ in _hulls(geo, name)]
    for qi, x in enumerate(Q):
        c, eta = halfplane.half_planes(x, rings)
        np.testing.assert_allclose(c, geo[f"{name}/c"][qi], rtol=0, atol=1e-12)
        np.testing.assert_allclose(eta, geo[f"{name}/eta"][qi], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", MAPS)
@pytest.mark.parametrize("rn", ("r15", "r30"))
def test_lidar_hits_bit_equal_to_reference(geo, lid, name, rn):
    """hit_xy bit-equal (and hit/no-hit pattern identical) to compute_lidar_readings (`:26-63`)."""
    obstacles = [pts for pts, _, _ in _hulls(geo, name)]   # ConvexHull.points, HumanoidMPCUnknownEnvironment.py:46
    rng = float(lid[f"{name}/{rn}/range"])
    for pos, ref in zip(lid[f"{name}/{rn}/positions"], lid[f"{name}/{rn}/readings"]):
        ho, he, xy = lidar.cast(pos, obstacles, rng, 360)
        assert np.array_equal(np.isnan(xy), np.isnan(ref))
        assert np.array_equal(xy[~np.isnan(xy)], ref[~np.isnan(ref)])
        assert np.all((ho >= 0) == ~np.isnan(xy[:, 0]))
        # the reported edge really contains the hit point
        for r in np.nonzero(ho >= 0)[0]:
            o = obstacles[ho[r]]
            a, b = o[he[r]], o[(he[r] + 1) % len(o)]
            cross = (b[0] - a[0]) * (xy[r, 1] - a[1]) - (b[1] - a[1]) * (xy[r, 0] - a[0])
            assert abs(cross) < 1e-9


def _circle_rings(geo):
    return [pts[verts] for pts, verts, _ in _hulls(geo, "circles")]


def test_step0_known_answer(geo):
    """SURVEY.md §8a KAT: x0=(0,0,3,0), theta0=0, goal (6,-3), CIRCLE_OBSTACLES, delta=0."""
    r = mpc.mpc_step((0, 0, 3, 0, 0), (6, -3), _circle_rings(geo), [1, -1, 1, -1])
    assert r["status"] == 0 and max(r["sol"]["kkt"]) < 1e-9
    np.testing.assert_allclose(r["theta"], [0, -0.1960353815840031, -0.3920707631680062, -0.549401723259783], atol=1e-15)
    np.testing.assert_allclose(r["omega"], [-0.4900884539600077, -0.4900884539600077, -0.39332740022944207], atol=1e-15)
    np.testing.assert_allclose(r["c"], [[5.03015369, -1.02898993], [3.0136387, 2.16459459], [1.1, 1.03923048]], atol=5e-9)
    np.testing.assert_allclose(r["eta"], [[-0.78050029, 0.62515542], [-0.96365932, 0.26713428], [-0.48926996, 0.87213239]], atol=5e-9)
    np.testing.assert_allclose(r["U"], [[-0.033515786476, 3.074353039892], [0.041962248849, 2.79245422353],
                                        [0.164198575344, 2.937686235951]], atol=1e-9)
    assert abs(r["obj"] - 279.52722776098) < 1e-9 * 279.5
    np.testing.assert_allclose(r["x_next"][:4], [0.02992878541529, 0.1687237253599, 2.933604536551, -0.3743048635156], atol=1e-9)
    rows = sorted(i for i, _ in r["sol"]["active"])
    kinds = [r["qp"]["kinds"][i] for i in rows]
    assert kinds == [("leg", 1, 1), ("man", 0, 0), ("man", 1, 0), ("man", 2, 0), ("walk", 1, 1), ("walk", 3, 1)]


@pytest.mark.parametrize("fname,max_first2", [("circles_traj.npz", 1e-6), ("circles_delta_traj.npz", 1e-6)])
def test_oracle_reproduces_reference_ipopt_trajectory(geo, fname, max_first2):
    """Per-step parity on identical inputs against the reference's own IPOPT run (recovered from its report PDFs).

    IPOPT stops at tol=1e-5 (HumanoidMpc.py:99) so only the first steps are tight; the heading
    schedule is exact everywhere; the objective of the golden next state is never better than the optimum.
    """
    g = np.load(os.path.join(G, fname))
    X, om, goal, delta = g["X"], g["omega"], g["goal"], float(g["delta"])
    rings = _circle_rings(geo)
    K = X.shape[1] - 1
    s_v = model.foot_parity(K + 4, True)
    dpos, dth, dom = [], [], []
    for k in range(K):
        r = mpc.mpc_step(X[:, k], goal, rings, s_v[k:k + 4], delta=delta)
        assert r["status"] == 0 and max(r["sol"]["kkt"]) < 1e-8
        dpos.append(np.max(np.abs(r["x_next"][[0, 2]] - X[[0, 2], k + 1])))
        dth.append(abs(r["x_next"][4] - X[4, k + 1]))
        dom.append(abs(r["omega"][0] - om[k]))
    dpos = np.array(dpos)
    assert dpos[0] < max_first2 and dpos[1] < max_first2
    assert max(dth) < 1e-6 and max(dom) < 1e-6
    assert np.median(dpos) < 5e-4 and dpos.max() < 5e-3      # IPOPT tol=1e-5 noise, documented in SURVEY §0


def test_lip_constants():
    conf = model.default_conf()
    A, B = model.lip_matrices(conf)
    assert conf["BETA"] == 3.132091952673165
    np.testing.assert_allclose(A[:2, :2], [[1.8929757753645815, 0.5131658337192296], [5.034156828785644, 1.8929757753645815]], rtol=1e-15)
    np.testing.assert_allclose(B[:2, 0], [-0.8929757753645815, -5.034156828785644], rtol=1e-15)


def test_closed_loop_reaches_goal(geo):
    X, U = mpc.run_simulation((6, -3), _circle_rings(geo), (0, 0, 3, 0, 0), 3, 300, 0.4)
    assert X.shape[0] == 5 and U.shape[0] == 3 and X.shape[1] == U.shape[1] + 1
    assert 80 <= U.shape[1] <= 92          # reference: 86 steps
    assert np.hypot(X[0, -1] - 6, X[2, -1] + 3) < 0.2


def test_pspace_oracle_equals_footstep_oracle(geo):
    """The position-space restatement (checker for long horizons) is the same QP as the reference-shaped one."""
    from oracle import qp_pspace
    conf = model.default_conf()
    rings = _circle_rings(geo)
    g = np.load(os.path.join(G, "circles_delta_traj.npz"))
    s_v = model.foot_parity(200, True)
    for k in range(0, 88, 3):
        for N in (2, 3, 4):
            a = mpc.mpc_step(g["X"][:, k], g["goal"], rings, s_v[k:k + N + 1], N=N, sampling_time=0.4, delta=0.3)
            b = qp_pspace.mpc_step(g["X"][:, k], g["goal"], rings, s_v[k:k + N + 1], N, 0.4, conf, delta=0.3)
            assert a["status"] == b["status"]
            if a["status"] == 0:
                np.testing.assert_allclose(b["U"], a["U"], atol=2e-8)
                np.testing.assert_allclose(b["X"], a["X"], atol=2e-8)
                assert abs(a["obj"] - b["obj"]) < 1e-8 * a["obj"]
    # long horizon: the footstep-space Hessian is singular in fp64, the position-space LDP still certifies
    b = qp_pspace.mpc_step((0, 0, 3, 0, 0), (6, -3), rings, model.foot_parity(41, True), 40, 0.4, conf)
    assert b["status"] == 0 and b["kkt"][0] < 1e-7 and b["kkt"][1] < 1e-8


def test_dbscan_restatement_equals_sklearn(geo, lid):
    """The numpy restatement of DBSCAN (what the device kernel implements) gives sklearn's labels on LiDAR scans."""
    from sklearn.cluster import DBSCAN
    from oracle import range_finder
    rs = np.random.default_rng(3)
    n = 0
    for name in MAPS:
        for rn in ("r15", "r30"):
            for reads in lid[f"{name}/{rn}/readings"][:6]:
                for sigma in (0.0, 0.01):
                    pts = reads[~np.isnan(reads[:, 0])]
                    if len(pts) == 0:
                        continue
                    pts = pts + rs.normal(0, sigma, pts.shape) if sigma else pts
                    ref = DBSCAN(eps=0.3, min_samples=3).fit(pts).labels_
                    assert np.array_equal(range_finder.dbscan_labels(pts), ref)
                    n += 1
    assert n > 40
