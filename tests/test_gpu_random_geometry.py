"""Randomised GPU parity of the geometry kernels against the oracle: random convex polygons (3..40 vertices, tiny
and huge edges, duplicated vertices, query points inside / outside / near vertices), both K1 mappings and the LiDAR
caster.  Everything is compared bit for bit."""
import numpy as np
import pytest
import torch
from scipy.spatial import ConvexHull

from oracle import halfplane, lidar

pytestmark = pytest.mark.gpu


def cu(a, dt=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()


def random_rings(rs, n_scen, max_obs, max_verts):
    out = []
    for _ in range(n_scen):
        rings = []
        for _ in range(int(rs.integers(0, max_obs + 1))):
            nv = int(rs.integers(3, max_verts + 1))
            c = rs.uniform(-4, 4, 2)
            scale = 10.0 ** rs.uniform(-2, 0.7)
            ang = np.sort(rs.uniform(0, 2 * np.pi, nv))
            pts = c + scale * np.column_stack((np.cos(ang), np.sin(ang))) * rs.uniform(0.6, 1.0, (nv, 1))
            ring = pts[ConvexHull(pts).vertices]
            if rs.random() < 0.2 and len(ring) < max_verts:          # a duplicated vertex: zero-length edge
                k = int(rs.integers(0, len(ring)))
                ring = np.insert(ring, k, ring[k], axis=0)
            rings.append(ring)
        out.append(rings)
    return out


@pytest.mark.parametrize("B,max_obs,max_verts", [(64, 3, 40), (3000, 5, 12), (20000, 2, 7)])
def test_halfplanes_random_polygons_bit_equal(B, max_obs, max_verts):
    """Covers the split-ring kernel (small batches), the thread-per-ring kernel (large) and max_verts <= 8."""
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios
    rs = np.random.default_rng(B)
    pool = random_rings(rs, 64, max_obs, max_verts)
    idx = rs.integers(0, 64, B)
    rings_all = [pool[i] for i in idx]
    Q = rs.uniform(-5, 5, (B, 2))
    for b in range(0, B, 7):                       # some queries on / next to a vertex or inside
        if rings_all[b]:
            r = rings_all[b][0]
            Q[b] = r[int(rs.integers(0, len(r)))] + rs.choice([0.0, 1e-9, 1e-3]) * rs.normal(size=2) if b % 2 else r.mean(0)
    verts, nverts, nobs = scenarios.pack_rings(rings_all, max_obs, max_verts)
    ce = L.half_planes(cu(Q), cu(verts), cu(nverts, torch.int32), cu(nobs, torch.int32)).cpu().numpy()
    check = range(B) if B <= 3000 else rs.choice(B, 1500, replace=False)
    for b in check:
        c, eta = halfplane.half_planes(Q[b], rings_all[b])
        n = len(rings_all[b])
        assert np.array_equal(ce[b, :n, :2], c), b
        assert np.array_equal(ce[b, :n, 2:], eta, equal_nan=True), b
        assert not ce[b, n:].any()


def test_lidar_random_polygons_bit_exact():
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios
    rs = np.random.default_rng(77)
    B = 96
    rings_all = random_rings(rs, B, 9, 16)
    pos = rs.uniform(-5, 5, (B, 2))
    verts, nverts, nobs = scenarios.pack_rings(rings_all, 9, 17)
    for rng_, R in ((1.5, 360), (6.0, 97)):
        ho, he, xy = L.lidar_cast(cu(pos), cu(verts), cu(nverts, torch.int32), cu(nobs, torch.int32), rng_, R)
        ho, he, xy = ho.cpu().numpy(), he.cpu().numpy(), xy.cpu().numpy()
        for b in range(B):
            o_ho, o_he, o_xy = lidar.cast(pos[b], rings_all[b], rng_, R)
            assert np.array_equal(ho[b], o_ho) and np.array_equal(he[b], o_he), (b, R)
            assert np.array_equal(xy[b], o_xy, equal_nan=True), (b, R)


@pytest.mark.parametrize("max_verts", [12, 24, 40])
def test_halfplanes_wide_loads_equal_narrow_loads(max_verts):
    """The thread-per-ring kernel reads two vertices per 256-bit load when the vertex pairs are 32-byte aligned (even
    max_verts, aligned base) and one vertex per 128-bit load otherwise.  Same map at a 32-byte aligned and at a
    16-byte-only aligned address: bit-identical half-planes (odd and even ring sizes, garbage in the padding), and
    equal to the oracle."""
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios
    B, max_obs = 20000, 3                                     # 480 000 lanes of work: the thread-per-ring kernel
    rs = np.random.default_rng(77 + max_verts)
    pool = random_rings(rs, 64, max_obs, max_verts)
    rings_all = [pool[i] for i in rs.integers(0, 64, B)]
    Q = rs.uniform(-5, 5, (B, 2))
    verts, nverts, nobs = scenarios.pack_rings(rings_all, max_obs, max_verts)
    verts = verts.reshape(B, max_obs, max_verts, 2)
    verts[np.arange(max_verts)[None, None, :] >= nverts[:, :, None]] = 1e300
    nv, no, q = cu(nverts, torch.int32), cu(nobs, torch.int32), cu(Q)
    buf = torch.empty(verts.size + 4, dtype=torch.float64, device="cuda")
    assert buf.data_ptr() % 32 == 0
    wide_in = buf[:verts.size].view(B, max_obs, max_verts, 2)
    wide_in.copy_(cu(verts))
    wide = L.half_planes(q, wide_in, nv, no).clone()
    narrow_in = buf[2:2 + verts.size].view(B, max_obs, max_verts, 2)     # 16 bytes further: no 256-bit loads
    assert narrow_in.data_ptr() % 32 == 16
    narrow_in.copy_(cu(verts))
    narrow = L.half_planes(q, narrow_in, nv, no)
    assert torch.equal(wide.view(torch.int64), narrow.view(torch.int64))
    got = wide.cpu().numpy()
    for b in rs.choice(B, 400, replace=False):
        c, eta = halfplane.half_planes(Q[b], rings_all[b])
        n = len(rings_all[b])
        assert np.array_equal(got[b, :n, :2], c) and np.array_equal(got[b, :n, 2:], eta, equal_nan=True), b
