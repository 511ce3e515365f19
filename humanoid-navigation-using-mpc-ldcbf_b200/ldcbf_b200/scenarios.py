"""Synthetic scenario generators for the benchmark configs of SURVEY.md §8d (host side, numpy only).

Pure numpy, seeded; used by `bench.py` (inputs only — the generated arrays are handed to the CUDA path) and by
the parity tests so both sides see identical inputs.  The obstacle shapes restate
`/root/reference/HumanoidNavigation/Utils/ObstaclesUtils.py:21-36` (`generate_circle_like_polygon`: num_points on
`linspace(0, 2*pi, num_points)`, i.e. the last point duplicates the first, leaving num_points-1 hull vertices) and
the CIRCLE_OBSTACLES map of `report_simulations/Scenario.py:202-206`.
"""
import math

import numpy as np

CIRCLES = ((10, 0.5, (5.5, -1.2)), (20, 1.0, (4.0, 2.0)), (25, 1.2, (1.7, 0.0)))


def circle_ring(num_points, radius, center):
    """Hull vertices (counter-clockwise) of the reference's circle-like polygon: num_points-1 distinct vertices."""
    th = np.linspace(0, 2 * np.pi, num_points)[:-1]
    return np.column_stack((center[0] + radius * np.cos(th), center[1] + radius * np.sin(th)))


def circle_rings(jitter=None):
    rings = []
    for i, (n, r, c) in enumerate(CIRCLES):
        cc = (c[0] + (jitter[i][0] if jitter is not None else 0.0), c[1] + (jitter[i][1] if jitter is not None else 0.0))
        rings.append(circle_ring(n, r, cc))
    return rings


def pack_rings(list_of_ring_lists, max_obs=None, max_verts=None):
    """-> verts[B,max_obs,max_verts,2] (zero padded), nverts[B,max_obs] int32, nobs[B] int32."""
    B = len(list_of_ring_lists)
    max_obs = max_obs or max(1, max(len(r) for r in list_of_ring_lists))
    max_verts = max_verts or max(1, max((len(p) for r in list_of_ring_lists for p in r), default=1))
    verts = np.zeros((B, max_obs, max_verts, 2))
    nverts = np.zeros((B, max_obs), dtype=np.int32)
    nobs = np.zeros(B, dtype=np.int32)
    for b, rings in enumerate(list_of_ring_lists):
        nobs[b] = len(rings)
        for o, ring in enumerate(rings):
            nverts[b, o] = len(ring)
            verts[b, o, :len(ring)] = ring
    return verts, nverts, nobs


def _dist_to_ring(p, ring):
    a = ring
    b = np.roll(ring, -1, axis=0)
    ab = b - a
    t = np.clip(((p - a) * ab).sum(1) / (ab * ab).sum(1), 0, 1)
    c = a + t[:, None] * ab
    return float(np.min(np.hypot(*(c - p).T)))


def _inside_convex(p, ring):
    a = ring
    b = np.roll(ring, -1, axis=0)
    cr = (b[:, 0] - a[:, 0]) * (p[1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (p[0] - a[:, 0])
    return bool(np.all(cr >= 0))


def config2(B=4096, seed=0):
    """Batched basic simulation.

    start p ~ U([-1,1]x[2,4]), v = 0, theta0 ~ U(-pi,pi), first foot ~ Bernoulli(1/2); goal ~ U([5,7]x[-4,-2]);
    the three config-1 circles with centres jittered U(-0.25,0.25)^2; a sample is redrawn while start or goal is
    inside / within 0.05 of a polygon.  Returns dict of arrays (state[B,5], goal[B,2], right_first[B] bool,
    verts, nverts, nobs).
    """
    rng = np.random.default_rng(seed)
    state = np.zeros((B, 5))
    goal = np.zeros((B, 2))
    right = np.zeros(B, dtype=bool)
    rings_all = []
    for b in range(B):
        while True:
            p = rng.uniform((-1, 2), (1, 4))
            g = rng.uniform((5, -4), (7, -2))
            th = rng.uniform(-math.pi, math.pi)
            rf = rng.random() < 0.5
            jit = rng.uniform(-0.25, 0.25, size=(3, 2))
            rings = circle_rings(jit)
            ok = all(not _inside_convex(q, r) and _dist_to_ring(q, r) > 0.05 for q in (p, g) for r in rings)
            if ok:
                break
        state[b] = (p[0], 0.0, p[1], 0.0, th)
        goal[b] = g
        right[b] = rf
        rings_all.append(rings)
    verts, nverts, nobs = pack_rings(rings_all, 3, 24)
    return dict(state=state, goal=goal, right_first=right, verts=verts, nverts=nverts, nobs=nobs, rings=rings_all)


def config2_sharded(B_total, lo, hi, seed=0, block=4096):
    """Scenarios [lo, hi) of the multi-GPU config-2 batch: the concatenation of config2(block, seed + i), i = 0, 1, ...
    (block i holds scenarios [i*block, (i+1)*block)), so a rank only draws the blocks its shard touches and the first
    `block` scenarios are exactly config2(block, seed)."""
    parts = []
    for i in range(lo // block, (min(hi, B_total) - 1) // block + 1):
        n = min(block, B_total - i * block)
        sc = config2(n, seed=seed + i)
        a, b = max(lo, i * block) - i * block, min(hi, i * block + n) - i * block
        parts.append({k: v[a:b] for k, v in sc.items()})
    out = {}
    for k in parts[0]:
        v = [p[k] for p in parts]
        out[k] = sum(v, []) if isinstance(v[0], list) else np.concatenate(v, axis=0)
    return out


def foot_window(right_first, step, N):
    """Parity window s_v[step:step+N+1] of HumanoidMpc.py:104-108,403 for arrays of scenarios -> int8[B,N+1]."""
    right_first = np.asarray(right_first, dtype=bool)
    idx = step + np.arange(N + 1)[None, :]
    even = (idx % 2) == 0
    s = np.where(even == right_first[:, None], 1, -1)
    return s.astype(np.int8)


def config3(B=16384, seed=0, pool=64, n_obstacles=20):
    """Unknown-environment shape (`report_simulations/simulation_1.py:201-231`): start (0,0)
    heading pi/2, goal (4, 3.5), `n_obstacles` convex polygons = hull of 5 uniform points in a 1x1 box around centres
    drawn in (-1,6)^2, rejecting a candidate whose box comes within 0.2 of an accepted box or that contains the start
    or the goal (a simplification of the rejection rules of `Utils/obstacles.py:167-194` with the same obstacle
    density).  `pool` distinct maps are generated and tiled to B scenarios; the LiDAR pose is jittered per scenario.
    Returns dict(state[B,5], goal[B,2], verts[B,20,5,2], nverts, nobs, rings (pool lists), pos[B,2])."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(seed)
    maps = []
    for _ in range(pool):
        rings, centres = [], []
        tries = 0
        while len(rings) < n_obstacles and tries < 400:
            tries += 1
            c = rng.uniform(-1, 6, 2)
            if any(np.max(np.abs(c - q)) < 1.2 for q in centres):
                continue
            pts = c + rng.uniform(-0.5, 0.5, (5, 2))
            ring = pts[ConvexHull(pts).vertices]
            if any(_inside_convex(np.array(q), ring) or _dist_to_ring(np.array(q), ring) < 0.1 for q in ((0.0, 0.0), (4.0, 3.5))):
                continue
            rings.append(ring)
            centres.append(c)
        maps.append(rings)
    verts_p, nverts_p, nobs_p = pack_rings(maps, n_obstacles, 5)
    idx = np.arange(B) % pool
    pos = np.column_stack((rng.uniform(-0.2, 4.2, B), rng.uniform(-0.2, 3.7, B)))
    state = np.zeros((B, 5))
    state[:, 0], state[:, 2], state[:, 4] = pos[:, 0], pos[:, 1], math.pi / 2
    return dict(state=state, goal=np.tile([4.0, 3.5], (B, 1)), verts=verts_p[idx], nverts=nverts_p[idx], nobs=nobs_p[idx],
                rings=maps, map_index=idx, pos=pos)


def config4(B=8192, seed=0, n_goals=6):
    """RRT* sub-goal variant (`report_simulations/simulation_rrt.py:18-24`): one wall
    `ConvexHull([[2,-3],[2,3],[3,-3],[3,3]])`, start at the origin, goal (5, 0).  RRT* planning is out of scope, so the
    way-points are an input: `n_goals` points routed around either end of the wall with U(-0.3, 0.3) jitter, the last
    one the goal itself.  Returns dict(state[B,5], goals[B,n_goals,2], verts[B,1,4,2], nverts, nobs, right_first)."""
    rng = np.random.default_rng(seed)
    wall = np.array([[2.0, -3.0], [3.0, -3.0], [3.0, 3.0], [2.0, 3.0]])          # hull vertices, counter-clockwise
    side = np.where(rng.random(B) < 0.5, 1.0, -1.0)                               # pass above (+) or below (-) the wall
    base = np.array([[1.0, 1.5], [1.6, 3.0], [2.5, 3.9], [3.6, 3.0], [4.4, 1.5], [5.0, 0.0]])
    if n_goals != len(base):
        t = np.linspace(0, len(base) - 1, n_goals)
        base = np.column_stack([np.interp(t, np.arange(len(base)), base[:, i]) for i in range(2)])
    goals = np.tile(base, (B, 1, 1))
    goals[:, :, 1] *= side[:, None]
    goals[:, :-1] += rng.uniform(-0.3, 0.3, (B, n_goals - 1, 2))
    goals[:, -1] = (5.0, 0.0)
    verts, nverts, nobs = pack_rings([[wall]] * B, 1, 4)
    return dict(state=np.zeros((B, 5)), goals=goals, verts=verts, nverts=nverts, nobs=nobs,
                right_first=np.ones(B, dtype=bool), rings=[wall])


def config5(B=1024, n_obs=8, seed=0, pool=32):
    """Scaling sweep: `n_obs` regular octagons of radius 0.3 on a jittered grid over [1,9]^2
    (the MAIN_PAPER extent of `report_simulations/Scenario.py:220-222`), goal (10,10); one open-loop step per scenario
    from a random obstacle-free position in [0,10]^2 at rest with a random heading and first foot.  `pool` distinct
    maps are generated and tiled to B scenarios.
    Returns dict(state[B,5], goal[B,2], right_first[B], verts[B,n_obs,8,2], nverts, nobs, rings (pool lists), map_index)."""
    rng = np.random.default_rng(seed)
    side = int(math.ceil(math.sqrt(n_obs)))
    pitch = 8.0 / side
    ang = np.arange(8) * (2 * np.pi / 8)
    maps, centres = [], []
    for _ in range(pool):
        cells = rng.permutation(side * side)[:n_obs]
        c = np.column_stack((1 + (cells % side + 0.5) * pitch, 1 + (cells // side + 0.5) * pitch))
        c += rng.uniform(-0.15, 0.15, c.shape) * pitch
        centres.append(c)
        maps.append([np.column_stack((cc[0] + 0.3 * np.cos(ang), cc[1] + 0.3 * np.sin(ang))) for cc in c])
    verts_p, nverts_p, nobs_p = pack_rings(maps, n_obs, 8)
    idx = np.arange(B) % pool
    state = np.zeros((B, 5))
    for b in range(B):
        while True:
            p = rng.uniform(0, 10, 2)
            if np.min(np.hypot(*(centres[idx[b]] - p).T)) > 0.45:
                break
        state[b] = (p[0], 0.0, p[1], 0.0, rng.uniform(-math.pi, math.pi))
    return dict(state=state, goal=np.tile([10.0, 10.0], (B, 1)), right_first=rng.random(B) < 0.5, verts=verts_p[idx],
                nverts=nverts_p[idx], nobs=nobs_p[idx], rings=maps, map_index=idx)
