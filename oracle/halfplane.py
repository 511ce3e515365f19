"""Oracle: LDCBF half-plane (closest point c and unit normal eta) per obstacle.  TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/HumanoidNavigation/Utils/ObstaclesUtils.py`:
* `:60-109` `get_closest_point_and_normal_vector_from_obs`  -> `closest_point_and_normal`
* `:50-57`  `is_point_inside_polygon` (matplotlib `Path.contains_point`)  -> `point_inside`
and `MPC/HumanoidMpc.py:296-319` (`_get_list_c_and_eta`) -> `half_planes`.

Arithmetic is spelled out in scalar IEEE fp64 operations in the reference's operation order (numpy's
2-element `dot` is fma(a1, b1, a0*b0) on the reference build — `model.dot2`; `np.power(np.linalg.norm(v), 2)`
is `sqrt(dot(v, v))**2`), so the functions here are bit-equal to the reference when given the reference's
edge list (`ConvexHull.simplices` order) — pinned in `tests/test_oracle_golden.py`.  The CUDA path
takes the hull vertices in counter-clockwise order and walks edges (i, i+1 mod V); `cyclic_edges`
builds that edge list so both sides see identical inputs.
"""
import math

import numpy as np

from .model import dot2


def cyclic_edges(n):
    """Edge list (i, i+1 mod n) of a polygon given as an ordered vertex ring."""
    return [(i, (i + 1) % n) for i in range(n)]


def point_inside(x, ring):
    """Crossing-number test of `matplotlib.path.Path(ring).contains_point(x)` (ObstaclesUtils.py:50-57).

    `ring` is `polygon.points[polygon.vertices]` (hull vertices, counter-clockwise, implicitly closed).
    Restates the Haines crossing test matplotlib uses (src/_path.h `point_in_path_impl`); it agrees with
    the plain even-odd rule everywhere except exactly on the boundary, where both are undefined.
    """
    tx, ty = float(x[0]), float(x[1])
    n = len(ring)
    inside = False
    vx0, vy0 = float(ring[n - 1][0]), float(ring[n - 1][1])
    yflag0 = vy0 >= ty
    for i in range(n):
        vx1, vy1 = float(ring[i][0]), float(ring[i][1])
        yflag1 = vy1 >= ty
        if yflag0 != yflag1:
            if ((vy1 - ty) * (vx0 - vx1) >= (vx1 - tx) * (vy0 - vy1)) == yflag1:
                inside = not inside
        yflag0 = yflag1
        vx0, vy0 = vx1, vy1
    return inside


def closest_point_on_edges(x, pts, edges):
    """First strict minimum over edges of the clamped projection of x (ObstaclesUtils.py:71-96).

    Returns (c[2], dist, edge_index).
    """
    px, py = float(x[0]), float(x[1])
    best = None
    min_dist = float("inf")
    best_e = -1
    for e, (ia, ib) in enumerate(edges):
        ax, ay = float(pts[ia][0]), float(pts[ia][1])
        bx, by = float(pts[ib][0]), float(pts[ib][1])
        apx, apy = px - ax, py - ay
        abx, aby = bx - ax, by - ay
        nrm = math.sqrt(dot2(abx, aby, abx, aby))
        den = nrm * nrm
        num = dot2(apx, apy, abx, aby)
        t = num / den if den != 0.0 else float("nan")
        # python's max(0, min(1, t)): NaN (zero-length edge) falls through to 1
        t = max(0.0, min(1.0, t)) if t == t else 1.0
        cx, cy = ax + t * abx, ay + t * aby
        dx, dy = cx - px, cy - py
        dist = math.sqrt(dot2(dx, dy, dx, dy))
        if dist < min_dist:
            best, min_dist, best_e = (cx, cy), dist, e
    return np.array(best), min_dist, best_e


def closest_point_and_normal(x, pts, edges, ring):
    """(c, eta) of ObstaclesUtils.py:60-109 with `unitary_normal_vector=True`.

    eta = (x - c)/||x - c||, negated when x lies inside the hull (`:106-107`).  A CoM exactly on the
    boundary gives 0/0 = NaN, as in the reference (`:104`).
    """
    c, _, _ = closest_point_on_edges(x, pts, edges)
    nx, ny = float(x[0]) - c[0], float(x[1]) - c[1]
    nrm = math.sqrt(dot2(nx, ny, nx, ny))
    with np.errstate(all="ignore"):
        eta = np.array([nx, ny]) / np.float64(nrm)
    if point_inside(x, ring):
        eta = eta * -1
    return c, eta


def half_planes(x, obstacles):
    """`_get_list_c_and_eta` (HumanoidMpc.py:296-319) over obstacles given as ordered vertex rings.

    obstacles: list of (V_o, 2) arrays, hull vertices counter-clockwise.  Returns c[n_obs,2], eta[n_obs,2].
    """
    cs, etas = [], []
    for ring in obstacles:
        ring = np.asarray(ring, dtype=np.float64)
        c, eta = closest_point_and_normal(x, ring, cyclic_edges(len(ring)), ring)
        cs.append(c)
        etas.append(eta)
    if not cs:
        return np.zeros((0, 2)), np.zeros((0, 2))
    return np.array(cs), np.array(etas)
