"""Mirror of the scenario-synthesis half of the reference's `Utils/obstacles.py` (host side, runs once per map).

Needed to regenerate the reference's seeded random maps bit for bit (`Scenario.CROWDED`, `BASE`, ...): the maps come
out of a rejection loop over Python's global `random` stream (reference `:167-194`), so every accept / reject
decision below has to be the decision the reference takes — same draws in the same order, same predicates
(`is_point_inside_polygon` `:32-37`, `polygons_intersect` `:144-154` with its `ccw` / `segment_intersection` tests
`:67-74`, `point_to_polygon_distance` `:41-63`).  Pinned by `tests/test_abi_cpu.py::test_crowded_map_reproduced`
against polygons produced by the reference itself.  The ray / segment intersection of the same reference file
(`:95-139`) is the K4 kernel, not this module.
"""
import random

import numpy as np
from scipy.spatial import ConvexHull


def set_seed(seed):
    random.seed(seed)


def _turn(o, a, b):
    """z-component of (a - o) x (b - o)."""
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])


def _ring_edges(poly):
    n = len(poly)
    return ((poly[i], poly[(i + 1) % n]) for i in range(n))


def generate_random_convex_polygon(num_points, x_range, y_range):
    """Hull (counter-clockwise vertex list) of `num_points` uniform draws: two `random.uniform` calls per point, x first."""
    cloud = [(random.uniform(*x_range), random.uniform(*y_range)) for _ in range(num_points)]
    return [cloud[v] for v in ConvexHull(cloud).vertices]


def is_point_inside_polygon(point, polygon):
    """True when `point` is on the left of (or on) every edge of a counter-clockwise polygon."""
    return all(_turn(a, b, point) >= 0 for a, b in _ring_edges(polygon))


def point_to_segment_distance(p, v, w):
    ex, ey = w[0] - v[0], w[1] - v[1]
    l2 = ex ** 2 + ey ** 2
    if l2 == 0:
        return np.hypot(p[0] - v[0], p[1] - v[1])
    t = max(0, min(1, ((p[0] - v[0]) * ex + (p[1] - v[1]) * ey) / l2))
    return np.hypot(p[0] - (v[0] + t * ex), p[1] - (v[1] + t * ey))


def point_to_polygon_distance(point, polygon):
    return min(point_to_segment_distance(point, a, b) for a, b in _ring_edges(polygon))


def ccw(p, q, r):
    return (r[1] - p[1]) * (q[0] - p[0]) > (q[1] - p[1]) * (r[0] - p[0])


def segment_intersection(a, b, c, d):
    return ccw(a, c, d) != ccw(b, c, d) and ccw(a, b, c) != ccw(a, b, d)


def segment_intersects_polygon(segment, polygon):
    return [(a, b) for a, b in _ring_edges(polygon) if segment_intersection(segment[0], segment[1], a, b)]


def polygons_intersect(polygon1, polygon2):
    """Edge-pair test of the reference: for every (edge of 1, edge of 2) pair, the edges cross or one of the two
    leading vertices lies inside the other polygon."""
    n1, n2 = len(polygon1), len(polygon2)
    for i in range(n1):
        s1 = (polygon1[i], polygon1[(i + 1) % n1])
        for j in range(n2):
            s2 = (polygon2[j], polygon2[(j + 1) % n2])
            if segment_intersects_polygon(s1, [s2[0], s2[1]]) or is_point_inside_polygon(polygon1[i], polygon2) \
                    or is_point_inside_polygon(polygon2[j], polygon1):
                return True
    return False


def polygon_intersect_with_list_of_polygons(polygon, polygons):
    return any(polygons_intersect(p, polygon) for p in polygons)


def generate_polygons(start, goal, num_obstacles, num_points, x_range, y_range, delta):
    """Rejection loop of the reference (`:167-194`): at most 500 candidates; a candidate is a random convex polygon
    in the unit box around a uniform centre and is rejected if it contains the start or the goal, intersects an
    accepted polygon, or its centre is closer than `delta` to one."""
    accepted = []
    for _ in range(500):
        if len(accepted) >= num_obstacles:
            break
        cx = random.uniform(*x_range)
        cy = random.uniform(*y_range)
        poly = generate_random_convex_polygon(num_points, (cx - 0.5, cx + 0.5), (cy - 0.5, cy + 0.5))
        if is_point_inside_polygon(start, poly) or is_point_inside_polygon(goal, poly):
            continue
        if any(polygons_intersect(poly, q) for q in accepted):
            continue
        if any(point_to_polygon_distance((cx, cy), q) < delta for q in accepted):
            continue
        accepted.append(np.array(poly))
    return accepted


def generate_obstacles(start, goal, num_obstacles=10, num_points=5, x_range=(-10, 10), y_range=(-10, 10), delta=1,
                       ch=True):
    polys = generate_polygons(start, goal, num_obstacles, num_points, x_range, y_range, delta)
    return [ConvexHull(p) for p in polys] if ch else polys
