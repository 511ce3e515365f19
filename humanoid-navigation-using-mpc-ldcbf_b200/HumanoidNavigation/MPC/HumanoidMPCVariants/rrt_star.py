"""RRT* on an occupancy grid with a clearance-weighted edge cost (host side, numpy).

The reference delegates this search to the third-party package `rrtplanner==0.1.2`
(`HumanoidMPCWithRRT.py:6-7,123-127`: `RRTStar(og, n=1500, r_rewire=80, costfn=cost_fn, seed=1).plan(start, goal)`),
which is not vendored and not installable offline, so its random tree cannot be reproduced (parity unpinned, DESIGN.md
§3).  This is an independent implementation of the textbook algorithm (Karaman & Frazzoli 2011) with the reference's
parameters and the reference's edge cost `cost(parent) + costs_matrix[x] * ||parent - x||` (`:113-117`); the map, the
distance transform and the cost matrix it consumes are computed on the GPU (`ldcbf_clearance_grid_f64`).
Sequential and branchy per scenario: it stays on the host by design.
"""
import numpy as np


class RRTStar:
    def __init__(self, og, costs_matrix, n=1500, r_rewire=80.0, seed=1):
        self.og = np.asarray(og) != 0
        self.costs = np.asarray(costs_matrix, dtype=np.float64)
        self.n = int(n)
        self.r = float(r_rewire)
        self.rng = np.random.default_rng(seed)
        self.free = np.argwhere(~self.og)

    def collision_free(self, a, b):
        """True when every cell on the segment a-b (sampled at half-cell spacing) is free."""
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        steps = max(2, int(np.ceil(2.0 * np.hypot(*(b - a)))) + 1)
        t = np.linspace(0.0, 1.0, steps)[:, None]
        cells = np.rint(a + t * (b - a)).astype(int)
        return not self.og[cells[:, 0], cells[:, 1]].any()

    def visible_from(self, points, x):
        """collision_free(p, x) for every row p of `points` at once (same half-cell sampling)."""
        points = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        if len(points) == 0:
            return np.zeros(0, dtype=bool)
        x = np.asarray(x, dtype=np.float64)
        length = np.hypot(points[:, 0] - x[0], points[:, 1] - x[1])
        steps = max(2, int(np.ceil(2.0 * length.max())) + 1)
        t = np.linspace(0.0, 1.0, steps)[None, :, None]
        cells = np.rint(points[:, None, :] + t * (x - points)[:, None, :]).astype(int)
        return ~self.og[cells[..., 0], cells[..., 1]].any(axis=1)

    def shortcut(self, path):
        """Removes way-points while that does not raise the path's cost under the same edge cost (cost matrix at the
        edge's end point times its length) and keeps every edge collision-free.  Not part of the reference's planner;
        it removes the zig-zags a 1500-sample tree leaves where the clearance cost is numerically flat."""
        path = [np.asarray(p) for p in path]
        i = 0
        while i < len(path) - 2:
            edge = lambda a, b: self.costs[b[0], b[1]] * float(np.hypot(*(a - b).astype(np.float64)))
            run = 0.0
            best = None
            for j in range(i + 1, len(path)):
                run += edge(path[j - 1], path[j])
                if j > i + 1 and edge(path[i], path[j]) <= run and self.collision_free(path[i], path[j]):
                    best = j
            if best is not None:
                del path[i + 1:best]
            i += 1
        return path

    def plan(self, start, goal):
        """-> list of grid cells from start to goal (both included), or None when no path was found."""
        start = np.asarray(start, dtype=int)
        goal = np.asarray(goal, dtype=int)
        P = np.zeros((self.n + 2, 2), dtype=int)
        parent = np.full(self.n + 2, -1, dtype=int)
        cost = np.zeros(self.n + 2)
        P[0] = start
        m = 1
        for _ in range(self.n):
            x_rand = self.free[self.rng.integers(len(self.free))]
            d = np.hypot(P[:m, 0] - x_rand[0], P[:m, 1] - x_rand[1])
            near = int(np.argmin(d))
            if d[near] == 0.0:
                continue
            x_new = x_rand if d[near] <= self.r else np.rint(P[near] + (x_rand - P[near]) * (self.r / d[near])).astype(int)
            if self.og[x_new[0], x_new[1]] or not self.collision_free(P[near], x_new):
                continue
            dn = np.hypot(P[:m, 0] - x_new[0], P[:m, 1] - x_new[1])
            near_ids = np.flatnonzero((dn <= self.r) & (dn > 0))
            cand = [int(v) for v in near_ids[self.visible_from(P[near_ids], x_new)]]
            if not cand:
                continue
            c_new = [cost[v] + self.costs[x_new[0], x_new[1]] * dn[v] for v in cand]
            best = int(np.argmin(c_new))
            P[m], parent[m], cost[m] = x_new, cand[best], c_new[best]
            for v in cand:                                               # rewire the neighbourhood through x_new
                if v == cand[best] or v == 0:
                    continue
                c_via = cost[m] + self.costs[P[v, 0], P[v, 1]] * dn[v]
                if c_via < cost[v] and not self._is_ancestor(parent, v, m):
                    delta = c_via - cost[v]
                    parent[v] = m
                    self._shift_subtree(parent, cost, v, delta, m + 1)
            m += 1
        # connect the goal to the cheapest visible vertex
        dg = np.hypot(P[:m, 0] - goal[0], P[:m, 1] - goal[1])
        order = np.argsort(cost[:m] + self.costs[goal[0], goal[1]] * dg)
        for v in order:
            if self.collision_free(P[v], goal):
                path = [goal] if dg[v] > 0 else []
                while v >= 0:
                    path.append(P[v].copy())
                    v = parent[v]
                return [np.asarray(p) for p in reversed(path)]
        return None

    @staticmethod
    def _is_ancestor(parent, v, of):
        while of >= 0:
            if of == v:
                return True
            of = parent[of]
        return False

    @staticmethod
    def _shift_subtree(parent, cost, root, delta, m):
        stack = [root]
        while stack:
            v = stack.pop()
            cost[v] += delta
            stack.extend(int(c) for c in np.flatnonzero(parent[:m] == v))
