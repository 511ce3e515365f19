"""Numpy prototype of the long-horizon solver (config 5): Goldfarb-Idnani dual active set in CoM-position space with
an orthogonal factorisation (J = Q, R) updated by a Householder reflection on add and Givens rotations on drop.
Development aid for csrc/mpc_long.cu; compared with oracle/qp_pspace.py.  Not part of the product path."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "humanoid-navigation-using-mpc-ldcbf_b200")
from oracle import qp_pspace, model          # noqa: E402


def gi_solve(G, h, g, tau=1e-9, tol=1e-10, max_iter=2000):
    """min 1/2||w-g||^2 s.t. G w <= h.  Returns (status, w, iters)."""
    n = len(g)
    Nn = -G
    b = -h
    norms = np.linalg.norm(Nn, axis=1)
    x = g.copy()
    J = np.eye(n)
    R = np.zeros((n, n))
    A, u = [], []
    it = 0
    while True:
        s = (Nn @ x - b) / norms
        s[A] = np.inf
        p = int(np.argmin(s))
        if s[p] > -tol:
            return 0, x, it
        npv = Nn[p]
        up = 0.0
        while True:
            it += 1
            if it > max_iter:
                return 1, x, it
            q = len(A)
            d = J.T @ npv
            d2n = np.linalg.norm(d[q:])
            r = np.linalg.solve(np.triu(R[:q, :q]), d[:q]) if q else np.zeros(0)
            t1, l = np.inf, -1
            for j in range(q):
                if r[j] > 1e-13 * norms[p]:
                    v = u[j] / r[j]
                    if v < t1:
                        t1, l = v, j
            sp = npv @ x - b[p]
            t2 = -sp / (d2n * d2n) if d2n > tau * norms[p] else np.inf
            t = min(t1, t2)
            if not np.isfinite(t):
                return 2, x, it
            if np.isfinite(t2):
                z = J[:, q:] @ d[q:]
                x = x + t * z
            u = [uj - t * rj for uj, rj in zip(u, r)]
            up += t
            if t == t2:
                # add p: Householder on d[q:]
                v = d[q:].copy()
                alpha = -math.copysign(d2n, v[0]) if v[0] != 0 else -d2n
                v[0] -= alpha
                vn = v @ v
                if vn > 0:
                    Jv = J[:, q:] @ v
                    J[:, q:] -= np.outer(Jv, 2.0 * v / vn)
                R[:q, q] = d[:q]
                R[q, q] = alpha
                A.append(p)
                u.append(up)
                break
            # drop l
            for i in range(l, q - 1):
                R[:, i] = R[:, i + 1]
            R[:, q - 1] = 0.0
            for i in range(l, q - 1):
                a_, b_ = R[i, i], R[i + 1, i]
                rr = math.hypot(a_, b_)
                if rr == 0.0:
                    continue
                c_, s_ = a_ / rr, b_ / rr
                Ri, Ri1 = R[i, :].copy(), R[i + 1, :].copy()
                R[i, :] = c_ * Ri + s_ * Ri1
                R[i + 1, :] = -s_ * Ri + c_ * Ri1
                Ji, Ji1 = J[:, i].copy(), J[:, i + 1].copy()
                J[:, i] = c_ * Ji + s_ * Ji1
                J[:, i + 1] = -s_ * Ji + c_ * Ji1
            A.pop(l)
            u.pop(l)


def config5(B, n_obs, seed=0):
    rng = np.random.default_rng(seed)
    side = int(math.ceil(math.sqrt(n_obs)))
    th = np.arange(8) * (2 * np.pi / 8)
    out = []
    for _ in range(B):
        cells = rng.permutation(side * side)[:n_obs]
        rings = []
        pitch = 8.0 / side
        for c in cells:
            cx = 1 + (c % side + 0.5) * pitch + rng.uniform(-0.15, 0.15) * pitch
            cy = 1 + (c // side + 0.5) * pitch + rng.uniform(-0.15, 0.15) * pitch
            rings.append(np.column_stack((cx + 0.3 * np.cos(th), cy + 0.3 * np.sin(th))))
        while True:
            p = rng.uniform(0, 10, 2)
            if all(np.hypot(*(p - r.mean(0))) > 0.45 for r in rings):
                break
        out.append(dict(state=np.array([p[0], 0.0, p[1], 0.0, rng.uniform(-np.pi, np.pi)]), rings=rings,
                        right=bool(rng.random() < 0.5)))
    return out


if __name__ == "__main__":
    conf = model.default_conf()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    n_obs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    goal = np.array([10.0, 10.0])
    from oracle.halfplane import half_planes
    bad = 0
    for i, sc in enumerate(config5(B, n_obs)):
        foot = model.foot_parity(N + 1, sc["right"])
        ref = qp_pspace.mpc_step(sc["state"], goal, sc["rings"], foot, N, 0.4, conf)
        x0, th0 = sc["state"][:4], sc["state"][4]
        c, eta = half_planes(np.array([x0[0], x0[2]]), sc["rings"])
        theta, omega = model.heading_schedule(x0, th0, goal, N, 0.4, conf)
        q = qp_pspace.assemble(x0, theta, omega, foot, c, eta, goal, conf)
        t0 = time.time()
        st, w, it = gi_solve(q["G"], q["h"], q["g"])
        dt = time.time() - t0
        if ref["status"] == 0 and st == 0:
            P = ref["X"][1:, [0, 2]].reshape(-1)
            err = np.max(np.abs(P - w))
        else:
            err = float("nan")
        flag = "" if (st == ref["status"] and (st != 0 or err < 1e-7)) else "  <<<<"
        bad += bool(flag)
        print(f"{i:3d} ref {ref['status']} gi {st} iters {it:4d} rows {len(q['h'])} err {err:.2e} {dt*1e3:.0f} ms{flag}")
    print("mismatches", bad, "of", B)
