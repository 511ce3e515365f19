"""Oracle for long horizons: the same QP assembled directly in CoM-position space.  TEST INFRASTRUCTURE ONLY.

`oracle/qp.py` condenses onto the footsteps exactly as the reference poses the problem; that Hessian has condition
number 1.8e13 at N = 10 and is numerically singular for N >= 12, so it cannot serve as
the checker for the scaling sweep (N = 10..40).  This module writes the identical problem in the unknowns
w = (p_1..p_N) using two exact consequences of the LIP step x_{k+1} = A x_k + B u_k
(`/root/reference/HumanoidNavigation/MPC/HumanoidMpc.py:34-48,335-343`):

    v_{k+1} = -v_k + g (p_{k+1} - p_k),            g = beta sinh(beta T) / (cosh(beta T) - 1)
    u_k     = (p_{k+1} - cosh(beta T) p_k - sinh(beta T)/beta v_k) / (1 - cosh(beta T))

so the cost `sum_k ||p_k - goal||^2` (`:321-333`) has Hessian 2I and the problem is a least-distance programme,
solved exactly by Lawson-Hanson NNLS with a KKT certificate.  It is pinned against `oracle/qp.py` (and therefore
against the reference's goldens) at N = 3 in `tests/test_oracle_golden.py::test_pspace_oracle_equals_footstep_oracle`.
Rows: same builders and order as `oracle/qp.py` (leg | maneuverability | walking | LDCBF k = 1..N); the constant
k = 0 LDCBF rows are checked separately.
"""
import math

import numpy as np
from scipy.optimize import nnls

from .model import FOOT_LATERAL_OFFSET, heading_schedule
from .halfplane import half_planes

INF = float("inf")


def lip_scalars(conf):
    beta, T = conf["BETA"], conf["DELTA_T"]
    ch, sh = math.cosh(beta * T), math.sinh(beta * T)
    return ch, sh / beta, beta * sh / (ch - 1.0)


def assemble(x0, theta, omega, foot, c, eta, goal, conf, delta=0.0):
    """Rows G w <= h (one-sided) in w = (p_1x, p_1y, ..., p_Nx, p_Ny); returns dict(G, h, g, const_violation)."""
    N = len(omega)
    n = 2 * N
    _, _, gt = lip_scalars(conf)
    p0 = np.array([x0[0], x0[2]], dtype=np.float64)
    v0 = np.array([x0[1], x0[3]], dtype=np.float64)
    G, h = [], []

    def pos_row(k, r):          # r . p_k as (coefficients on w, constant)
        a = np.zeros(n)
        if k == 0:
            return a, float(r @ p0)
        a[2 * (k - 1):2 * k] = r
        return a, 0.0

    def vel_row(k, r):          # r . v_k
        a = np.zeros(n)
        const = (-1.0) ** k * float(r @ v0)
        # v_k = (-1)^k v_0 + g sum_{j<k} (-1)^{k-1-j} (p_{j+1} - p_j)
        for j in range(k):
            sgn = (-1.0) ** (k - 1 - j)
            ap, cp = pos_row(j + 1, r)
            am, cm = pos_row(j, r)
            a += gt * sgn * (ap - am)
            const += gt * sgn * (cp - cm)
        return a, const

    def two_sided(a, const, lo, hi):
        if hi < INF:
            G.append(a); h.append(hi - const)
        if lo > -INF:
            G.append(-a); h.append(-(lo - const))

    for k in range(N):                                        # leg reachability (HumanoidMpc.py:183-202,233-236)
        ct, st = math.cos(theta[k]), math.sin(theta[k])
        for r, lo, hi in ((np.array([ct, st]), conf["L_MIN_X"], conf["L_MAX_X"]),
                          (np.array([-st, ct]), conf["L_MIN_Y"] - foot[k] * FOOT_LATERAL_OFFSET,
                           conf["L_MAX_Y"] - foot[k] * FOOT_LATERAL_OFFSET)):
            a1, c1 = pos_row(k + 1, r)
            a0, c0 = pos_row(k, r)
            two_sided(a1 - a0, c1 - c0, lo, hi)
    for k in range(N):                                        # maneuverability (:204-219,238-243)
        ct, st = math.cos(theta[k + 1]), math.sin(theta[k + 1])
        a, cst = vel_row(k + 1, np.array([ct, st]))
        two_sided(a, cst, -INF, conf["V_MAX"][0] - (conf["ALPHA"] / np.pi) * abs(omega[k]))
    for k in range(1, N + 1):                                 # walking velocities (:162-181,245-249)
        ct, st = math.cos(theta[k]), math.sin(theta[k])
        a, cst = vel_row(k, np.array([ct, st]))
        two_sided(a, cst, conf["V_MIN"][0], conf["V_MAX"][0])
        a, cst = vel_row(k, np.array([-st, ct * foot[k]]))
        two_sided(a, cst, conf["V_MIN"][1], conf["V_MAX"][1])
    cviol = 0.0
    for k in range(N + 1):                                    # LDCBF (:252-294, delta variant)
        for o in range(len(c)):
            a, cst = pos_row(k, np.asarray(eta[o], dtype=np.float64))
            lo = float(eta[o] @ c[o]) + delta
            if k == 0:
                cviol = max(cviol, lo - cst)
            else:
                two_sided(a, cst, lo, INF)
    return dict(G=np.array(G).reshape(-1, n), h=np.array(h), g=np.tile(np.asarray(goal, dtype=np.float64), N),
                const_violation=cviol, p0=p0, v0=v0)


def solve_ldp(G, h, g):
    """min ||w - g||^2 s.t. G w <= h  by Lawson-Hanson LDP.  Returns (status, w, lam, kkt)."""
    n = len(g)
    if G.shape[0] == 0:
        return 0, g.copy(), np.zeros(0), (0.0, 0.0, 0.0)
    # y = w - g:  min ||y||^2 s.t. (-G) y >= G g - h
    E, f = -G, G @ g - h
    M = np.vstack([E.T, f[None, :]])
    e = np.zeros(n + 1)
    e[n] = 1.0
    u, rnorm = nnls(M, e, maxiter=200 * M.shape[1])
    r = M @ u - e
    if abs(r[n]) < 1e-12 or rnorm < 1e-10:
        return 2, np.full(n, np.nan), None, None
    y = -r[:n] / r[n]
    w = g + y
    lam = 2.0 * u / (-r[n])                       # multipliers of G w <= h for the cost ||w - g||^2
    res = G @ w - h
    kkt = (float(np.max(np.abs(2 * (w - g) + G.T @ lam))), float(max(0.0, res.max())), float(np.max(np.abs(lam * res))))
    return 0, w, lam, kkt


def mpc_step(state, goal, obstacles, foot, N, sampling_time, conf, delta=0.0, const_row_tol=1e-6):
    """One MPC step through the position-space formulation.  Same outputs as `oracle.mpc.mpc_step`."""
    state = np.asarray(state, dtype=np.float64)
    x0, th0 = state[:4], state[4]
    c, eta = half_planes(np.array([x0[0], x0[2]]), obstacles)
    theta, omega = heading_schedule(x0, th0, goal, N, sampling_time, conf)
    out = dict(theta=theta, omega=omega, c=c, eta=eta)
    if not np.all(np.isfinite(eta)):
        out.update(status=3)
        return out
    q = assemble(x0, theta, omega, foot, c, eta, goal, conf, delta)
    if q["const_violation"] > const_row_tol:
        out.update(status=2)
        return out
    status, w, lam, kkt = solve_ldp(q["G"], q["h"], q["g"])
    out.update(status=status, kkt=kkt)
    if status != 0:
        return out
    ch, sob, gt = lip_scalars(conf)
    P = np.vstack([q["p0"], w.reshape(N, 2)])
    V = [q["v0"]]
    for k in range(N):
        V.append(-V[-1] + gt * (P[k + 1] - P[k]))
    V = np.array(V)
    U = np.array([(P[k + 1] - ch * P[k] - sob * V[k]) / (1.0 - ch) for k in range(N)])
    X = np.column_stack((P[:, 0], V[:, 0], P[:, 1], V[:, 1]))
    obj = float(((P - np.asarray(goal)) ** 2).sum())
    out.update(U=U, X=X, obj=obj, x_next=np.concatenate([X[1], [theta[1]]]))
    return out
