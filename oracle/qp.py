"""Oracle: the per-step LDCBF-MPC quadratic program and its exact optimum.  TEST INFRASTRUCTURE ONLY.

The reference states the problem symbolically and hands it to CasADi `Opti` + IPOPT
(`/root/reference/HumanoidNavigation/MPC/HumanoidMpc.py:97-100,417`; `casadi` is un-vendored and unpinned,
`requirements.txt:3`, and is not installable offline).  With theta/omega precomputed the problem is a
strictly convex QP; this module restates it in condensed form and solves it exactly.

Rows follow the reference's constraint builders:
* dynamics        `HumanoidMpc.py:226-230`  (eliminated: x_k = Sx_k x0 + Su_k z)
* leg reach       `:183-202,233-236`        k = 0..N-1, (longitudinal, lateral)
* maneuverability `:204-219,238-243`        k = 0..N-1, uses theta_{k+1}, omega_k
* walking velocity`:162-181,245-249`        k = 1..N,   (longitudinal, lateral); only the cos term of the
                                            lateral row carries the foot parity (`:177-178`)
* LDCBF           `:252-261,284-292` and the delta variant `HumanoidMPCVariants/HumanoidMPCCustomLCBF.py:30-31`
                                            k = 0..N, obstacle-major inside k; the k = 0 rows are constant
* cost            `:321-333`                sum_{k=0..N} ||p_k - goal||^2 (k = 0 term constant)
Row order: leg | maneuverability | walking | LDCBF.

`solve_exact` is Lawson-Hanson least-distance programming through `scipy.optimize.nnls`
 and returns a KKT certificate; it is the "tight-tolerance solve of the same QP"
that BASELINE.md §4 names as the parity target.
"""
import math

import numpy as np
from scipy.optimize import nnls

from .model import FOOT_LATERAL_OFFSET, condensing

INF = float("inf")


def assemble_condensed(x0, theta, omega, foot, c, eta, goal, conf, delta=0.0):
    """Condensed QP  min 1/2 z'Pz + q'z + const  s.t.  lo <= A z <= hi,  z = (u_0..u_{N-1}) in R^{2N}.

    x0[4] (p_x, v_x, p_y, v_y); theta[N+1]; omega[N]; foot[N+1] (+1/-1, the s_v window of
    `HumanoidMpc.py:403`); c[n_obs,2], eta[n_obs,2]; goal[2].
    Returns dict(P, q, const, A, lo, hi, Sx, Su, kinds) — `kinds[i]` = (type, k, sub) of row i.
    """
    x0 = np.asarray(x0, dtype=np.float64)
    N = len(omega)
    n = 2 * N
    Sx, Su = condensing(N, conf)
    rows, lo, hi, kinds = [], [], [], []

    def add(r, k_state, lo_v, hi_v, kind, k_prev=None):
        """row r (length 4) applied to x_{k_state} (minus x_{k_prev} if given)."""
        a = r @ Su[k_state]
        b = r @ Sx[k_state] @ x0
        if k_prev is not None:
            a = a - r @ Su[k_prev]
            b = b - r @ Sx[k_prev] @ x0
        rows.append(a)
        lo.append(lo_v - b)
        hi.append(hi_v - b)
        kinds.append(kind)

    # leg reachability, k = 0..N-1
    for k in range(N):
        ct, st = math.cos(theta[k]), math.sin(theta[k])
        add(np.array([ct, 0.0, st, 0.0]), k + 1, conf["L_MIN_X"], conf["L_MAX_X"], ("leg", k, 0), k_prev=k)
        off = foot[k] * FOOT_LATERAL_OFFSET
        add(np.array([-st, 0.0, ct, 0.0]), k + 1, conf["L_MIN_Y"] - off, conf["L_MAX_Y"] - off,
            ("leg", k, 1), k_prev=k)
    # maneuverability, k = 0..N-1 on x_{k+1}
    for k in range(N):
        ct, st = math.cos(theta[k + 1]), math.sin(theta[k + 1])
        bound = conf["V_MAX"][0] - (conf["ALPHA"] / np.pi) * abs(omega[k])
        add(np.array([0.0, ct, 0.0, st]), k + 1, -INF, bound, ("man", k, 0))
    # walking velocities, k = 1..N on x_k
    for k in range(1, N + 1):
        ct, st = math.cos(theta[k]), math.sin(theta[k])
        add(np.array([0.0, ct, 0.0, st]), k, conf["V_MIN"][0], conf["V_MAX"][0], ("walk", k, 0))
        add(np.array([0.0, -st, 0.0, ct * foot[k]]), k, conf["V_MIN"][1], conf["V_MAX"][1], ("walk", k, 1))
    # LDCBF, k = 0..N, obstacles inside k
    for k in range(N + 1):
        for o in range(len(c)):
            r = np.array([eta[o][0], 0.0, eta[o][1], 0.0])
            add(r, k, float(eta[o] @ c[o]) + delta, INF, ("ldcbf", k, o))

    A = np.array(rows).reshape(-1, n)
    lo = np.array(lo)
    hi = np.array(hi)

    P = np.zeros((n, n))
    q = np.zeros(n)
    const = 0.0
    for k in range(N + 1):
        for row, g in ((0, goal[0]), (2, goal[1])):
            a = Su[k][row]
            b = Sx[k][row] @ x0 - g
            P += 2.0 * np.outer(a, a)
            q += 2.0 * a * b
            const += b * b
    return dict(P=P, q=q, const=const, A=A, lo=lo, hi=hi, Sx=Sx, Su=Su, kinds=kinds)


def one_sided(A, lo, hi, tol_zero_row=0.0):
    """Split two-sided rows into G z <= h.  Constant (all-zero) rows are dropped after checking their bound.

    Returns (G, h, src, constant_violation) with src[i] = (row index, +1 upper / -1 lower).
    """
    G, h, src = [], [], []
    const_violation = 0.0
    for i in range(A.shape[0]):
        if np.max(np.abs(A[i])) <= tol_zero_row:
            const_violation = max(const_violation, lo[i] - 0.0, 0.0 - hi[i])
            continue
        if hi[i] < INF:
            G.append(A[i]); h.append(hi[i]); src.append((i, +1))
        if lo[i] > -INF:
            G.append(-A[i]); h.append(-lo[i]); src.append((i, -1))
    n = A.shape[1]
    return np.array(G).reshape(-1, n), np.array(h), src, const_violation


def solve_exact(qp, const_row_tol=1e-6):
    """Exact optimum of the strictly convex QP by least-distance programming + NNLS.

    Returns dict(status, z, obj, lam (per one-sided row), kkt (stationarity, primal, complementarity), active).
    status: 0 solved, 2 primal infeasible (the reference's IPOPT raises and the loop breaks,
    `HumanoidMpc.py:419-429`).  A constant row (the k = 0 LDCBF rows) makes the problem infeasible only
    when it is violated by more than `const_row_tol` = 1e-6, the LDCBF tolerance of BASELINE.json
    (IPOPT itself accepts 1e-5, `HumanoidMpc.py:99`).
    """
    P, q, const = qp["P"], qp["q"], qp["const"]
    G, h, src, cviol = one_sided(qp["A"], qp["lo"], qp["hi"])
    n = P.shape[0]
    if cviol > const_row_tol:
        return dict(status=2, z=np.full(n, np.nan), obj=np.nan, lam=None, kkt=None, active=[])
    L = np.linalg.cholesky(P)
    Linv_q = np.linalg.solve(L, q)
    if G.shape[0] == 0:
        z = -np.linalg.solve(L.T, Linv_q)
        lam = np.zeros(0)
    else:
        # y = L'z + L^-1 q ; min 1/2||y||^2 s.t. E y >= f
        E = -np.linalg.solve(L, G.T).T
        f = -(h + G @ np.linalg.solve(P, q))
        M = np.vstack([E.T, f[None, :]])
        e = np.zeros(n + 1)
        e[n] = 1.0
        u, rnorm = nnls(M, e, maxiter=50 * M.shape[1])
        r = M @ u - e
        if abs(r[n]) < 1e-12 or rnorm < 1e-10:
            return dict(status=2, z=np.full(n, np.nan), obj=np.nan, lam=None, kkt=None, active=[])
        y = -r[:n] / r[n]
        z = np.linalg.solve(L.T, y - Linv_q)
        lam = u / (-r[n])
    obj = 0.5 * z @ P @ z + q @ z + const
    if G.shape[0]:
        res = G @ z - h
        kkt = (float(np.max(np.abs(P @ z + q + G.T @ lam))), float(max(0.0, np.max(res))),
               float(np.max(np.abs(lam * res))))
        active = [src[i] for i in range(len(lam)) if lam[i] > 0]
    else:
        kkt = (float(np.max(np.abs(P @ z + q))), 0.0, 0.0)
        active = []
    return dict(status=0, z=z, obj=float(obj), lam=lam, kkt=kkt, active=active, src=src)


def predicted_states(qp, x0, z):
    """X[N+1,4] predicted by the condensed solution (x_k = Sx_k x0 + Su_k z)."""
    x0 = np.asarray(x0, dtype=np.float64)
    return np.array([Sx @ x0 + Su @ z for Sx, Su in zip(qp["Sx"], qp["Su"])])
