"""Shared helpers for the parity tests (CPU side)."""
import os

import numpy as np

from oracle import model, mpc

G = os.path.join(os.path.dirname(__file__), "golden")


def load_geo():
    return np.load(os.path.join(G, "geometry_golden.npz"))


def map_rings(geo, name):
    return [geo[f"{name}/obs{o}/points"][geo[f"{name}/obs{o}/vertices"]] for o in range(int(geo[f"{name}/n_obs"]))]


def map_points(geo, name):
    return [geo[f"{name}/obs{o}/points"] for o in range(int(geo[f"{name}/n_obs"]))]


def golden_step_inputs():
    """Per-step inputs taken from the reference's own trajectories (both circle runs)."""
    geo = load_geo()
    rings = map_rings(geo, "circles")
    states, goals, foots, deltas = [], [], [], []
    for f in ("circles_traj.npz", "circles_delta_traj.npz"):
        g = np.load(os.path.join(G, f))
        X = g["X"]
        s_v = model.foot_parity(X.shape[1] + 4, True)
        for k in range(X.shape[1] - 1):
            states.append(X[:, k]); goals.append(g["goal"]); foots.append(s_v[k:k + 4]); deltas.append(float(g["delta"]))
    return rings, np.array(states), np.array(goals), np.array(foots, dtype=np.int8), np.array(deltas)


def oracle_steps(states, goals, foots, rings_list, deltas, N=3, sampling_time=0.4, conf=None):
    out = []
    for s, g, f, r, d in zip(states, goals, foots, rings_list, deltas):
        out.append(mpc.mpc_step(s, g, r, [int(v) for v in f], N=N, sampling_time=sampling_time, conf=conf, delta=float(d)))
    return out


def check_constraints(U, X, theta, omega, foot, c_eta, nobs, delta, conf, tol=1e-6):
    """Size-independent property: every row of the reference's constraint set holds for a returned solution.

    Returns the worst violation over (dynamics, leg reach, maneuverability, walking velocity, LDCBF)."""
    A, Bm = model.lip_matrices(conf)
    B, N = U.shape[0], U.shape[1]
    worst = 0.0
    for k in range(N):
        pred = X[:, k] @ A.T + U[:, k] @ Bm.T
        worst = max(worst, np.abs(pred - X[:, k + 1]).max())
        ct, st = np.cos(theta[:, k]), np.sin(theta[:, k])
        dp = X[:, k + 1][:, [0, 2]] - X[:, k][:, [0, 2]]
        lg = ct * dp[:, 0] + st * dp[:, 1]
        lt = -st * dp[:, 0] + ct * dp[:, 1] + foot[:, k] * model.FOOT_LATERAL_OFFSET
        worst = max(worst, (lg - conf["L_MAX_X"]).max(), (conf["L_MIN_X"] - lg).max(),
                    (lt - conf["L_MAX_Y"]).max(), (conf["L_MIN_Y"] - lt).max())
        c1, s1 = np.cos(theta[:, k + 1]), np.sin(theta[:, k + 1])
        v = X[:, k + 1][:, [1, 3]]
        man = c1 * v[:, 0] + s1 * v[:, 1] - (conf["V_MAX"][0] - conf["ALPHA"] / np.pi * np.abs(omega[:, k]))
        wl = c1 * v[:, 0] + s1 * v[:, 1]
        wt = -s1 * v[:, 0] + c1 * foot[:, k + 1] * v[:, 1]
        worst = max(worst, man.max(), (wl - conf["V_MAX"][0]).max(), (conf["V_MIN"][0] - wl).max(),
                    (wt - conf["V_MAX"][1]).max(), (conf["V_MIN"][1] - wt).max())
    ldcbf_worst = 0.0
    for k in range(N + 1):
        p = X[:, k][:, [0, 2]]
        for o in range(c_eta.shape[1]):
            m = nobs > o
            h = (c_eta[:, o, 2] * (p[:, 0] - c_eta[:, o, 0]) + c_eta[:, o, 3] * (p[:, 1] - c_eta[:, o, 1])) - delta
            if m.any():
                ldcbf_worst = max(ldcbf_worst, (-h[m]).max())
    return worst, ldcbf_worst
