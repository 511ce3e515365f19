"""Oracle: constants, LIP transition matrices, heading schedule.  TEST INFRASTRUCTURE ONLY.

Follows (reference, read-only, `/root/reference/HumanoidNavigation/`):
* `config.yml:2-17` and the derived keys of `MPC/HumanoidMpc.py:16-22`  -> `default_conf`
* `MPC/HumanoidMpc.py:34-48` (A_l, B_l; math in `Report/chapters/LIP.tex:67-92`) -> `lip_matrices`
* `MPC/HumanoidMpc.py:104-108` (foot parity list s_v)                              -> `foot_parity`
* `MPC/HumanoidMpc.py:137-160` (`_precompute_theta_omega_naive`)                   -> `heading_schedule`
"""
import ctypes
import ctypes.util
import math
from fractions import Fraction

import numpy as np

try:
    _libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    _libm.fma.restype = ctypes.c_double
    _libm.fma.argtypes = (ctypes.c_double, ctypes.c_double, ctypes.c_double)

    def fma(a, b, c):
        """Correctly rounded a*b + c (C99 `fma`)."""
        return _libm.fma(a, b, c)
except (OSError, AttributeError):  # pragma: no cover
    def fma(a, b, c):
        return float(Fraction(a) * Fraction(b) + Fraction(c))


def dot2(a0, a1, b0, b1):
    """numpy's 2-element `dot` as the reference build evaluates it: fma(a1, b1, a0*b0).

    Probed against numpy 2.3.5 / OpenBLAS 0.3.30 in the build container (2000 random pairs, 100 % bit-equal;
    the un-fused a0*b0 + a1*b1 matches only 75 %).  `np.linalg.norm` of a 2-vector is sqrt of this dot.
    """
    return fma(a1, b1, a0 * b0)


def default_conf():
    """The values of `config.yml:2-17` plus the keys derived at import time (`HumanoidMpc.py:20-22`)."""
    conf = {
        "DELTA_T": 0.4, "GRAVITY_CONST": 9.81, "COM_HEIGHT": 1, "ALPHA": 3.6,
        "L_MAX_X": 0.10, "L_MAX_Y": 0.10, "L_MIN_X": -0.1, "L_MIN_Y": -0.1,
        "V_MIN": [-0.1, 0.1], "V_MAX": [0.8, 0.4], "RIGHT_FOOT": 1, "LEFT_FOOT": -1,
    }
    conf["BETA"] = float(np.sqrt(conf["GRAVITY_CONST"] / conf["COM_HEIGHT"]))
    conf["OMEGA_MAX"] = 0.156 * math.pi
    conf["OMEGA_MIN"] = -conf["OMEGA_MAX"]
    return conf


# Hard-coded in the reference, not in config.yml:
FOOT_LATERAL_OFFSET = 0.05   # HumanoidMpc.py:200
STOP_OBJECTIVE = 0.05        # HumanoidMpc.py:392


def lip_matrices(conf):
    """A_l (4x4), B_l (4x2) of HumanoidMpc.py:34-48.  State (p_x, v_x, p_y, v_y), input (f_x, f_y)."""
    beta, T = conf["BETA"], conf["DELTA_T"]
    ch, sh = math.cosh(beta * T), math.sinh(beta * T)
    Ad = np.array([[ch, sh / beta], [sh * beta, ch]])
    Bd = np.array([1 - ch, -beta * sh])
    A = np.zeros((4, 4))
    A[:2, :2] = Ad
    A[2:, 2:] = Ad
    B = np.zeros((4, 2))
    B[:2, 0] = Bd
    B[2:, 1] = Bd
    return A, B


def foot_parity(num, start_with_right_foot=True, conf=None):
    """s_v of HumanoidMpc.py:104-108: +1 (right) on even indices when starting with the right foot."""
    conf = conf or default_conf()
    r, l = conf["RIGHT_FOOT"], conf["LEFT_FOOT"]
    return [r if i % 2 == (0 if start_with_right_foot else 1) else l for i in range(num)]


def heading_schedule(x0, theta0, goal, N, sampling_time, conf):
    """theta[N+1], omega[N] of `_precompute_theta_omega_naive` (HumanoidMpc.py:137-160).

    Quirks kept: the target angle uses the *current* CoM for every k, the difference is not wrapped,
    theta advances by omega * sampling_time.
    """
    theta = [float(theta0)]
    omega = []
    for _ in range(N):
        target = math.atan2(goal[1] - x0[2], goal[0] - x0[0]) - theta[-1]
        w = min(max(target, conf["OMEGA_MIN"]), conf["OMEGA_MAX"])
        omega.append(w)
        theta.append(theta[-1] + w * sampling_time)
    return np.array(theta), np.array(omega)


def condensing(N, conf):
    """Sx[k] (4x4), Su[k] (4x2N) with x_k = Sx[k] x0 + Su[k] z, z = (u_0..u_{N-1})."""
    A, B = lip_matrices(conf)
    Sx = [np.eye(4)]
    Su = [np.zeros((4, 2 * N))]
    for k in range(N):
        nSu = A @ Su[-1]
        nSu[:, 2 * k:2 * k + 2] += B
        Sx.append(A @ Sx[-1])
        Su.append(nSu)
    return Sx, Su
