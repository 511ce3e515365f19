"""Oracle: one LDCBF-MPC step and the closed loop around it.  TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/HumanoidNavigation/MPC/HumanoidMpc.py`:
* one step = `:387` half-planes at the current CoM -> `:406-411` heading schedule -> `:417` solve ->
  `:432-447` record u*_0, integrate one LIP step, theta <- theta_1                      -> `mpc_step`
* loop control `:380-459` (stop when the previous objective < 0.05, `mpc_step > 1` sub-stepping, output
  trimming) and the foot-parity window `:401-403`                                        -> `run_simulation`
* sub-goal sequencing `MPC/HumanoidMPCVariants/HumanoidMPCWithRRT.py:153-181`           -> `run_subgoals`
"""
import numpy as np

from .halfplane import half_planes
from .model import STOP_OBJECTIVE, default_conf, foot_parity, heading_schedule, lip_matrices
from .qp import assemble_condensed, predicted_states, solve_exact


def mpc_step(state, goal, obstacles, foot, N=3, sampling_time=None, conf=None, delta=0.0, c_eta=None):
    """One MPC step for one scenario.

    state[5] = (p_x, v_x, p_y, v_y, theta); foot[N+1] = parity window; obstacles = list of vertex rings
    (hull, counter-clockwise).  `c_eta=(c, eta)` overrides the half-plane computation (unknown-environment
    variant feeds inferred hulls).  Returns dict(status, U[N,2], X[N+1,4], theta, omega, c, eta, obj, x_next[5]).
    """
    conf = conf or default_conf()
    sampling_time = conf["DELTA_T"] if sampling_time is None else sampling_time
    state = np.asarray(state, dtype=np.float64)
    x0, th0 = state[:4], state[4]
    if c_eta is None:
        c, eta = half_planes(np.array([x0[0], x0[2]]), obstacles)
    else:
        c, eta = c_eta
    theta, omega = heading_schedule(x0, th0, goal, N, sampling_time, conf)
    out = dict(theta=theta, omega=omega, c=c, eta=eta)
    if not np.all(np.isfinite(eta)):
        out.update(status=3, U=None, X=None, obj=np.nan, x_next=None, qp=None, sol=None)
        return out
    qp = assemble_condensed(x0, theta, omega, foot, c, eta, goal, conf, delta=delta)
    sol = solve_exact(qp)
    out.update(status=sol["status"], qp=qp, sol=sol, obj=sol["obj"])
    if sol["status"] != 0:
        out.update(U=None, X=None, x_next=None)
        return out
    z = sol["z"]
    X = predicted_states(qp, x0, z)
    A, B = lip_matrices(conf)
    x_next = np.empty(5)
    x_next[:4] = A @ x0 + B @ z[:2]          # HumanoidMpc.py:441-442
    x_next[4] = theta[1]                     # :447
    out.update(U=z.reshape(N, 2), X=X, x_next=x_next)
    return out


def run_simulation(goal, obstacles, init_state, N_horizon=3, N_mpc_timesteps=100, sampling_time=1e-3,
                   start_with_right_foot=True, conf=None, delta=0.0, info=None):
    """Closed loop of `HumanoidMPC.run_simulation` (HumanoidMpc.py:345-459).  Returns X_pred[5,K+1], U_pred[3,K].
    `info` (optional dict) receives `solves` (MPC solves executed, the failed one included) and `end`
    ("stop_rule", "step_budget" or "status<k>" of the failed solve)."""
    conf = conf or default_conf()
    mpc_steps = int(conf["DELTA_T"] / sampling_time) or 1                    # :74-75
    num_inputs = mpc_steps * N_mpc_timesteps                                 # :78
    s_v = foot_parity(num_inputs + N_horizon + 1, start_with_right_foot, conf)
    X_pred = np.zeros((5, num_inputs + 1))
    U_pred = np.zeros((3, num_inputs))
    X_pred[:, 0] = np.asarray(init_state, dtype=np.float64)
    last_obj = float("inf")
    u0 = np.zeros(2)
    foot = None
    k = 0
    solves, end = 0, "step_budget"
    for k in range(num_inputs):
        is_mpc = k % mpc_steps == 0
        if last_obj < STOP_OBJECTIVE:                                        # :392
            end = "stop_rule"
            break
        if is_mpc:
            step_number = k // mpc_steps
            foot = s_v[step_number:step_number + N_horizon + 1]             # :401-403
            r = mpc_step(X_pred[:, k], goal, obstacles, foot, N_horizon, sampling_time, conf, delta)
            solves += 1
            if r["status"] != 0:                                             # :419-429
                end = f"status{r['status']}"
                break
            last_obj = r["obj"]
            u0 = r["U"][0]
            theta, omega = r["theta"], r["omega"]
        else:
            from .model import heading_schedule as _hs
            theta, omega = _hs(X_pred[:4, k], X_pred[4, k], goal, N_horizon, sampling_time, conf)
        U_pred[:2, k] = u0                                                   # :432-433
        U_pred[2, k] = omega[0]
        if is_mpc:
            A, B = lip_matrices(conf)
            X_pred[:4, k + 1] = A @ X_pred[:4, k] + B @ u0                   # :441-442
        else:
            X_pred[:4, k + 1] = X_pred[:4, k]                                # :446
        X_pred[4, k + 1] = theta[1]                                          # :447
    if info is not None:
        info.update(solves=solves, end=end)
    return X_pred[:, :k + 1], U_pred[:, :k]                                  # :458-459


def run_subgoals(sub_goals, obstacles, N_horizon=3, N_mpc_timesteps=100, sampling_time=1e-3,
                 start_with_right_foot=True, conf=None, start_state=(0, 0, 0, 0, 0), delta=0.0):
    """Sequential fresh runs per sub-goal with state carry-over (HumanoidMPCWithRRT.py:153-181)."""
    Xg, Ug = None, None
    for sg in sub_goals:
        X, U = run_simulation(sg, obstacles, start_state, N_horizon, N_mpc_timesteps, sampling_time,
                              start_with_right_foot, conf, delta)
        start_state = tuple(X[:, -1])                                        # :178
        Xg = X if Xg is None else np.concatenate((Xg, X), axis=1)
        Ug = U if Ug is None else np.concatenate((Ug, U), axis=1)
    return Xg, Ug
