"""Mirror of the reference's `MPC/HumanoidMpc.py`: same `conf`, class, constructor and `run_simulation` contract.

What changed underneath (reference lines in brackets): the CasADi `Opti` problem and IPOPT solve [:97-135, :417]
are replaced by the fused CUDA step (`ldcbf_b200.mpc_qp`: heading schedule [:137-160], constraint rows
[:162-249, :263-294], cost [:321-333], exact QP solve, LIP integration [:335-343]); the per-obstacle Python
loop [:296-319] by the K1 kernel.  The loop control of `run_simulation` [:380-459] is kept statement for
statement, including the `mpc_step` sub-stepping, the `< 0.05` stop, the break on a failed solve and the output
trimming.  When no subclass hook is overridden the whole loop runs inside one kernel launch
(`ldcbf_b200.rollout`).
"""
import math
import os
from typing import Union

import numpy as np
import torch
from yaml import safe_load

import ldcbf_b200
from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils, hull_ring

this_dir = os.path.dirname(os.path.realpath(__file__))
config_dir = os.path.dirname(this_dir)
with open(config_dir + '/config.yml', 'r') as file:
    conf = safe_load(file)
conf["BETA"] = np.sqrt(conf["GRAVITY_CONST"] / conf["COM_HEIGHT"])
conf["OMEGA_MAX"] = 0.156 * math.pi
conf["OMEGA_MIN"] = -conf["OMEGA_MAX"]

ASSETS_PATH = os.path.dirname(config_dir) + "/Assets/Animations/res.gif"

# The reference hands the QP to an interior-point solver (IPOPT, tol 1e-5, :98-100): its iterates stay strictly inside
# every LDCBF half-plane, so the CoM never comes to lie exactly ON an obstacle edge.  The exact solver used here puts
# it there whenever an LDCBF row is active, and the next step's normal (x - c)/||x - c|| [ObstaclesUtils.py:98-107]
# is then decided by rounding (inside / outside): the closed loop of the basic simulation ends "infeasible" around
# step 28 instead of at the goal.  The mirror therefore keeps the clearance an interior-point iterate has, 1e-6 m:
# within BASELINE.json's tolerances (LDCBF rows to 1e-6, trajectories to 1e-4 m), and the basic simulation reaches the
# goal in 85 steps (reference: 86).  Set to 0.0 for the bare QP; the C ABI and the batched front-end take delta as given.
INTERIOR_MARGIN = 1e-6


class HumanoidMPC:
    """MPC of "Real-Time Safe Bipedal Robot Navigation using Linear Discrete Control Barrier Functions" (Peng et al.)."""

    def __init__(self, goal, obstacles, N_horizon=3, N_mpc_timesteps=100, sampling_time=1e-3,
                 init_state: Union[np.ndarray, tuple] = np.array([0, 0, 0, 0, 0]),
                 start_with_right_foot: bool = True, verbosity: int = 1):
        assert conf['DELTA_T'] % sampling_time <= 1e-8, \
            "The sampling time must be lower than and divisible by the duration of the step."
        if not torch.cuda.is_available():
            raise RuntimeError("HumanoidMPC runs its step on the GPU (ldcbf_b200); no CUDA device is visible")
        self.N_horizon = N_horizon
        self.N_simul = N_mpc_timesteps
        self.sampling_time = sampling_time
        self.mpc_step = int((conf['DELTA_T'] / self.sampling_time))
        self.mpc_step = 1 if self.mpc_step == 0 else self.mpc_step
        self.num_inputs = self.mpc_step * self.N_simul
        self.start_with_right_foot = start_with_right_foot
        self.verbosity = verbosity
        self.goal = goal
        self.obstacles = obstacles
        self.list_inferred_obstacles = []
        self.list_lidar_readings = []
        self.precomputed_omega = None
        self.precomputed_theta = None
        assert (self.goal is not None and self.obstacles is not None)
        self.state_dim = 4
        self.control_dim = 2
        self.s_v = [conf["RIGHT_FOOT"] if i % 2 == (0 if start_with_right_foot else 1) else conf["LEFT_FOOT"]
                    for i in range(self.num_inputs + self.N_horizon + 1)]
        if isinstance(init_state, tuple):
            init_state = np.array(init_state)
        init_state = np.asarray(init_state, dtype=np.float64)
        assert init_state.shape[0] == 5, "The initial state must be a vector with 5 components."
        self.init_state = init_state
        self.distance_from_obstacles = getattr(self, "distance_from_obstacles", 0.0)
        self.last_status = 0
        self._dev = torch.device("cuda")

    # ---- hooks kept from the reference ----------------------------------------------------------------------------
    def _get_list_c_and_eta(self, x_k: float, y_k: float):
        """Lists of c[2,1] and eta[2,1] per obstacle at the current CoM (reference :296-319) — one K1 launch."""
        c, eta = ObstaclesUtils.closest_points_and_normals(np.array([x_k, y_k]), self.obstacles)
        return [ci.reshape(2, 1) for ci in c], [ei.reshape(2, 1) for ei in eta]

    def _delta(self):
        """LDCBF margin handed to the kernels: the variant's `distance_from_obstacles` plus INTERIOR_MARGIN."""
        return float(self.distance_from_obstacles) + float(INTERIOR_MARGIN)

    def _params(self):
        # conf is read at call time: bounds_tuning.py:22-26 mutates it between runs
        return ldcbf_b200.params_from_conf(conf, self.sampling_time)

    def _solve(self, state, foot, list_c, list_eta):
        """One K2+K3 launch for this scenario given the half-planes."""
        n = len(list_c)
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=self._dev)
        ce = np.zeros((1, max(n, 1), 4))                 # any number of obstacles: 8 in registers, the rest streamed
        for o in range(n):
            ce[0, o, :2] = np.asarray(list_c[o]).ravel()
            ce[0, o, 2:] = np.asarray(list_eta[o]).ravel()
        out = ldcbf_b200.mpc_qp(self._params(), t(state[None, :4]), t(state[None, 4]), t(np.asarray(self.goal)[None, :]),
                                t(np.asarray(foot)[None, :], torch.int8), t(ce), t([n], torch.int32),
                                delta=t([self._delta()]))
        return {k: v.cpu().numpy()[0] for k, v in out.items()}

    def _hooks_overridden(self):
        return type(self)._get_list_c_and_eta is not HumanoidMPC._get_list_c_and_eta

    # ---- closed loop ----------------------------------------------------------------------------------------------
    def run_simulation(self, path_to_gif: str = None, make_fast_plot: bool = True, plot_animation: bool = False,
                       fill_animator: bool = True, initial_animator=None):
        """Returns (X_pred[5,K+1], U_pred[3,K], animator) like the reference (:345-494); animator is passed through
        (plotting is out of scope)."""
        # any number of obstacles runs fused (8 register-resident, the rest streamed); long horizons use the
        # block-per-scenario step kernel, one launch per MPC timestep
        if self._hooks_overridden() or self.N_horizon > ldcbf_b200.binding.MAX_HORIZON:
            X_pred, U_pred = self._run_stepwise()
        else:
            X_pred, U_pred = self._run_fused()
        return X_pred, U_pred, initial_animator

    def _run_fused(self):
        """Whole loop in one launch (csrc/rollout.cu), then the reference's output trimming (:458-459)."""
        from ldcbf_b200.scenarios import pack_rings
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=self._dev)
        verts, nverts, nobs = pack_rings([[hull_ring(o) for o in self.obstacles]])
        state = t(self.init_state[None, :])
        r = ldcbf_b200.rollout(self._params(), state, t(np.asarray(self.goal, dtype=np.float64)[None, None, :]),
                               t([1 if self.start_with_right_foot else 0], torch.int8), t(verts),
                               t(nverts, torch.int32), t(nobs, torch.int32), T=self.num_inputs, N=self.N_horizon,
                               max_steps_per_goal=self.num_inputs,
                               delta=t([self._delta()]))
        K = int(r["steps"].item())
        self.last_status = int(r["status"].item())
        X = r["traj_X"][0].cpu().numpy().T            # [5, T+1]
        U = r["traj_U"][0].cpu().numpy().T            # [3, T]
        if K == self.num_inputs:                      # loop ran to completion: k = num_inputs-1 (:383 defect kept)
            return X[:, :K], U[:, :K - 1]
        return X[:, :K + 1], U[:, :K]

    def _run_stepwise(self):
        """The reference's loop (:380-459) with the subclass hooks in the loop, one solve launch per MPC timestep."""
        X_pred = np.zeros((self.state_dim + 1, self.num_inputs + 1))
        U_pred = np.zeros((self.control_dim + 1, self.num_inputs))
        X_pred[:, 0] = self.init_state
        last_obj = float('inf')
        sol = None
        k = 0
        for k in range(self.num_inputs):
            is_mpc_timestep = k % self.mpc_step == 0
            list_c, list_eta = self._get_list_c_and_eta(x_k=X_pred[0, k], y_k=X_pred[2, k])       # :387
            if last_obj < 0.05:                                                                     # :392
                break
            if is_mpc_timestep:
                step_number = math.floor(k / self.mpc_step)
                foot = self.s_v[step_number:step_number + self.N_horizon + 1]                       # :401-403
                sol = self._solve(X_pred[:, k], foot, list_c, list_eta)
                self.last_status = int(sol["status"])
                if sol["status"] != 0:                                                              # :419-429
                    if self.verbosity > 0:
                        print(f"===== ERROR ({k}) ===== solver status {int(sol['status'])} "
                              "(1 max-iter, 2 infeasible, 3 degenerate geometry)")
                    break
                last_obj = float(sol["obj"])
                self.precomputed_theta, self.precomputed_omega = sol["theta"], sol["omega"]
                theta1, omega0 = sol["theta"][1], sol["omega"][0]
            else:
                target = math.atan2(self.goal[1] - X_pred[2, k], self.goal[0] - X_pred[0, k]) - X_pred[4, k]
                omega0 = min(max(target, conf["OMEGA_MIN"]), conf["OMEGA_MAX"])                    # :150-156
                theta1 = X_pred[4, k] + omega0 * self.sampling_time
            U_pred[:2, k] = sol["U"][0]                                                             # :432-433
            U_pred[2, k] = omega0
            if is_mpc_timestep:
                X_pred[:4, k + 1] = sol["X"][1]                                                     # :441-442
            else:
                X_pred[:4, k + 1] = X_pred[:4, k]                                                   # :446
            X_pred[4, k + 1] = theta1                                                               # :447
        return X_pred[:, :k + 1], U_pred[:, :k]                                                     # :458-459
