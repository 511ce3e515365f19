"""Mirror of `report_simulations/bounds_tuning.py`: the reference's own batch workload.

The reference runs its 16 x 4 x 35 x 12 = 26 880 hyper-parameter combinations one after the other, mutating the module
global `conf` (ALPHA, V_MAX, OMEGA_MAX/MIN) before each full closed-loop simulation (reference `:13-46`).  Here every
combination is one scenario of ONE rollout launch: the limits are per-scenario inputs of the kernels
(`limits[b] = (ALPHA, V_MAX[0], V_MAX[1], OMEGA_MAX, OMEGA_MIN, -)`), the loop control (sampling_time = 0.1 ->
mpc_step = 4, 300 MPC timesteps, the `< 0.05` stop) runs on the device, and the selection rule of `:43-48` is applied
to the returned trajectories.  Plotting the winner (`:57-73`) is out of scope.
"""
import itertools

import numpy as np
import torch

import ldcbf_b200
from HumanoidNavigation.MPC import HumanoidMpc


def hyperparameter_grid():
    v_max_x = np.arange(0.2, 1, 0.05)
    v_max_y = np.arange(0.2, 0.4, 0.05)
    alpha = np.arange(0.5, 4, 0.1)
    omega_range = np.arange(0.4, 1, 0.05)
    return np.array(list(itertools.product(v_max_x, v_max_y, alpha, omega_range)))      # same order as the reference


def bounds_tuning(grid=None, goal=(5, 5), init_state=(0, 0, 0, 0, 0), N_horizon=3, N_mpc_timesteps=300,
                  sampling_time=1e-1, return_all=False):
    """Returns (best_combination, best_res) like the reference prints them (`:50-51`); with `return_all` also the
    per-combination score array (inf where the run did not end within 1 m of the goal in both coordinates)."""
    conf = HumanoidMpc.conf
    grid = hyperparameter_grid() if grid is None else np.asarray(grid, dtype=np.float64)
    B = len(grid)
    dev = torch.device("cuda")
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    limits = np.full((B, 6), np.nan)
    limits[:, 0] = grid[:, 2]          # ALPHA
    limits[:, 1] = grid[:, 0]          # V_MAX[0]
    limits[:, 2] = grid[:, 1]          # V_MAX[1]
    limits[:, 3] = grid[:, 3]          # OMEGA_MAX
    limits[:, 4] = -grid[:, 3]         # OMEGA_MIN
    sub = int(conf["DELTA_T"] / sampling_time) or 1
    T = sub * N_mpc_timesteps
    prm = ldcbf_b200.params_from_conf(conf, sampling_time)
    state = t(np.tile(np.asarray(init_state, dtype=np.float64), (B, 1)))
    goals = t(np.tile(np.asarray(goal, dtype=np.float64), (B, 1, 1)))
    verts = torch.zeros((B, 1, 1, 2), dtype=torch.float64, device=dev)            # obstacles = []
    nverts = torch.zeros((B, 1), dtype=torch.int32, device=dev)
    nobs = torch.zeros((B,), dtype=torch.int32, device=dev)
    r = ldcbf_b200.rollout(prm, state, goals, torch.ones(B, dtype=torch.int8, device=dev), verts, nverts, nobs, T=T,
                           N=N_horizon, max_steps_per_goal=T, limits=t(limits))
    steps = r["steps"].cpu().numpy()
    X = r["traj_X"]                                                                # [B, T+1, 5]
    # the reference's X_pred_glob[:, -1] is column K (or K-1 when the loop ran to completion, HumanoidMpc.py:458)
    last = np.where(steps == T, steps - 1, steps)
    idx = torch.as_tensor(last, device=dev, dtype=torch.long)
    final = X[torch.arange(B, device=dev), idx][:, [0, 2]].cpu().numpy()
    reached = np.all((final - np.asarray(goal)) ** 2 <= 1, axis=1)                # `:43`
    ncol = np.minimum(last + 1, 50)                                                # X_pred_glob[3, :50]
    vy = X[:, :50, 3].abs().cpu().numpy()
    mask = np.arange(50)[None, :] < ncol[:, None]
    val = (vy * mask).sum(1) / ncol
    score = np.where(reached, val, np.inf)
    best = int(np.argmin(score))                                                   # first strict minimum, like `:46`
    out = (tuple(grid[best]), float(score[best]))
    return out + (score, steps) if return_all else out


if __name__ == "__main__":
    best_combin, best_res = bounds_tuning()
    print(f'best_combination: {best_combin}')
    print(f'best_res: {best_res}')
