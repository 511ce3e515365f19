"""Mirror of the sub-goal sequencing of `MPC/HumanoidMPCVariants/HumanoidMPCWithRRT.py:153-181`.

The occupancy grid, clearance cost and RRT* planning (reference :21-135, third-party `rrtplanner==0.1.2`) are out
of scope: the way-points are an input.  The sequencing itself — a fresh MPC per sub-goal,
state carried over, foot parity restarted, start hard-coded to the origin (:155) — runs inside one rollout launch.
"""
import numpy as np
import torch

import ldcbf_b200
from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
from HumanoidNavigation.Utils.ObstaclesUtils import hull_ring


class HumanoidMPCWithRRT(HumanoidMPC):
    def run_simulation(self, path_to_gif: str = None, make_fast_plot: bool = True, plot_animation: bool = False,
                       fill_animator: bool = True, initial_animator=None, visualize_rrt_path: bool = False,
                       path_to_rrt_pdf: str = None, sub_goals=None):
        if sub_goals is None:
            raise NotImplementedError("RRT* planning is out of scope here (rrtplanner is a third-party dependency of "
                                      "the reference); pass the way-points as sub_goals=[(x, y), ...]")
        from ldcbf_b200.scenarios import pack_rings
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=self._dev)
        goals = np.asarray(sub_goals, dtype=np.float64).reshape(1, -1, 2)
        G = goals.shape[1]
        verts, nverts, nobs = pack_rings([[hull_ring(o) for o in self.obstacles]])
        state = t(np.zeros((1, 5)))                                            # start_state = (0,0,0,0,0), :155
        T = G * self.num_inputs
        r = ldcbf_b200.rollout(self._params(), state, t(goals), t([1 if self.start_with_right_foot else 0], torch.int8),
                               t(verts), t(nverts, torch.int32), t(nobs, torch.int32), T=T, N=self.N_horizon,
                               max_steps_per_goal=self.num_inputs)
        gs = r["goal_steps"][0].cpu().numpy()
        X = r["traj_X"][0].cpu().numpy().T
        U = r["traj_U"][0].cpu().numpy().T
        # the reference concatenates the per-run arrays, so the junction state appears twice (:180-181)
        Xs, Us, s = [], [], 0
        for g in range(G):
            k = int(gs[g])
            if k == self.num_inputs:                                           # run exhausted: last column dropped
                Xs.append(X[:, s:s + k]); Us.append(U[:, s:s + k - 1])
            else:
                Xs.append(X[:, s:s + k + 1]); Us.append(U[:, s:s + k])
            s += k
        return np.concatenate(Xs, axis=1), np.concatenate(Us, axis=1), initial_animator
