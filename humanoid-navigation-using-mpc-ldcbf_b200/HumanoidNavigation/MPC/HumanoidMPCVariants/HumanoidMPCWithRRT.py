"""Mirror of `MPC/HumanoidMPCVariants/HumanoidMPCWithRRT.py`.

* `_build_occupancy_grid` (:21-88) and the clearance cost (:103-108) run on the GPU (`ldcbf_clearance_grid_f64`,
  bit-exact grid and distances).
* The RRT* search (:110-135) is the third-party `rrtplanner==0.1.2` in the reference, absent offline: `rrt_star.py` is an
  independent host implementation with the same parameters and edge cost; its random tree is not the reference's
  (parity unpinned).  Way-points can also be passed in directly as `sub_goals=[...]`.
* The sequencing (:153-181) — a fresh MPC per sub-goal, state carried over, foot parity restarted, start hard-coded
  to the origin (:155) — runs inside one rollout launch.
"""
import numpy as np
import torch

import ldcbf_b200
from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
from HumanoidNavigation.Utils.ObstaclesUtils import hull_ring


class HumanoidMPCWithRRT(HumanoidMPC):
    def _clearance(self, width_grid_size):
        from ldcbf_b200.scenarios import pack_rings
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=self._dev)
        verts, nverts, nobs = pack_rings([[hull_ring(o) for o in self.obstacles]])
        r = ldcbf_b200.clearance_grid(t(np.asarray(self.goal, dtype=np.float64)[None, :]), t(verts),
                                      t(nverts, torch.int32), t(nobs, torch.int32), width=width_grid_size)
        min_x, min_y, max_x, max_y, height, _ = r["meta"][0].cpu().numpy()
        height = int(height)
        cut = lambda a: a[0, :, :height + 1].cpu().numpy()
        return cut(r["og"]).astype(np.float64), cut(r["dist"]), cut(r["cost"]), (min_x, min_y, max_x, max_y, height)

    def _build_occupancy_grid(self, width_grid_size: int):
        """(occupancy_grid[width+1, height+1], world -> grid, grid -> world) like the reference (:21-88)."""
        og, _, _, (min_x, min_y, max_x, max_y, height) = self._clearance(width_grid_size)
        fwd = lambda x_glob, y_glob: np.array([                                                     # :55-58
            np.round(((x_glob - min_x) / (max_x - min_x)) * width_grid_size),
            np.round(((y_glob - min_y) / (max_y - min_y)) * height),
        ]).astype(int)
        inv = lambda x_og, y_og: np.array([                                                         # :60-63
            min_x + ((x_og * (max_x - min_x)) / width_grid_size),
            min_y + ((y_og * (max_y - min_y)) / height),
        ])
        return og, fwd, inv

    def plan_sub_goals(self, width_grid_size=250, n=1500, r_rewire=80, seed=1, shortcut=True):
        """Way-points from the start (0,0) to the goal (:97-135): GPU map + clearance cost, host RRT*
        (+ a cost-non-increasing shortcut pass, see rrt_star.RRTStar.shortcut)."""
        from HumanoidNavigation.MPC.HumanoidMPCVariants.rrt_star import RRTStar
        og, fwd, inv = self._build_occupancy_grid(width_grid_size)
        _, _, costs, _ = self._clearance(width_grid_size)
        planner = RRTStar(og, costs, n=n, r_rewire=r_rewire, seed=seed)
        path = planner.plan(fwd(0, 0), fwd(self.goal[0], self.goal[1]))
        if path is None:
            raise RuntimeError("RRT*: no collision-free path found")
        if shortcut:
            path = planner.shortcut(path)
        return np.array([inv(p[0], p[1]) for p in path[1:]])                                       # edge end points, :131-135

    def run_simulation(self, path_to_gif: str = None, make_fast_plot: bool = True, plot_animation: bool = False,
                       fill_animator: bool = True, initial_animator=None, visualize_rrt_path: bool = False,
                       path_to_rrt_pdf: str = None, sub_goals=None):
        if sub_goals is None:
            sub_goals = self.plan_sub_goals()
        self.sub_goals = np.asarray(sub_goals, dtype=np.float64)
        if self.N_horizon > ldcbf_b200.binding.MAX_HORIZON or self._hooks_overridden():
            return self._run_per_sub_goal() + (initial_animator,)
        from ldcbf_b200.scenarios import pack_rings
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=self._dev)
        goals = np.asarray(sub_goals, dtype=np.float64).reshape(1, -1, 2)
        G = goals.shape[1]
        verts, nverts, nobs = pack_rings([[hull_ring(o) for o in self.obstacles]])
        state = t(np.zeros((1, 5)))                                            # start_state = (0,0,0,0,0), :155
        T = G * self.num_inputs
        # every sub-goal run is a plain HumanoidMPC in the reference (:160-168): same LDCBF margin as the base class
        r = ldcbf_b200.rollout(self._params(), state, t(goals), t([1 if self.start_with_right_foot else 0], torch.int8),
                               t(verts), t(nverts, torch.int32), t(nobs, torch.int32), T=T, N=self.N_horizon,
                               max_steps_per_goal=self.num_inputs, delta=t([self._delta()]))
        gs = r["goal_steps"][0].cpu().numpy()
        X = r["traj_X"][0].cpu().numpy().T
        U = r["traj_U"][0].cpu().numpy().T
        # the reference concatenates the per-run arrays, so the junction state appears twice (:180-181); a run that
        # used up num_inputs steps returns X_pred[:, :num_inputs] (:458) and the next run starts from ITS last column,
        # the state before the last integration (:178) - the kernel restarts from that state too
        Xs, Us, s, first = [], [], 0, X[:, 0]
        for g in range(G):
            k = int(gs[g])
            cols = np.concatenate((first[:, None], X[:, s + 1:s + k + 1]), axis=1)
            if k == self.num_inputs:                                           # run exhausted: last column dropped
                Xs.append(cols[:, :k]); Us.append(U[:, s:s + k - 1])
            else:
                Xs.append(cols); Us.append(U[:, s:s + k])
            first = Xs[-1][:, -1]
            s += k
        return np.concatenate(Xs, axis=1), np.concatenate(Us, axis=1), initial_animator

    def _run_per_sub_goal(self):
        """The reference's own sequencing (:153-181), one HumanoidMPC per sub-goal: used for horizons the fused rollout
        does not cover (N > 4) and when a subclass overrides the half-plane hook."""
        start_state = (0, 0, 0, 0, 0)                                          # :155
        Xg, Ug = None, None
        for sg in self.sub_goals:
            mpc = HumanoidMPC(goal=tuple(sg), obstacles=self.obstacles, N_horizon=self.N_horizon,
                              N_mpc_timesteps=self.N_simul, sampling_time=self.sampling_time, init_state=start_state,
                              start_with_right_foot=self.start_with_right_foot, verbosity=self.verbosity)
            X, U, _ = mpc.run_simulation(fill_animator=False)
            start_state = tuple(X[:, -1])                                      # :178
            Xg = X if Xg is None else np.concatenate((Xg, X), axis=1)
            Ug = U if Ug is None else np.concatenate((Ug, U), axis=1)
        return Xg, Ug
