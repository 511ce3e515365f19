"""Batched front-end: B independent scenarios, one MPC step or a whole closed loop per call.

This is the new caller of the C ABI next to the reference-shaped `HumanoidNavigation.MPC.HumanoidMpc.HumanoidMPC`
(B = 1).  It owns the device-resident scenario tensors (laid out for the kernels) and pinned
host staging buffers for the end-to-end path.
"""
import numpy as np
import torch

from . import binding as _b


class BatchedHumanoidMPC:
    """B scenarios sharing N_horizon and the solver parameters.

    obstacles are given packed: verts[B,max_obs,max_verts,2] (hull vertices, CCW, zero padded), nverts[B,max_obs],
    nobs[B] (see `scenarios.pack_rings`).  All arrays may be numpy (copied to the device once) or CUDA tensors.
    """

    def __init__(self, goal, verts, nverts, nobs, N_horizon=3, sampling_time=0.4, conf=None, delta=None,
                 limits=None, device="cuda", **param_overrides):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedHumanoidMPC needs a CUDA device: ldcbf_b200 has no CPU path")
        self.device = torch.device(device)
        self.N = int(N_horizon)
        self.prm = (_b.params_from_conf(conf, sampling_time, **param_overrides) if conf is not None
                    else _b.default_params(sampling_time, **param_overrides))
        dev = self._dev
        self.goal = dev(goal, torch.float64)
        self.verts = dev(verts, torch.float64)
        self.nverts = dev(nverts, torch.int32)
        self.nobs = dev(nobs, torch.int32)
        self.delta = None if delta is None else dev(delta, torch.float64)
        self.limits = None if limits is None else dev(limits, torch.float64)
        self.B = self.goal.shape[0]
        self._out = None
        self._pinned = None
        self._step_graph, self._step_graph_key, self._step_graph_ok = None, None, None

    def _scene_key(self):
        """Addresses of the scenario tensors a captured graph has baked in (a replaced tensor forces a re-capture)."""
        return tuple(0 if t is None else t.data_ptr() for t in (self.goal, self.verts, self.nverts, self.nobs,
                                                                 self.delta, self.limits)) + (bytes(self.prm), self.N)

    def _dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    # ---- device-resident step ---------------------------------------------------------------------------
    def step(self, x0, theta0, foot, goal=None, graph=False):
        """One MPC step for all scenarios.  x0[B,4], theta0[B], foot[B,N+1] int8 CUDA tensors.
        Returns dict(U, X, theta, omega, c_eta, obj, status, iters) of CUDA tensors (reused between calls).

        graph=True: the two launches (K1, K2+K3) are captured once into a CUDA graph, keyed on the addresses of the
        input tensors, and replayed while the caller keeps updating those tensors in place — for small batches the
        time of a step is otherwise mostly the Python-side marshalling of the 24 arguments between the launches."""
        g = self.goal if goal is None else goal
        if not graph or self.B >= 148 * 2 * 128 * 4 or self._step_graph_ok is False:
            self._out = _b.mpc_step(self.prm, x0, theta0, g, foot, self.verts, self.nverts, self.nobs,
                                    delta=self.delta, limits=self.limits, out=self._out)
            return self._out
        key = (x0.data_ptr(), theta0.data_ptr(), foot.data_ptr(), g.data_ptr(), tuple(x0.shape)) + self._scene_key()
        if self._step_graph_key != key:
            self._out = _b.mpc_step(self.prm, x0, theta0, g, foot, self.verts, self.nverts, self.nobs,
                                    delta=self.delta, limits=self.limits, out=self._out)      # eager once: allocations
            torch.cuda.current_stream().synchronize()
            try:
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    _b.mpc_step(self.prm, x0, theta0, g, foot, self.verts, self.nverts, self.nobs, delta=self.delta,
                                limits=self.limits, out=self._out)
                self._step_graph, self._step_graph_key = cg, key
            except Exception:
                self._step_graph, self._step_graph_key, self._step_graph_ok = None, None, False
                torch.cuda.synchronize()
                return self.step(x0, theta0, foot, goal)
        self._step_graph.replay()
        return self._out

    # ---- end-to-end step: host buffers in, host buffers out -----------------------------------------------
    def step_host(self, state_host):
        """state_host: pinned [B,6] fp64 tensor (p_x, v_x, p_y, v_y, theta, first stance foot +-1).
        One upload, the loop-shaped step (`ldcbf_mpc_step_packed_f64`), one download and a stream synchronise.
        Returns the pinned host tensor next[B,10] = (x_next[4], theta_1, u0_x, u0_y, omega_0, objective, status),
        reused between calls.

        The four stream operations (H2D copy, K1, K2+K3, D2H copy) are captured once into a CUDA graph, keyed on the
        address of `state_host`, and replayed on later calls with the same buffer (a loop that updates its state
        tensor in place): one graph launch per step instead of four enqueues through Python.  A different buffer,
        a batch large enough for the two-kernel solve (which allocates stream-ordered memory) or a failed capture
        use the plain enqueue path."""
        B = self.B
        if self._pinned is None:
            self._pinned = torch.empty((B, 10), dtype=torch.float64).pin_memory()
            self._d_state = torch.empty((B, 6), dtype=torch.float64, device=self.device)
            self._packed = None
            self._graph, self._graph_key, self._graph_ok = None, None, B < 148 * 2 * 128 * 4
        key = (state_host.data_ptr(), tuple(state_host.shape)) + self._scene_key()
        if self._graph_ok and self._graph_key != key:
            self._enqueue_host_step(state_host)               # eager once: allocations, lazy module loading
            torch.cuda.current_stream().synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue_host_step(state_host)
                self._graph, self._graph_key = g, key
            except Exception:                                 # capture refused: keep the plain path
                self._graph, self._graph_key, self._graph_ok = None, None, False
                torch.cuda.synchronize()
        if self._graph is not None and self._graph_key == key:
            self._graph.replay()
        else:
            self._enqueue_host_step(state_host)
        torch.cuda.current_stream().synchronize()
        return self._pinned

    def _enqueue_host_step(self, state_host):
        self._d_state.copy_(state_host, non_blocking=True)
        o = self._packed = _b.mpc_step_packed(self.prm, self._d_state, self.goal, self.verts, self.nverts, self.nobs,
                                              N=self.N, delta=self.delta, limits=self.limits, out=self._packed)
        self._pinned.copy_(o["next"], non_blocking=True)

    @property
    def h2d_bytes_per_step(self):
        return self.B * 6 * 8

    @property
    def d2h_bytes_per_step(self):
        return self.B * 10 * 8

    # ---- closed loop -----------------------------------------------------------------------------------------
    def rollout(self, state, right_first, T, goals=None, max_steps_per_goal=None, record=True):
        """Closed loop (HumanoidMpc.py:380-459) for all scenarios in one launch.  state[B,5] CUDA, updated in place."""
        goals = self.goal[:, None, :].contiguous() if goals is None else goals
        return _b.rollout(self.prm, state, goals, right_first, self.verts, self.nverts, self.nobs, T, N=self.N,
                          max_steps_per_goal=max_steps_per_goal, delta=self.delta, limits=self.limits, record=record)


    def rollout_host(self, state_host, right_first_host, T, goals=None, max_steps_per_goal=None):
        """End-to-end closed loop: pinned host state[B,5] and right_first[B] (int8) in, one upload, ONE launch for all
        the steps of all scenarios, one download and a stream synchronise.  Returns the pinned host tensor
        result[B,8] = (final state[5], executed steps, status of the last solve, LDCBF_END_* code), reused between
        calls."""
        B = self.B
        if getattr(self, "_ro_pinned", None) is None:
            self._ro_pinned = torch.empty((B, 8), dtype=torch.float64).pin_memory()
            self._ro_state = torch.empty((B, 5), dtype=torch.float64, device=self.device)
            self._ro_rf = torch.empty((B,), dtype=torch.int8, device=self.device)
            self._ro_res = torch.empty((B, 8), dtype=torch.float64, device=self.device)
        self._ro_state.copy_(state_host, non_blocking=True)
        self._ro_rf.copy_(right_first_host, non_blocking=True)
        r = self.rollout(self._ro_state, self._ro_rf, T, goals=goals, max_steps_per_goal=max_steps_per_goal, record=False)
        res = self._ro_res
        res[:, :5] = self._ro_state
        res[:, 5] = r["steps"]
        res[:, 6] = r["status"]
        res[:, 7] = r["end_code"]
        self._ro_pinned.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._ro_last = r
        return self._ro_pinned

    rollout_h2d_bytes = property(lambda self: self.B * (5 * 8 + 1))
    rollout_d2h_bytes = property(lambda self: self.B * 8 * 8)


class BatchedUnknownEnvMPC(BatchedHumanoidMPC):
    """Unknown-environment variant, batched and entirely on the device: every step scans the true map with the LiDAR
    caster (K4), clusters the readings and builds the convex hulls (f1), and solves the MPC step against the
    *inferred* obstacles (K1 + K2+K3) — `HumanoidMPCUnknownEnvironment._get_list_c_and_eta` + one loop iteration of
    the reference, for B scenarios at once.  `verts` are the true obstacles in the order the reference casts
    against (`ConvexHull.points` rows)."""

    def __init__(self, goal, verts, nverts, nobs, lidar_range=3.0, lidar_resolution=360, max_hulls=8,
                 max_hull_verts=64, eps=0.3, min_samples=3, **kw):
        super().__init__(goal, verts, nverts, nobs, **kw)
        self.lidar_range = float(lidar_range)
        self.rays = _b.ray_table(lidar_range, lidar_resolution).to(self.device)
        self.max_hulls, self.max_hull_verts = max_hulls, max_hull_verts
        self.eps, self.min_samples = eps, min_samples
        self.check_overflow = True

    def sense(self, pos, noise=None):
        """pos[B,2] -> dict(hit_obs, hit_edge, hit_xy, labels, verts, nverts, nobs, overflow) (inferred obstacles)."""
        ho, he, xy = _b.lidar_cast(pos, self.verts, self.nverts, self.nobs, self.lidar_range, rays=self.rays)
        cl = _b.lidar_clusters(xy, noise=noise, eps=self.eps, min_samples=self.min_samples,
                               max_hulls=self.max_hulls, max_hull_verts=self.max_hull_verts)
        cl.update(hit_obs=ho, hit_edge=he, hit_xy=xy)
        return cl

    def step(self, x0, theta0, foot, goal=None, noise=None):
        """One step for all scenarios.  Raises when a scan produced more hulls / hull vertices than max_hulls /
        max_hull_verts hold (the reference constrains against EVERY inferred obstacle): rebuild the engine with larger
        capacities.  The check reads one flag back from the device; `check_overflow=False` on the instance skips it."""
        pos = x0[:, [0, 2]].contiguous()
        sensed = self.sense(pos, noise)
        if self.check_overflow and bool(sensed["overflow"].any()):
            raise RuntimeError(f"LiDAR clustering overflow: more than max_hulls={self.max_hulls} hulls or "
                               f"max_hull_verts={self.max_hull_verts} vertices in a scan; raise the capacities")
        self._out = _b.mpc_step(self.prm, x0, theta0, self.goal if goal is None else goal, foot, sensed["verts"],
                                sensed["nverts"], sensed["nobs"], delta=self.delta, limits=self.limits, out=self._out)
        out = dict(self._out)
        out["sensed"] = sensed
        return out

    def rollout(self, state, right_first, T, noise=None, record=True):
        """Closed loop of the variant for all scenarios, entirely on the device (`ldcbf_rollout_unknown_f64`): every step
        scans the true map from the current CoM, infers the obstacles and solves against them.  state[B,5] CUDA,
        updated in place; `noise` [B,R,2] is added to the readings of every step.  Raises on clustering overflow."""
        r = _b.rollout_unknown(self.prm, state, self.goal, right_first, self.verts, self.nverts, self.nobs, T, self.rays,
                               self.lidar_range, N=self.N, noise=noise, eps=self.eps, min_samples=self.min_samples,
                               max_hulls=self.max_hulls, max_hull_verts=self.max_hull_verts, delta=self.delta,
                               limits=self.limits, record=record)
        if self.check_overflow and bool(r["overflow"].any()):
            raise RuntimeError(f"LiDAR clustering overflow during the rollout (max_hulls={self.max_hulls}, "
                               f"max_hull_verts={self.max_hull_verts}); raise the capacities")
        return r
