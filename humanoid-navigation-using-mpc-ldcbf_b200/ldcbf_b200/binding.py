"""ctypes binding of the C ABI declared in include/ldcbf_mpc.h.  torch is used for device memory and streams only."""
import ctypes
import math
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDCBF_B200_LIB") or os.path.join(_HERE, "libldcbf_b200.so")   # env override: A/B builds

MAX_OBSTACLES = 8      # LDCBF_MAX_OBSTACLES of include/ldcbf_mpc.h
MAX_HORIZON = 4        # LDCBF_MAX_HORIZON: fused rollout / thread-per-scenario solver
MAX_HORIZON_LONG = 48  # LDCBF_MAX_HORIZON_LONG: block-per-scenario solver for 5..48

EXPORTS = ("ldcbf_abi_version", "ldcbf_params_default", "ldcbf_last_cuda_error", "ldcbf_workspace_bytes",
           "ldcbf_trim_workspace",
           "ldcbf_halfplanes_f64", "ldcbf_mpc_qp_f64", "ldcbf_mpc_step_f64", "ldcbf_mpc_step_packed_f64",
           "ldcbf_lidar_cast_f64",
           "ldcbf_lidar_clusters_f64", "ldcbf_clearance_grid_f64",
           "ldcbf_rollout_f64", "ldcbf_rollout_unknown_f64", "ldcbf_probe_fp64_fma")


class LdcbfParams(ctypes.Structure):
    """Mirror of `struct ldcbf_params` (include/ldcbf_mpc.h)."""
    _fields_ = [("delta_t", c_double), ("gravity", c_double), ("com_height", c_double), ("alpha", c_double),
                ("l_max_x", c_double), ("l_max_y", c_double), ("l_min_x", c_double), ("l_min_y", c_double),
                ("v_min", c_double * 2), ("v_max", c_double * 2), ("omega_max", c_double), ("omega_min", c_double),
                ("foot_offset", c_double), ("stop_objective", c_double), ("sampling_time", c_double),
                ("eps_active", c_double), ("eps_const_row", c_double), ("max_iter", c_int32), ("flags", c_int32),
                ("eps_infeasible", c_double)]


FLAG_FAST_GEOMETRY = 1
FLAG_COLD_START = 2
FLAG_COOP_LANES = 4


class Status:
    SOLVED, MAX_ITER, INFEASIBLE, DEGENERATE, DONE = range(5)


END_NAMES = ("stop_rule", "step_budget", "infeasible_k0_row", "infeasible_future_rows", "degenerate", "iteration_cap")


_lib = None


def lib():
    """Load libldcbf_b200.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                               "(nvcc, sm_100a).  ldcbf_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.ldcbf_abi_version.restype = c_int
        L.ldcbf_params_default.argtypes = [POINTER(LdcbfParams)]
        L.ldcbf_params_default.restype = None
        L.ldcbf_last_cuda_error.restype = c_char_p
        L.ldcbf_workspace_bytes.argtypes = [c_int] * 4
        L.ldcbf_workspace_bytes.restype = c_size_t
        P = c_void_p
        L.ldcbf_halfplanes_f64.argtypes = [c_int, c_int, c_int, P, P, P, P, P, P]
        L.ldcbf_mpc_qp_f64.argtypes = [POINTER(LdcbfParams), c_int, c_int, c_int] + [P] * 16
        L.ldcbf_mpc_step_f64.argtypes = [POINTER(LdcbfParams), c_int, c_int, c_int, c_int] + [P] * 19
        L.ldcbf_mpc_step_packed_f64.argtypes = [POINTER(LdcbfParams), c_int, c_int, c_int, c_int] + [P] * 11
        L.ldcbf_lidar_cast_f64.argtypes = [c_int, c_int, P, c_double, P, c_int, c_int, P, P, P, P, P, P, P]
        L.ldcbf_lidar_clusters_f64.argtypes = [c_int, c_int, P, P, c_double, c_int, c_int, c_int, P, P, P, P, P, P]
        L.ldcbf_clearance_grid_f64.argtypes = [c_int, c_int, c_int, c_int, c_int] + [P] * 10
        L.ldcbf_rollout_f64.argtypes = [POINTER(LdcbfParams)] + [c_int] * 7 + [P] * 17
        L.ldcbf_rollout_unknown_f64.argtypes = ([POINTER(LdcbfParams)] + [c_int] * 4 + [P, c_double, c_int, c_int] + [P] * 7
                                                + [c_double, c_int, c_int, c_int] + [P] * 10)
        L.ldcbf_probe_fp64_fma.argtypes = [c_int, c_int, c_int, P, P]
        for name in EXPORTS:
            getattr(L, name)
            if name not in ("ldcbf_params_default", "ldcbf_last_cuda_error", "ldcbf_workspace_bytes"):
                getattr(L, name).restype = c_int
        _lib = L
    return _lib


def abi_version():
    return lib().ldcbf_abi_version()


def default_params(sampling_time=None, **overrides):
    p = LdcbfParams()
    lib().ldcbf_params_default(ctypes.byref(p))
    if sampling_time is not None:
        p.sampling_time = float(sampling_time)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def params_from_conf(conf, sampling_time, **overrides):
    """ldcbf_params from the reference's `conf` dict (config.yml keys + BETA/OMEGA_* of HumanoidMpc.py:16-22)."""
    p = default_params(sampling_time)
    p.delta_t = float(conf["DELTA_T"]); p.gravity = float(conf["GRAVITY_CONST"]); p.com_height = float(conf["COM_HEIGHT"])
    p.alpha = float(conf["ALPHA"])
    p.l_max_x = float(conf["L_MAX_X"]); p.l_max_y = float(conf["L_MAX_Y"])
    p.l_min_x = float(conf["L_MIN_X"]); p.l_min_y = float(conf["L_MIN_Y"])
    p.v_min[0], p.v_min[1] = float(conf["V_MIN"][0]), float(conf["V_MIN"][1])
    p.v_max[0], p.v_max[1] = float(conf["V_MAX"][0]), float(conf["V_MAX"][1])
    p.omega_max = float(conf.get("OMEGA_MAX", 0.156 * math.pi))
    p.omega_min = float(conf.get("OMEGA_MIN", -p.omega_max))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def _check(rc, what):
    if rc != 0:
        err = lib().ldcbf_last_cuda_error().decode()
        raise RuntimeError(f"{what} failed with code {rc} (-1 argument, -2 unsupported shape, -3 CUDA launch: {err})")


def _ptr(t, dtype, name):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name}: expected a CUDA tensor (ldcbf_b200 has no CPU path)")
    if t.dtype != dtype or not t.is_contiguous():
        raise TypeError(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


F64, I32, I8, I64 = torch.float64, torch.int32, torch.int8, torch.int64


def half_planes(pos, verts, nverts, nobs, c_eta=None):
    """K1.  pos[B,2], verts[B,max_obs,max_verts,2], nverts[B,max_obs], nobs[B] -> c_eta[B,max_obs,4]."""
    B, max_obs, max_verts = verts.shape[0], verts.shape[1], verts.shape[2]
    if c_eta is None:
        c_eta = torch.empty((B, max_obs, 4), dtype=F64, device=verts.device)
    _check(lib().ldcbf_halfplanes_f64(B, max_obs, max_verts, _ptr(pos, F64, "pos"), _ptr(verts, F64, "verts"),
                                      _ptr(nverts, I32, "nverts"), _ptr(nobs, I32, "nobs"),
                                      _ptr(c_eta, F64, "c_eta"), _stream()), "ldcbf_halfplanes_f64")
    return c_eta


def _alloc_step_out(B, N, dev, out):
    out = {} if out is None else out
    def get(name, shape, dt):
        t = out.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dt, device=dev)
            out[name] = t
        return t
    get("U", (B, N, 2), F64); get("X", (B, N + 1, 4), F64); get("theta", (B, N + 1), F64); get("omega", (B, N), F64)
    get("obj", (B,), F64); get("status", (B,), I32); get("iters", (B,), I32)
    return out


def mpc_qp(prm, x0, theta0, goal, foot, c_eta, nobs, delta=None, limits=None, out=None):
    """K2+K3 given the half-planes.  Returns dict(U, X, theta, omega, obj, status, iters)."""
    B, N, max_obs = x0.shape[0], foot.shape[1] - 1, c_eta.shape[1]
    out = _alloc_step_out(B, N, x0.device, out)
    _check(lib().ldcbf_mpc_qp_f64(ctypes.byref(prm), B, N, max_obs, _ptr(x0, F64, "x0"), _ptr(theta0, F64, "theta0"),
                                  _ptr(goal, F64, "goal"), _ptr(foot, I8, "foot"), _ptr(c_eta, F64, "c_eta"),
                                  _ptr(nobs, I32, "nobs"), _ptr(delta, F64, "delta"), _ptr(limits, F64, "limits"),
                                  _ptr(out["U"], F64, "U"), _ptr(out["X"], F64, "X"), _ptr(out["theta"], F64, "theta"),
                                  _ptr(out["omega"], F64, "omega"), _ptr(out["obj"], F64, "obj"),
                                  _ptr(out["status"], I32, "status"), _ptr(out["iters"], I32, "iters"), _stream()),
           "ldcbf_mpc_qp_f64")
    return out


def mpc_step(prm, x0, theta0, goal, foot, verts, nverts, nobs, delta=None, warm=None, limits=None, out=None):
    """One full batched MPC step (K1 + K2+K3).  Returns dict(U, X, theta, omega, c_eta, obj, status, iters)."""
    B, N = x0.shape[0], foot.shape[1] - 1
    max_obs, max_verts = verts.shape[1], verts.shape[2]
    out = _alloc_step_out(B, N, x0.device, out)
    if out.get("c_eta") is None:
        out["c_eta"] = torch.empty((B, max_obs, 4), dtype=F64, device=x0.device)
    _check(lib().ldcbf_mpc_step_f64(ctypes.byref(prm), B, N, max_obs, max_verts, _ptr(x0, F64, "x0"),
                                    _ptr(theta0, F64, "theta0"), _ptr(goal, F64, "goal"), _ptr(foot, I8, "foot"),
                                    _ptr(verts, F64, "verts"), _ptr(nverts, I32, "nverts"), _ptr(nobs, I32, "nobs"),
                                    _ptr(delta, F64, "delta"), _ptr(warm, F64, "warm"), _ptr(limits, F64, "limits"),
                                    _ptr(out["U"], F64, "U"), _ptr(out["X"], F64, "X"),
                                    _ptr(out["theta"], F64, "theta"), _ptr(out["omega"], F64, "omega"),
                                    _ptr(out["c_eta"], F64, "c_eta"), _ptr(out["obj"], F64, "obj"),
                                    _ptr(out["status"], I32, "status"), _ptr(out["iters"], I32, "iters"), _stream()),
           "ldcbf_mpc_step_f64")
    return out


def mpc_step_packed(prm, state, goal, verts, nverts, nobs, N=3, delta=None, limits=None, out=None):
    """Loop-shaped step: state[B,6] = (p_x, v_x, p_y, v_y, theta, first stance foot +-1) in ->
    dict(next[B,10] = (x_next[4], theta_1, u0_x, u0_y, omega_0, objective, status), c_eta, iters)."""
    B = state.shape[0]
    max_obs, max_verts = verts.shape[1], verts.shape[2]
    out = {} if out is None else out
    dev = state.device
    for name, shape, dt in (("next", (B, 10), F64), ("c_eta", (B, max_obs, 4), F64), ("iters", (B,), I32)):
        if out.get(name) is None:
            out[name] = torch.empty(shape, dtype=dt, device=dev)
    _check(lib().ldcbf_mpc_step_packed_f64(ctypes.byref(prm), B, int(N), max_obs, max_verts, _ptr(state, F64, "state"),
                                           _ptr(goal, F64, "goal"), _ptr(verts, F64, "verts"),
                                           _ptr(nverts, I32, "nverts"), _ptr(nobs, I32, "nobs"),
                                           _ptr(delta, F64, "delta"), _ptr(limits, F64, "limits"),
                                           _ptr(out["next"], F64, "next"), _ptr(out["c_eta"], F64, "c_eta"),
                                           _ptr(out["iters"], I32, "iters"), _stream()), "ldcbf_mpc_step_packed_f64")
    return out


def ray_table(lidar_range, resolution=360):
    """[R,2] host table lidar_range*(cos, sin)(i*2*pi/R), evaluated with libm exactly as the reference does
    (RangeFinder/range_finder_wth_polygons_dbscan.py:28-37)."""
    step = 2 * math.pi / resolution
    return torch.tensor([[lidar_range * math.cos(i * step), lidar_range * math.sin(i * step)]
                         for i in range(resolution)], dtype=F64)


def lidar_cast(pos, verts, nverts, nobs, lidar_range, resolution=360, rays=None):
    """K4.  pos[B,2] -> hit_obs[B,R], hit_edge[B,R] (int32, -1 = none), hit_xy[B,R,2] (NaN = none)."""
    B, max_obs, max_verts = verts.shape[0], verts.shape[1], verts.shape[2]
    if rays is None:
        rays = ray_table(lidar_range, resolution).to(pos.device)
    R = rays.shape[0]
    hit_obs = torch.empty((B, R), dtype=I32, device=pos.device)
    hit_edge = torch.empty((B, R), dtype=I32, device=pos.device)
    hit_xy = torch.empty((B, R, 2), dtype=F64, device=pos.device)
    CH = 65535
    for s in range(0, B, CH):
        e = min(B, s + CH)
        _check(lib().ldcbf_lidar_cast_f64(e - s, R, _ptr(rays, F64, "rays"), float(lidar_range),
                                          _ptr(pos[s:e], F64, "pos"), max_obs, max_verts, _ptr(verts[s:e], F64, "verts"),
                                          _ptr(nverts[s:e], I32, "nverts"), _ptr(nobs[s:e], I32, "nobs"),
                                          _ptr(hit_obs[s:e], I32, "hit_obs"), _ptr(hit_edge[s:e], I32, "hit_edge"),
                                          _ptr(hit_xy[s:e], F64, "hit_xy"), _stream()), "ldcbf_lidar_cast_f64")
    return hit_obs, hit_edge, hit_xy


def clearance_grid(goal, verts, nverts, nobs, width=250, h_cap=None, with_cost=True, out=None, check=True):
    """f3: occupancy grid + exact distance transform + clearance cost of the planner front-end
    (HumanoidMPCWithRRT.py:21-88,103-108) for a batch of maps.
    Returns dict(meta[B,6] = (min_x, min_y, max_x, max_y, height, occupied cells), og[B,width+1,h_cap+1] uint8,
    dist, cost[B,width+1,h_cap+1] fp64); cells with j > height are zero in a fresh result.  `h_cap` defaults to 2*width
    (a frame at most twice as tall as wide); a map that needs more raises.
    Hot loops: pass the previous result as `out` (no allocation, no zero fill: cells with j > height then keep whatever
    an earlier, taller map left there - meta[:, 4] says where the grid ends) and `check=False` (the h_cap check reads a
    flag back from the device, i.e. synchronises; meta[:, 5] < 0 marks a map that did not fit)."""
    B = goal.shape[0]
    h_cap = int(h_cap) if h_cap is not None else 2 * int(width)
    dev = goal.device
    shape = (B, int(width) + 1, h_cap + 1)
    if out is None or out["og"].shape != shape or (with_cost and out.get("cost") is None):
        out = dict(meta=torch.zeros((B, 6), dtype=F64, device=dev), og=torch.zeros(shape, dtype=torch.uint8, device=dev),
                   dist=torch.zeros(shape, dtype=F64, device=dev),
                   cost=torch.zeros(shape, dtype=F64, device=dev) if with_cost else None,
                   work=torch.empty(shape, dtype=I32, device=dev))
    _check(lib().ldcbf_clearance_grid_f64(B, int(width), h_cap, verts.shape[1], verts.shape[2], _ptr(goal, F64, "goal"),
                                          _ptr(verts, F64, "verts"), _ptr(nverts, I32, "nverts"),
                                          _ptr(nobs, I32, "nobs"), _ptr(out["meta"], F64, "meta"),
                                          _ptr(out["og"], torch.uint8, "og"), _ptr(out["dist"], F64, "dist"),
                                          _ptr(out["cost"] if with_cost else None, F64, "cost"),
                                          _ptr(out["work"], I32, "work"), _stream()),
           "ldcbf_clearance_grid_f64")
    if check and bool((out["meta"][:, 5] < 0).any()):
        raise ValueError("ldcbf_clearance_grid_f64: a map needs more than h_cap grid rows (or has no obstacle)")
    return out


def lidar_clusters(hit_xy, noise=None, eps=0.3, min_samples=3, max_hulls=MAX_OBSTACLES, max_hull_verts=16):
    """f1.  hit_xy[B,R,2] (NaN = no reading) -> dict(labels[B,R], verts[B,max_hulls,max_hull_verts,2], nverts, nobs,
    overflow[B]); verts/nverts/nobs feed `half_planes` / `mpc_step` directly."""
    B, R = hit_xy.shape[0], hit_xy.shape[1]
    dev = hit_xy.device
    labels = torch.empty((B, R), dtype=I32, device=dev)
    verts = torch.empty((B, max_hulls, max_hull_verts, 2), dtype=F64, device=dev)
    nverts = torch.empty((B, max_hulls), dtype=I32, device=dev)
    nobs = torch.empty((B,), dtype=I32, device=dev)
    overflow = torch.empty((B,), dtype=I32, device=dev)
    _check(lib().ldcbf_lidar_clusters_f64(B, R, _ptr(hit_xy, F64, "hit_xy"), _ptr(noise, F64, "noise"), float(eps),
                                          int(min_samples), max_hulls, max_hull_verts, _ptr(labels, I32, "labels"),
                                          _ptr(verts, F64, "verts"), _ptr(nverts, I32, "nverts"), _ptr(nobs, I32, "nobs"),
                                          _ptr(overflow, I32, "overflow"), _stream()), "ldcbf_lidar_clusters_f64")
    return dict(labels=labels, verts=verts, nverts=nverts, nobs=nobs, overflow=overflow)


def rollout(prm, state, goals, right_first, verts, nverts, nobs, T, N=3, max_steps_per_goal=None, delta=None,
            limits=None, record=True):
    """Closed loop in one launch.  state[B,5] is updated in place.  goals[B,n_goals,2].
    Returns dict(traj_X[B,T+1,5], traj_U[B,T,3], steps[B], goal_steps[B,n_goals], status[B], total_solves,
    total_iters, end_code[B] (END_* below))."""
    B, n_goals = state.shape[0], goals.shape[1]
    dev = state.device
    max_obs, max_verts = verts.shape[1], verts.shape[2]
    max_steps_per_goal = T if max_steps_per_goal is None else max_steps_per_goal
    tX = torch.zeros((B, T + 1, 5), dtype=F64, device=dev) if record else None
    tU = torch.zeros((B, T, 3), dtype=F64, device=dev) if record else None
    steps = torch.empty((B,), dtype=I32, device=dev)
    goal_steps = torch.empty((B, n_goals), dtype=I32, device=dev)
    status = torch.empty((B,), dtype=I32, device=dev)
    total = torch.zeros((2,), dtype=I64, device=dev)          # (solves, active-set iterations)
    end_code = torch.empty((B,), dtype=I32, device=dev)
    _check(lib().ldcbf_rollout_f64(ctypes.byref(prm), B, N, T, n_goals, max_steps_per_goal, max_obs, max_verts,
                                   _ptr(state, F64, "state"), _ptr(goals, F64, "goals"),
                                   _ptr(right_first, I8, "right_first"), _ptr(verts, F64, "verts"),
                                   _ptr(nverts, I32, "nverts"), _ptr(nobs, I32, "nobs"), _ptr(delta, F64, "delta"),
                                   _ptr(limits, F64, "limits"), _ptr(tX, F64, "traj_X"), _ptr(tU, F64, "traj_U"),
                                   _ptr(steps, I32, "steps"), _ptr(goal_steps, I32, "goal_steps"),
                                   _ptr(status, I32, "status"), total.data_ptr(), total.data_ptr() + 8,
                                   _ptr(end_code, I32, "end_code"), _stream()),
           "ldcbf_rollout_f64")
    return dict(traj_X=tX, traj_U=tU, steps=steps, goal_steps=goal_steps, status=status, total_solves=total[:1],
                total_iters=total[1:], end_code=end_code)


def rollout_unknown(prm, state, goal, right_first, verts, nverts, nobs, T, rays, lidar_range, N=3, noise=None, eps=0.3,
                    min_samples=3, max_hulls=MAX_OBSTACLES, max_hull_verts=64, delta=None, limits=None, record=True):
    """Closed loop of the unknown-environment variant in one call (K4 -> f1 -> K1 -> K2+K3 -> advance per step, all on
    the device).  state[B,5] is updated in place.  Returns dict(traj_X, traj_U, steps, status, end_code, overflow,
    total_solves)."""
    B, dev = state.shape[0], state.device
    max_obs, max_verts, R = verts.shape[1], verts.shape[2], rays.shape[0]
    tX = torch.zeros((B, T + 1, 5), dtype=F64, device=dev) if record else None
    tU = torch.zeros((B, T, 3), dtype=F64, device=dev) if record else None
    steps = torch.empty((B,), dtype=I32, device=dev)
    status = torch.empty((B,), dtype=I32, device=dev)
    end_code = torch.empty((B,), dtype=I32, device=dev)
    overflow = torch.empty((B,), dtype=I32, device=dev)
    total = torch.zeros((1,), dtype=I64, device=dev)
    _check(lib().ldcbf_rollout_unknown_f64(ctypes.byref(prm), B, int(N), int(T), R, _ptr(rays, F64, "rays"),
                                           float(lidar_range), max_obs, max_verts, _ptr(state, F64, "state"),
                                           _ptr(goal, F64, "goal"), _ptr(right_first, I8, "right_first"),
                                           _ptr(verts, F64, "verts"), _ptr(nverts, I32, "nverts"), _ptr(nobs, I32, "nobs"),
                                           _ptr(noise, F64, "noise"), float(eps), int(min_samples), int(max_hulls),
                                           int(max_hull_verts), _ptr(delta, F64, "delta"), _ptr(limits, F64, "limits"),
                                           _ptr(tX, F64, "traj_X"), _ptr(tU, F64, "traj_U"), _ptr(steps, I32, "steps"),
                                           _ptr(status, I32, "status"), _ptr(end_code, I32, "end_code"),
                                           _ptr(overflow, I32, "overflow"), _ptr(total, I64, "total_solves"), _stream()),
           "ldcbf_rollout_unknown_f64")
    return dict(traj_X=tX, traj_U=tU, steps=steps, status=status, end_code=end_code, overflow=overflow, total_solves=total)


def probe_fp64(blocks=148 * 8, threads=256, iters=20000):
    """Run the FP64 FMA-chain probe; returns achieved TFLOP/s (2 flop per FMA)."""
    out = torch.empty((blocks * threads,), dtype=F64, device="cuda")
    L = lib()
    _check(L.ldcbf_probe_fp64_fma(blocks, threads, 200, _ptr(out, F64, "out"), _stream()), "probe")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(L.ldcbf_probe_fp64_fma(blocks, threads, iters, _ptr(out, F64, "out"), _stream()), "probe")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return (2.0 * 8 * iters * blocks * threads) / (ms * 1e-3) / 1e12
