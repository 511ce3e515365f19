"""Development aid: the handful of timings iterated on in round 2 (one process; LDCBF_PREP_TRIPS etc. are read once)."""
import os, sys, statistics, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
what = sys.argv[1:] or ["rollout", "f1", "large"]
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("LDCBF_")}}
sc = scenarios.config2(4096, seed=0)
if "rollout" in what:
    eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4, delta=np.full(4096, 1e-6))
    st0, rf = cu(sc["state"]), cu(sc["right_first"].astype(np.int8), torch.int8)
    res = {}
    def ro(): res["r"] = eng.rollout(st0.clone(), rf, 150, record=False)
    out["rollout_ms"] = timeit(ro); out["rollout_solves"] = int(res["r"]["total_solves"].item())
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    a = (cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8), cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32))
    o = {}
    out["step0_us"] = 1e3 * timeit(lambda: L.mpc_step(L.default_params(0.4), *a, out=o), n=50)
if "f1" in what:
    B = 16384
    c3 = scenarios.config3(B, seed=0)
    eng = L.BatchedUnknownEnvMPC(c3["goal"], c3["verts"], c3["nverts"], c3["nobs"], lidar_range=1.5, sampling_time=0.4, N_horizon=3)
    foots = scenarios.foot_window(np.ones(B, bool), 0, 3)
    x0, th, ft = cu(c3["state"][:, :4]), cu(c3["state"][:, 4]), cu(foots, torch.int8)
    noise = torch.randn((B, 360, 2), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0)) * 0.01
    o = eng.step(x0, th, ft, noise=noise)
    xy = o["sensed"]["hit_xy"]
    out["f1_noisy_ms"] = timeit(lambda: L.lidar_clusters(xy, noise=noise))
    out["f1_clean_ms"] = timeit(lambda: L.lidar_clusters(xy))
    out["unknown_step_ms"] = timeit(lambda: eng.step(x0, th, ft, noise=noise))
    pos = x0[:, [0, 2]].contiguous()
    out["k4_ms"] = timeit(lambda: L.lidar_cast(pos, eng.verts, eng.nverts, eng.nobs, 1.5, rays=eng.rays))
if "large" in what:
    rep = (1 << 20) // 4096
    foots = scenarios.foot_window(sc["right_first"], 0, 3)
    t = lambda a, dt=torch.float64: cu(np.tile(a, (rep,) + (1,) * (a.ndim - 1)), dt)
    a = (t(sc["state"][:, :4]), t(sc["state"][:, 4]), t(sc["goal"]), t(foots, torch.int8), t(sc["verts"]), t(sc["nverts"], torch.int32), t(sc["nobs"], torch.int32))
    o = L.mpc_step(L.default_params(0.4), *a)
    ce = o["c_eta"]
    out["large_qp_ms"] = timeit(lambda: L.mpc_qp(L.default_params(0.4), a[0], a[1], a[2], a[3], ce, a[6], out=o))
    out["large_step_ms"] = timeit(lambda: L.mpc_step(L.default_params(0.4), *a, out=o))
    out["large_iters_mean"] = float(o["iters"].double().mean().item())
    small = L.mpc_step(L.default_params(0.4), *[x[:4096].contiguous() for x in a])
    out["large_equals_small"] = bool(torch.equal(small["U"], o["U"][:4096])) and bool(torch.equal(small["status"], o["status"][:4096]))
print(json.dumps(out), flush=True)
