"""Development aid: time of the closed-loop pass against the batch size (same scenarios, prefixes of config2(seed=0))."""
import os, sys, statistics, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
sc = scenarios.config2(4096, seed=0)
for B in (1, 8, 64, 512, 1024, 2048, 4096):
    eng = L.BatchedHumanoidMPC(sc["goal"][:B], sc["verts"][:B], sc["nverts"][:B], sc["nobs"][:B], N_horizon=3, sampling_time=0.4, delta=np.full(B, 1e-6))
    st0 = torch.as_tensor(sc["state"][:B]).cuda(); rf = torch.as_tensor(sc["right_first"][:B].astype(np.int8)).cuda()
    for _ in range(3): r = eng.rollout(st0.clone(), rf, 150, record=False)
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        st = st0.clone(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = eng.rollout(st, rf, 150, record=False); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts); mx = int(r["steps"].max().item())
    print(json.dumps({"B": B, "ms": round(ms, 3), "max_steps": mx, "us_per_step_of_longest": round(1e3 * ms / mx, 2),
                      "solves": int(r["total_solves"].item()), "iters_mean": round(float(r["total_iters"].item()) / int(r["total_solves"].item()), 2)}), flush=True)
