"""Development aid: does the ORDER of the scenarios in the batch (= which scenarios share a warp in the closed-loop
kernel) change the pass time?  The pass is the sum over steps of the slowest scenario of every warp."""
import os, sys, statistics, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
B = 4096
sc = scenarios.config2(B, seed=0)
def timeit(fn, n=6, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
def run(order, name):
    s = {k: (v[order] if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    eng = L.BatchedHumanoidMPC(s["goal"], s["verts"], s["nverts"], s["nobs"], N_horizon=3, sampling_time=0.4, delta=np.full(B, 1e-6))
    st0, rf = cu(s["state"]), cu(s["right_first"].astype(np.int8), torch.int8)
    res = {}
    def ro(): res["r"] = eng.rollout(st0.clone(), rf, 150, record=False)
    ms = timeit(ro)
    steps = res["r"]["steps"].cpu().numpy()
    print(json.dumps({"order": name, "ms": round(ms, 3), "solves": int(res["r"]["total_solves"].item())}), flush=True)
    return steps
ident = np.arange(B)
steps = run(ident, "as generated")
p, g = sc["state"][:, [0, 2]], sc["goal"]
dist = np.hypot(*(g - p).T)
run(np.argsort(dist), "by start-goal distance")
run(np.argsort(steps, kind="stable"), "by executed steps (oracle knowledge)")
run(np.argsort(-steps, kind="stable"), "by executed steps, longest first")
run(np.lexsort((p[:, 1], np.floor(p[:, 0] * 4))), "by start position (x strips, y)")
th = sc["state"][:, 4]
run(np.argsort(th), "by initial heading")
ang = np.arctan2(*(g - p).T[::-1])
run(np.argsort(ang), "by bearing of the goal")
# spread: the longest loops one per warp (16 scenarios per warp at this batch): round-robin over the sorted list
o = np.argsort(-steps, kind="stable"); nw = B // 16
spread = o.reshape(16, nw).T.reshape(-1)
run(spread, "longest loops spread one per warp (oracle knowledge)")
