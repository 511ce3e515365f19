"""Development aid: time the config-2 closed-loop pass (bench.py's headline) for the rollout block sizes selectable with
LDCBF_ROLLOUT_BLOCK (one process per setting: the library reads the variable once)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os, statistics, json
sys.path.insert(0, os.path.join(%r, "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
B = int(os.environ.get("ROLL_B", "4096"))
sc = scenarios.config2(B, seed=0)
eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4, delta=np.full(B, 1e-6))
st0 = torch.as_tensor(sc["state"]).cuda(); rf = torch.as_tensor(sc["right_first"].astype(np.int8)).cuda()
for _ in range(3): r = eng.rollout(st0.clone(), rf, 150, record=False)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    st = st0.clone(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = eng.rollout(st, rf, 150, record=False); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print(json.dumps({"block": os.environ.get("LDCBF_ROLLOUT_BLOCK"), "G": os.environ.get("LDCBF_ROLLOUT_G"), "ms": statistics.median(ts), "solves": int(r["total_solves"].item()),
                  "steps_sum": int(r["steps"].sum().item()), "start": os.environ.get("LDCBF_ROLLOUT_START"),
                  "iters_mean": float(r["total_iters"].item()) / int(r["total_solves"].item())}))
''' % ROOT
for blk in sys.argv[1:] or ["32", "16", "8"]:
    env = (dict(os.environ, LDCBF_ROLLOUT_G=blk[1:]) if blk.startswith("g") else
           dict(os.environ, LDCBF_ROLLOUT_START=blk[1:]) if blk.startswith("s") else dict(os.environ, LDCBF_ROLLOUT_BLOCK=blk))
    print(subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
