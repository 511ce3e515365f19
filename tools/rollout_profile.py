"""Driver for the ncu capture of the headline kernel: config-2 closed loops (bench.py's pass), B = 4096, margin 1e-6."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
B = 4096
sc = scenarios.config2(B, seed=0)
eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4, delta=np.full(B, 1e-6))
st0 = torch.as_tensor(sc["state"]).cuda()
rf = torch.as_tensor(sc["right_first"].astype(np.int8)).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    r = eng.rollout(st0.clone(), rf, 150, record=False)
torch.cuda.synchronize()
print("ok", int(r["total_solves"].item()))
