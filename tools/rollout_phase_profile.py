import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/humanoid-navigation-using-mpc-ldcbf_b200"]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
sc = scenarios.config2(4096, seed=0)
for B in (1, 64, 2048, 4096):
    eng = L.BatchedHumanoidMPC(sc["goal"][:B], sc["verts"][:B], sc["nverts"][:B], sc["nobs"][:B], N_horizon=3, sampling_time=0.4, delta=np.full(B, 1e-6))
    st0 = torch.as_tensor(sc["state"][:B]).cuda(); rf = torch.as_tensor(sc["right_first"][:B].astype(np.int8)).cuda()
    for _ in range(2):
        r = eng.rollout(st0.clone(), rf, 150, record=False); torch.cuda.synchronize()
    print("B", B, flush=True)
