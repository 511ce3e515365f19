"""Development aid: every transition of closed loops replayed through the batched K1 + K2+K3 step; LDCBF_ROLLOUT_NOPRUNE=1
runs the rollout with the full ring walk."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
B, T, M = 96, 150, 1e-6
sc = scenarios.config2(B, seed=17)
eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4, delta=np.full(B, M))
r = eng.rollout(cu(sc["state"]), cu(sc["right_first"].astype(np.int8), torch.int8), T)
tX, steps = r["traj_X"].cpu().numpy(), r["steps"].cpu().numpy()
np.save(os.path.join(ROOT, "gpurun_out", "prune_ab_%s.npy" % ("off" if os.environ.get("LDCBF_ROLLOUT_NOPRUNE") else "on")), tX)
worst, where = 0.0, None
for k in range(int(steps.max())):
    alive = np.flatnonzero(steps > k)
    foots = scenarios.foot_window(sc["right_first"][alive], k, 3)
    o = L.mpc_step(L.default_params(0.4), cu(tX[alive, k, :4]), cu(tX[alive, k, 4]), cu(sc["goal"][alive]), cu(foots, torch.int8),
                   cu(sc["verts"][alive]), cu(sc["nverts"][alive], torch.int32), cu(sc["nobs"][alive], torch.int32), delta=cu(np.full(len(alive), M)))
    nxt = np.column_stack((o["X"][:, 1].cpu().numpy(), o["theta"][:, 1].cpu().numpy()))
    d = np.abs(nxt - tX[alive, k + 1]).max(1)
    if np.nanmax(d) > worst: worst, where = float(np.nanmax(d)), (int(alive[np.nanargmax(d)]), k)
print("prune", "off" if os.environ.get("LDCBF_ROLLOUT_NOPRUNE") else "on", "worst", worst, "at", where, "steps sum", int(steps.sum()))
