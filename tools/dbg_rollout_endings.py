"""Development aid: closed loops of the bench batch on the GPU; for every loop that ended on a failed solve, ask the
oracle and the open-loop step kernel about the same final state and save the disagreements for offline analysis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
from oracle import model, mpc

B, T, D = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 150, 1e-6
sc = scenarios.config2(B, seed=0)
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4, delta=np.full(B, D))
r = eng.rollout(cu(sc["state"]), cu(sc["right_first"].astype(np.int8), torch.int8), T)
tX, steps, status, end = (r[k].cpu().numpy() for k in ("traj_X", "steps", "status", "end_code"))
out = []
for b in np.flatnonzero(status != 0):
    k = int(steps[b])
    s_v = model.foot_parity(T + 8, bool(sc["right_first"][b]))
    foot = np.array(s_v[k:k + 4], dtype=np.int8)
    o = mpc.mpc_step(tX[b, k], sc["goal"][b], sc["rings"][b], [int(v) for v in foot], sampling_time=0.4, delta=D)
    g = L.mpc_step(L.default_params(0.4), cu(tX[b:b + 1, k, :4]), cu(tX[b:b + 1, k, 4]), cu(sc["goal"][b:b + 1]), cu(foot[None], torch.int8),
                   cu(sc["verts"][b:b + 1]), cu(sc["nverts"][b:b + 1], torch.int32), cu(sc["nobs"][b:b + 1], torch.int32), delta=cu([D]))
    gs = int(g["status"].item())
    if o["status"] != status[b] or gs != status[b]:
        print("scenario", b, "step", k, "rollout", status[b], "end", end[b], "open-loop kernel", gs, "oracle", o["status"], flush=True)
        out.append(np.concatenate(([b, k, status[b], gs, o["status"]], tX[b, k], tX[b, max(k - 1, 0)], sc["goal"][b], foot,
                                   g["c_eta"].cpu().numpy().ravel())))
print("failed loops", int((status != 0).sum()), "disagreements", len(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "dbg_endings.npy"), np.array(out))
