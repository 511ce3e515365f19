"""Small driver for compute-sanitizer (memcheck / racecheck): every K2+K3 launch shape, the half-plane kernels and the
rollout kernel once, at sizes that stay quick under the tool.
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
from ldcbf_b200.binding import FLAG_COLD_START, FLAG_COOP_LANES

def inputs(B, seed=0):
    sc = scenarios.config2(min(B, 2048), seed=seed)
    rep = (B + len(sc["state"]) - 1) // len(sc["state"])
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B]), dtype=dt).cuda()
    ft = scenarios.foot_window(sc["right_first"], 0, 3)
    return sc, (cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(ft, torch.int8), cu(sc["verts"]),
                cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32))

for B, flags, what in ((301, 0, "race, 8-lane blocks"), (3001, 0, "race, 16-lane blocks"), (64, FLAG_COOP_LANES, "coop"),
                       (10001, 0, "plain 32"), (151552 + 5, 0, "prepare + resume"), (301, FLAG_COLD_START, "cold small")):
    sc, args = inputs(B)
    out = L.mpc_step(L.default_params(0.4, flags=flags), *args)
    torch.cuda.synchronize()
    print(what, B, torch.bincount(out["status"], minlength=4).tolist(), flush=True)
sc, args = inputs(200)
eng = L.BatchedHumanoidMPC(args[2], args[4], args[5], args[6], N_horizon=3, sampling_time=0.4)
st = torch.cat((args[0], args[1][:, None]), 1).contiguous()
r = eng.rollout(st, torch.ones(200, dtype=torch.int8, device="cuda"), 30, record=True)
torch.cuda.synchronize()
print("rollout (4 lanes per scenario)", int(r["total_solves"].item()), flush=True)
L.lib().ldcbf_trim_workspace()
print("ok")
