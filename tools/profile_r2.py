"""Driver for the round-2 ncu captures committed under profiles/ (run it under ncu, see profiles/README.md):
one launch of every hot-path kernel after its warm-up — the headline closed-loop pass (rollout_kernel, B = 4096), the
open-loop step at B = 4096 and B = 2^20 (K1 + K2+K3), the LiDAR caster and the clustering on 16384 config-3 scans, the
unknown-environment loop kernels and the FP64 probe that supplies the roofline denominator."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
REP = int(os.environ.get("PROFILE_REP", "3"))
sc = scenarios.config2(4096, seed=0)
foots = scenarios.foot_window(sc["right_first"], 0, 3)
prm = L.default_params(0.4)
# headline: closed loops
eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=3, sampling_time=0.4, delta=np.full(4096, 1e-6))
st0, rf = cu(sc["state"]), cu(sc["right_first"].astype(np.int8), torch.int8)
for _ in range(REP):
    eng.rollout(st0.clone(), rf, 150, record=False)
# open-loop step, B = 4096 and B = 2^20
a = (cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8), cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32))
dl = cu(np.full(4096, 1e-6))
for _ in range(REP):
    L.mpc_step(prm, *a, delta=dl)
rep = (1 << 20) // 4096
t = lambda x, dt=torch.float64: cu(np.tile(x, (rep,) + (1,) * (x.ndim - 1)), dt)
big = (t(sc["state"][:, :4]), t(sc["state"][:, 4]), t(sc["goal"]), t(foots, torch.int8), t(sc["verts"]), t(sc["nverts"], torch.int32), t(sc["nobs"], torch.int32))
o = None
for _ in range(REP):
    o = L.mpc_step(prm, *big, out=o)
del big, o
torch.cuda.empty_cache()
# config 3: caster, clustering, the loop
B3 = 16384
c3 = scenarios.config3(B3, seed=0)
ue = L.BatchedUnknownEnvMPC(c3["goal"], c3["verts"], c3["nverts"], c3["nobs"], lidar_range=1.5, sampling_time=0.4, N_horizon=3, delta=np.full(B3, 1e-6))
f3 = scenarios.foot_window(np.ones(B3, bool), 0, 3)
x0, th, ft = cu(c3["state"][:, :4]), cu(c3["state"][:, 4]), cu(f3, torch.int8)
noise = torch.randn((B3, 360, 2), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0)) * 0.01
for _ in range(REP):
    ue.step(x0, th, ft, noise=noise)
s0 = np.zeros((B3, 5)); s0[:, 4] = np.pi / 2
ue.rollout(cu(s0), cu(np.ones(B3, np.int8), torch.int8), 3, noise=noise, record=False)
# FP64 probe (roofline denominator)
for _ in range(REP):
    L.probe_fp64()
torch.cuda.synchronize()
print("ok")
