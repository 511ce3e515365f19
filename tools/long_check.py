"""Config-5 check of the long-horizon kernel against oracle/qp_pspace.py (development aid; the pytest version is
tests/test_gpu_long_horizon.py).  usage: long_check.py N n_obs B"""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "humanoid-navigation-using-mpc-ldcbf_b200")
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
from oracle import qp_pspace, model

N, n_obs, B = (int(a) for a in sys.argv[1:4])
ncheck = int(sys.argv[4]) if len(sys.argv) > 4 else min(B, 24)
sc = scenarios.config5(B, n_obs, seed=0)
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
foots = scenarios.foot_window(sc["right_first"], 0, N)  # horizon = foot.shape[1]-1
prm = L.default_params(0.4)
args = (prm, cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8), cu(sc["verts"]),
        cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32))
out = L.mpc_step(*args)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    out = L.mpc_step(*args)
torch.cuda.synchronize()
dt = (time.time() - t0) / 3
st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
print(f"N={N} n_obs={n_obs} B={B}: {dt*1e3:.2f} ms/step  {B/dt:.3e} solves/s  status counts {np.bincount(st, minlength=4)}  iters mean {it.mean():.1f} max {it.max()}")
conf = model.default_conf()
X = out["X"].cpu().numpy(); U = out["U"].cpu().numpy(); obj = out["obj"].cpu().numpy()
bad = 0
for b in range(ncheck):
    ref = qp_pspace.mpc_step(sc["state"][b], sc["goal"][b], sc["rings"][sc["map_index"][b]], foots[b], N, 0.4, conf)
    if ref["status"] != st[b]:
        print(b, "status", ref["status"], st[b]); bad += 1; continue
    if st[b] == 0:
        ex, eu, eo = np.abs(ref["X"] - X[b]).max(), np.abs(ref["U"] - U[b]).max(), abs(ref["obj"] - obj[b]) / ref["obj"]
        if ex > 1e-6 or eu > 1e-6 or eo > 1e-8:
            print(b, "err X %.2e U %.2e obj %.2e" % (ex, eu, eo)); bad += 1
print("checked", ncheck, "bad", bad)
