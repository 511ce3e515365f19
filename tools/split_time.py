"""Developer probe: how much of the QP kernel is setup/finish vs active-set trips (max_iter sweep)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
sc = scenarios.config2(4096, seed=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rep = B // 4096
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))), dtype=dt).cuda()
x0, th, g = cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"])
ft = cu(scenarios.foot_window(sc["right_first"], 0, 3), torch.int8)
v, nv, no = cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32)
out = L.mpc_step(L.default_params(0.4), x0, th, g, ft, v, nv, no)
for mi in (0, 1, 2, 4, 8, 16, 32, 200):
    prm = L.default_params(0.4, max_iter=mi)
    for _ in range(3): L.mpc_qp(prm, x0, th, g, ft, out["c_eta"], no, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.mpc_qp(prm, x0, th, g, ft, out["c_eta"], no, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"max_iter {mi:4d}: {e0.elapsed_time(e1)/10*1e3:8.1f} us   mean iters {out['iters'].float().mean().item():.2f}")
