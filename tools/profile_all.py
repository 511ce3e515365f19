"""Small driver for the ncu captures committed under profiles/: one step at B=2^18, one LiDAR scan batch + clustering."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
B = 1 << 18
sc = scenarios.config2(4096, seed=0)
rep = B // 4096
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))), dtype=dt).cuda()
x0, th, g = cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"])
ft = cu(scenarios.foot_window(sc["right_first"], 0, 3), torch.int8)
v, nv, no = cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32)
prm = L.default_params(0.4)
out = None
for _ in range(3):
    out = L.mpc_step(prm, x0, th, g, ft, v, nv, no, out=out)
# small batch step (B = 4096: the bench workload)
s = slice(0, 4096)
for _ in range(3):
    L.mpc_step(prm, x0[s].contiguous(), th[s].contiguous(), g[s].contiguous(), ft[s].contiguous(), v[s].contiguous(), nv[s].contiguous(), no[s].contiguous())
c3 = scenarios.config3(16384, seed=0)
c = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
pos, v3, nv3, no3 = c(c3["pos"], torch.float64), c(c3["verts"], torch.float64), c(c3["nverts"], torch.int32), c(c3["nobs"], torch.int32)
rays = L.binding.ray_table(1.5, 360).cuda()
for _ in range(3):
    ho, he, xy = L.lidar_cast(pos, v3, nv3, no3, 1.5, 360, rays=rays)
    cl = L.lidar_clusters(xy)
torch.cuda.synchronize()
print("ok")
