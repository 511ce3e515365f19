import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
B = 16384
c3 = scenarios.config3(B, seed=0)
cu = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
pos, v, nv, no = cu(c3["pos"], torch.float64), cu(c3["verts"], torch.float64), cu(c3["nverts"], torch.int32), cu(c3["nobs"], torch.int32)
rays = L.binding.ray_table(1.5, 360).cuda()
for _ in range(3): L.lidar_cast(pos, v, nv, no, 1.5, 360, rays=rays)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ho, he, xy = L.lidar_cast(pos, v, nv, no, 1.5, 360, rays=rays)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"lidar B={B}: {ms*1e3:.1f} us  {B/ms*1e3:.3e} scans/s  hit fraction {(ho>=0).float().mean().item():.3f}")
