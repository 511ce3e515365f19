import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch, ldcbf_b200 as L
from ldcbf_b200 import scenarios
c3=scenarios.config3(16384,seed=0); c=lambda a,dt: torch.as_tensor(np.ascontiguousarray(a),dtype=dt).cuda()
pos,v,nv,no=c(c3["pos"],torch.float64),c(c3["verts"],torch.float64),c(c3["nverts"],torch.int32),c(c3["nobs"],torch.int32)
ho,he,xy=L.lidar_cast(pos,v,nv,no,1.5,360)
noise=torch.randn((16384,360,2),dtype=torch.float64,device="cuda",generator=torch.Generator("cuda").manual_seed(0))*0.01
for nz in (None,noise):
    for _ in range(3): L.lidar_clusters(xy,noise=nz,max_hull_verts=64)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): o=L.lidar_clusters(xy,noise=nz,max_hull_verts=64)
    e1.record(); torch.cuda.synchronize(); print("f1 ms %.3f"%(e0.elapsed_time(e1)/10),"mean hulls",o["nobs"].double().mean().item(),"overflow",o["overflow"].sum().item())
