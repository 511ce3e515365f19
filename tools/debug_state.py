import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[R, os.path.join(R,"humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios
from oracle import mpc, model, halfplane
from tests import helpers
geo=helpers.load_geo(); rings=helpers.map_rings(geo,'circles')
state=np.array([ 2.13893923,  0.40524595,  1.10594404,  0.01519288, -0.80080179]); ft=np.array([[1,-1,1,-1]],dtype=np.int8)
d=np.load(os.path.join(R,"gpurun_out","substep_dump.npz")) if os.path.exists(os.path.join(R,"gpurun_out","substep_dump.npz")) else None
if d is not None: state=d['X'][:,120]
verts,nverts,nobs=scenarios.pack_rings([rings])
cu=lambda a,dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a),dtype=dt).cuda()
for Ts in (0.1,0.4):
    prm=L.default_params(Ts)
    out=L.mpc_step(prm,cu(state[None,:4]),cu(state[None,4]),cu([[6.,-3.]]),cu(ft,torch.int8),cu(verts),cu(nverts,torch.int32),cu(nobs,torch.int32))
    o={k:v.cpu().numpy()[0] for k,v in out.items()}
    r=mpc.mpc_step(state,(6,-3),rings,[1,-1,1,-1],sampling_time=Ts)
    qp=r['qp']; z=o['U'].ravel(); res=qp['A']@z; v=np.maximum(res-qp['hi'],qp['lo']-res)
    print("Ts",Ts,"gpu obj",o['obj'],"oracle",r['obj'],"iters",o['iters'],"status",o['status'])
    print("  gpu viol rows:",[(qp['kinds'][i],"%.3e"%v[i]) for i in range(len(v)) if v[i]>1e-9])
    print("  theta diff",np.abs(o['theta']-r['theta']).max(),"c_eta diff",np.abs(o['c_eta'][:,:2]-r['c']).max(),np.abs(o['c_eta'][:,2:]-r['eta']).max())
    print("  gpu U",o['U'].ravel(),"\n  oracle U",r['sol']['z'])
