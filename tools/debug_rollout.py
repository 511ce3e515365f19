import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[R, os.path.join(R,"humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC, conf
from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils
from oracle import mpc, model
hulls=[ObstaclesUtils.generate_circle_like_polygon(10, 0.5, (5.5, -1.2)),ObstaclesUtils.generate_circle_like_polygon(20, 1, (4, 2)),ObstaclesUtils.generate_circle_like_polygon(25, 1.2, (1.7, 0))]
m=HumanoidMPC(N_horizon=3,N_mpc_timesteps=300,sampling_time=0.4,goal=(6,-3),init_state=(0,0,3,0,0),obstacles=hulls,verbosity=1)
X,U,_=m.run_simulation(None)
print("fused",X.shape,m.last_status)
Xs,Us=m._run_stepwise(); print("stepwise",Xs.shape,m.last_status)
rings=[h.points[h.vertices] for h in hulls]
Xo,Uo=mpc.run_simulation((6,-3),rings,(0,0,3,0,0),3,300,0.4)
n=min(X.shape[1],Xo.shape[1]); print("max diff fused vs oracle", np.abs(X[:,:n]-Xo[:,:n]).max(0)[-5:])
k=X.shape[1]-1
s_v=model.foot_parity(400)
r=mpc.mpc_step(X[:,k],(6,-3),rings,s_v[k:k+4],sampling_time=0.4); print("oracle at last fused state: status",r['status'],r['obj'])
sol=m._solve(X[:,k],s_v[k:k+4],*m._get_list_c_and_eta(X[0,k],X[2,k])); print("gpu step:",sol['status'],sol['iters'],sol['obj'])
print("---- per-step check of fused trajectory")
for k in range(X.shape[1]-1):
    r=mpc.mpc_step(X[:,k],(6,-3),rings,s_v[k:k+4],sampling_time=0.4)
    sol=m._solve(X[:,k],s_v[k:k+4],*m._get_list_c_and_eta(X[0,k],X[2,k]))
    d=np.abs(r['x_next']-X[:,k+1]).max() if r['status']==0 else np.nan
    d2=np.abs(sol['X'][1]-X[:4,k+1]).max()
    print(k, "oracle-vs-fused %.2e"%d, "stepkernel-vs-fused %.2e"%d2, "iters",sol['iters'], "obj gpu %.10f oracle %.10f"%(sol['obj'],r['obj']), "kkt",["%.1e"%v for v in r['sol']['kkt']] if r['status']==0 else None)
