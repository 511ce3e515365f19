import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[R, os.path.join(R,"humanoid-navigation-using-mpc-ldcbf_b200")]
import numpy as np, torch
from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC
from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils
from oracle import mpc, model
hulls=[ObstaclesUtils.generate_circle_like_polygon(10, 0.5, (5.5, -1.2)),ObstaclesUtils.generate_circle_like_polygon(20, 1, (4, 2)),ObstaclesUtils.generate_circle_like_polygon(25, 1.2, (1.7, 0))]
rings=[h.points[h.vertices] for h in hulls]
m=HumanoidMPC(N_horizon=3,N_mpc_timesteps=40,sampling_time=0.1,goal=(6,-3),init_state=(0,0,3,0,0),obstacles=hulls,verbosity=0)
X,U,_=m.run_simulation(None)
s_v=model.foot_parity(100)
bad=[]
for k in range(0,U.shape[1],4):
    r=mpc.mpc_step(X[:,k],(6,-3),rings,s_v[k//4:k//4+4],sampling_time=0.1)
    sol=m._solve(X[:,k],s_v[k//4:k//4+4],*m._get_list_c_and_eta(X[0,k],X[2,k]))
    d=np.abs(r['x_next']-X[:,k+1]).max() if r['status']==0 else np.nan
    d2=np.abs(sol['X'][1]-X[:4,k+1]).max()
    print(k,"oracle-vs-rollout %.2e stepkernel-vs-rollout %.2e"%(d,d2),"iters",sol['iters'],"status",sol['status'],r['status'],"obj gpu %.9f oracle %.9f"%(sol['obj'],r['obj']))
    if d>1e-6 or d2>1e-6: bad.append(k)
np.savez(os.path.join(R,"gpurun_out","substep_dump.npz"),X=X,U=U,bad=np.array(bad))
