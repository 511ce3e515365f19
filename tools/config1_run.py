import sys, os, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/humanoid-navigation-using-mpc-ldcbf_b200")
from HumanoidNavigation.MPC.HumanoidMpc import HumanoidMPC, conf
from HumanoidNavigation.MPC.HumanoidMPCVariants.HumanoidMPCCustomLCBF import HumanoidMPCCustomLCBF
from tests.test_gpu_host_mirror import _hulls
hulls = _hulls()
for delta in (0.0, 1e-6, 0.3):
    if delta == 0.0:
        m = HumanoidMPC(N_horizon=3, N_mpc_timesteps=300, sampling_time=conf['DELTA_T'], goal=(6, -3), init_state=(0, 0, 3, 0, 0), obstacles=hulls, verbosity=0)
    else:
        m = HumanoidMPCCustomLCBF(N_horizon=3, N_mpc_timesteps=300, sampling_time=conf['DELTA_T'], goal=(6, -3), init_state=np.array([0, 0, 3, 0, 0.0]), obstacles=hulls, verbosity=0, distance_from_obstacles=delta)
    X, U, _ = m.run_simulation(path_to_gif=None, make_fast_plot=False, plot_animation=False, fill_animator=False)
    print("delta", delta, "steps", U.shape[1], "status", m.last_status, "final dist %.3f" % np.hypot(X[0, -1] - 6, X[2, -1] + 3))
