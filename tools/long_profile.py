"""Per-phase cycle split of the long-horizon kernel (needs the -DLDCBF_LONG_PROFILE build, see csrc/mpc_long.cu).
usage: LDCBF_B200_LIB=.../_prof_libldcbf.so python tools/long_profile.py N n_obs B"""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "humanoid-navigation-using-mpc-ldcbf_b200")
import ldcbf_b200 as L
from ldcbf_b200 import scenarios

N, n_obs, B = (int(a) for a in sys.argv[1:4])
sc = scenarios.config5(B, n_obs, seed=0)
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
foots = scenarios.foot_window(sc["right_first"], 0, N)
args = (L.default_params(0.4), cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8),
        cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32))
lib = L.lib()
buf = (ctypes.c_ulonglong * 12)()
L.mpc_step(*args)
lib.ldcbf_debug_long_profile(buf, 1)
out = L.mpc_step(*args)
lib.ldcbf_debug_long_profile(buf, 1)
it = out["iters"].cpu().numpy().astype(np.float64)
names = ["between", "setup", "velocities", "scan+pick", "d=J'n", "solve|z", "decide", "add(H)", "drop:shift", "drop:chain", "drop:Jrot", "tri-solve"]
tot = float(sum(buf))
print(f"N={N} n_obs={n_obs} B={B}: total iterations {it.sum():.0f}, cycles/iteration {tot / it.sum():.0f}")
for n_, v in zip(names, buf):
    print(f"  {n_:12s} {v / it.sum():9.0f} cycles/iter  {100 * v / tot:5.1f} %")
