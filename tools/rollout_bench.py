"""Developer timing of the closed-loop rollout kernel (config 2, 4096 scenarios, <= 150 steps), warm vs cold start."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch, ldcbf_b200 as L
from ldcbf_b200 import scenarios
from ldcbf_b200.binding import FLAG_COLD_START
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = scenarios.config2(4096, seed=0); rep = max(1, B // 4096)
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))), dtype=dt).cuda()
for flags, name in ((0, "warm"), (FLAG_COLD_START, "cold")):
    for delta in (0.0, 1e-6):
        eng = L.BatchedHumanoidMPC(cu(sc["goal"]), cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32), N_horizon=3,
                                   sampling_time=0.4, delta=None if not delta else torch.full((B,), delta, dtype=torch.float64, device="cuda"), flags=flags)
        rf = cu(sc["right_first"].astype(np.int8), torch.int8); st0 = cu(sc["state"])
        for _ in range(2): r = eng.rollout(st0.clone(), rf, 150, record=False)
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            st = st0.clone(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = eng.rollout(st, rf, 150, record=False); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]; solves = int(r["total_solves"].item())
        print(f"{name} delta={delta:g}: {ms:.2f} ms, {solves} solves, {solves/ms*1e3:.3e} solves/s, stop-rule endings {(r['status']==0).sum().item()}, final-state checksum {st.sum().item():.9f}")
