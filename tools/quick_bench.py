"""Developer timing loop (not the judged bench): step time vs batch size on one GPU."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios

def main():
    print("fp64 probe TFLOP/s", [round(L.probe_fp64(), 2) for _ in range(2)])
    sc = scenarios.config2(4096, seed=0)
    for B in [int(a) for a in sys.argv[1:]] or (4096, 65536, 1048576):
        rep = max(1, B // 4096)
        cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(np.tile(a[:B], (rep,) + (1,) * (a.ndim - 1))), dtype=dt).cuda()
        x0, th, g = cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"])
        ft = cu(scenarios.foot_window(sc["right_first"], 0, 3), torch.int8)
        v, nv, no = cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32)
        prm = L.default_params(0.4)
        out = None
        for _ in range(5):
            out = L.mpc_step(prm, x0, th, g, ft, v, nv, no, out=out)
        torch.cuda.synchronize()
        n = 20
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(n):
            L.half_planes(x0, v, nv, no, c_eta=out["c_eta"]) if False else L.lib()
        for _ in range(n):
            out = L.mpc_step(prm, x0, th, g, ft, v, nv, no, out=out)
        ev[1].record()
        for _ in range(n):
            L.mpc_qp(prm, x0, th, g, ft, out["c_eta"], no, out=out)
        ev[2].record()
        torch.cuda.synchronize()
        ms, ms_qp = ev[0].elapsed_time(ev[1]) / n, ev[1].elapsed_time(ev[2]) / n
        it = out["iters"].float()
        print(f"B={x0.shape[0]} step {ms*1e3:.1f} us  qp-only {ms_qp*1e3:.1f} us  solves/s {x0.shape[0]/ms*1e3:.3e}  "
              f"iters mean {it.mean().item():.2f} max {int(it.max().item())} status {torch.bincount(out['status']).tolist()}")

if __name__ == "__main__":
    main()
