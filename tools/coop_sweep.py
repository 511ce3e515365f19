"""Developer sweep: K2+K3 kernel time vs batch for the thread-per-scenario and the cooperative (G lanes per scenario)
kernels.  Run once per setting: LDCBF_COOP_MAX_B=0 (thread kernels) or LDCBF_COOP_MAX_B=<big> LDCBF_COOP_G=8|16|32."""
import os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch
import ldcbf_b200 as L
from ldcbf_b200 import scenarios

def main():
    tag = f"max_b={os.environ.get('LDCBF_COOP_MAX_B')} g={os.environ.get('LDCBF_COOP_G')} cold={os.environ.get('LDCBF_COLD')}"
    sc = scenarios.config2(4096, seed=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for B in [int(a) for a in sys.argv[1:]] or (1, 64, 512, 4096, 16384, 65536):
        rep = max(1, B // 4096)
        cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(np.tile(a[:B], (rep,) + (1,) * (a.ndim - 1))), dtype=dt).cuda()
        x0, th, g = cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"])
        ft = cu(scenarios.foot_window(sc["right_first"], 0, 3), torch.int8)
        v, nv, no = cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32)
        prm = L.default_params(0.4)
        if os.environ.get("LDCBF_COLD") == "1":
            prm.flags |= 2
        out = None
        for _ in range(5):
            out = L.mpc_step(prm, x0, th, g, ft, v, nv, no, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.mpc_qp(prm, x0, th, g, ft, out["c_eta"], no, out=out)
            e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        it = out["iters"].float()
        print(f"[{tag}] B={x0.shape[0]} qp-only p50 {statistics.median(ts):.1f} us min {min(ts):.1f}  iters mean {it.mean().item():.2f} "
              f"max {int(it.max().item())} status {torch.bincount(out['status']).tolist()} objsum {out['obj'].nansum().item():.9f}", flush=True)

if __name__ == "__main__":
    main()
