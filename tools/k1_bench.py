"""Developer timing of the half-plane kernel alone at B = 2^20 (exact and fast arithmetic)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "humanoid-navigation-using-mpc-ldcbf_b200"))
import numpy as np, torch, ldcbf_b200 as L
from ldcbf_b200 import scenarios
from ldcbf_b200.binding import FLAG_FAST_GEOMETRY
sc = scenarios.config2(4096, seed=0)
B = 1 << 20; rep = B // 4096
cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))), dtype=dt).cuda()
import numpy as np
fw = scenarios.foot_window(sc["right_first"], 0, 3)
st = cu(np.column_stack((sc["state"], fw[:, 0].astype(np.float64)))); g = cu(sc["goal"])
pos = cu(sc["state"][:, [0, 2]]); v, nv, no = cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32)
ce = torch.empty((B, 3, 4), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
ms = timeit(lambda: L.half_planes(pos, v, nv, no, c_eta=ce))
print(f"K1 exact B={B}: {ms*1e3:.1f} us  {984*B/ms/1e6:.0f} GB/s")
if "--k1-only" in sys.argv: sys.exit(0)
out = {}
for flags, name in ((0, "exact"), (FLAG_FAST_GEOMETRY, "fast")):
    prm = L.default_params(0.4, flags=flags)
    ms2 = timeit(lambda: L.mpc_step_packed(prm, st, g, v, nv, no, out=out))
    print(f"packed step ({name} geometry): {ms2*1e3:.1f} us  {B/ms2*1e3:.3e} solves/s")
