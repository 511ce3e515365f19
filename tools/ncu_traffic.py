"""profiles/r2_kernels.json -> profiles/kernel_summary.json: DRAM bytes (read + write) per launch of every profiled kernel,
keyed the way bench.py looks them up for `roofline.traffic` ("rollout" = the headline pass, "k2k3" by batch size)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_kernels.json")
ks = json.load(open(src))
out = {"source": os.path.relpath(src, ROOT) + " (ncu --set full --clock-control none, tools/profile_r2.py on one B200; last = warm "
                 "launch of each kernel; dram__bytes_read.sum + dram__bytes_write.sum)", "kernels": {}, "k2k3": {"dram_bytes_per_launch": {}}}
k2k3_big = 0.0
for k in ks:
    name, grid = k["kernel"], int(k.get("grid", 0))
    b = k.get("dram_rd", 0.0) + k.get("dram_wr", 0.0)
    out["kernels"][f"{name} grid={grid}"] = {"dram_bytes_per_launch": b, "time_us": k.get("time_us")}
    if "rollout_kernel" in name:
        out["rollout"] = {"dram_bytes_per_launch": b, "batch": 4096}
    if "mpc_qp_race_kernel" in name:
        out["k2k3"]["dram_bytes_per_launch"]["4096"] = b
    if "mpc_qp_prepare_kernel" in name or ("mpc_qp_refill_kernel" in name):
        k2k3_big += b
    if "halfplane_kernel" in name and grid >= 8192:          # the B = 2^20 launch (the driver also runs K1 on small batches)
        out["halfplane_kernel"] = {"dram_bytes_per_launch": b, "batch": 1 << 20}
if k2k3_big:
    out["k2k3"]["dram_bytes_per_launch"][str(1 << 20)] = k2k3_big
json.dump(out, open(os.path.join(ROOT, "profiles", "kernel_summary.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
