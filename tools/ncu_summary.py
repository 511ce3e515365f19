"""Turn `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary committed under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/<round>_ncu_full_raw.csv
    python tools/ncu_summary.py profiles/<round>_ncu_full_raw.csv profiles/<round>_kernels.json
Keeps, per kernel name, the LAST profiled launch (the warm one)."""
import csv, json, re, sys

KEEP = {
    "time_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "warp_inst": "smsp__inst_executed.sum",
    "cycles_per_issue": "smsp__average_warp_latency_per_inst_issued.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("ldcbf::", "").strip()
        rec = {"kernel": name}
        for k, m in KEEP.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", ""))
                v *= UNIT_SCALE.get(units[col[m]], 1.0) if k in ("time_us", "dram_rd", "dram_wr") else 1.0
                rec[k] = round(v, 3)
        key = f"{name} grid={int(rec.get('grid', 0))}"
        out[key] = rec
    json.dump(list(out.values()), open(dst, "w"), indent=1)
    for rec in out.values():
        print({k: rec.get(k) for k in ("kernel", "grid", "block", "time_us", "fp64_pipe_pct", "issue_pct", "threads_per_inst", "dram_pct")})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
