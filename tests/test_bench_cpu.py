"""CPU checks of bench.py's contract: the reference arm (the oracle port on the host cores) prints exactly one JSON
line with the same `config` object the GPU arm prints, and the closed-loop solve counter of the oracle matches its
trajectory."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_shared_config():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["config"] == bench.make_config(4096)            # what run_ours prints for the default batch
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_oracle_closed_loop_counts_its_solves():
    from ldcbf_b200 import scenarios
    from oracle import mpc
    sc = scenarios.config2(6, seed=0)
    for b in range(6):
        info = {}
        X, U = mpc.run_simulation(sc["goal"][b], sc["rings"][b], sc["state"][b], N_horizon=3, N_mpc_timesteps=40,
                                  sampling_time=0.4, start_with_right_foot=bool(sc["right_first"][b]), delta=1e-6, info=info)
        k = U.shape[1]
        if info["end"] == "stop_rule":
            assert info["solves"] == k
        elif info["end"].startswith("status"):
            assert info["solves"] == k + 1                      # the failed solve is counted, as on the GPU arm
        else:
            assert info["end"] == "step_budget" and info["solves"] == 40 and k == 39   # :458-459 drop the last column
