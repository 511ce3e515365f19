#!/usr/bin/env python
"""bench.py — LDCBF-MPC QP solves/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one batched MPC step (half-planes K1 + heading/assembly/solve/integrate K2+K3) over the batch of
BASELINE.json config 2 ("Batched basic simulation: 4096 randomized start/goal poses x 3 obstacles"), one solve per
scenario, inputs resident in HBM.  Each rank owns its own copy of the 4096 scenarios (weak scaling with equal work per GPU, no data-path collective);
the time of a step is measured with CUDA events on the launching stream, L2 is flushed between timed steps, and
the job time is the max over ranks.

Extra keys on the JSON line: `roofline` (dominant kernel, FP64-pipe bound, peak measured live with an FMA-chain
probe), `roofline_hbm` (half-plane builder vs MEASURED_PEAKS.json), `cpu_baseline` (the numpy oracle port on the
host cores), `e2e` (host buffers -> H2D -> step -> D2H through BatchedHumanoidMPC.step_host), `p50_step_us`,
`large_batch` (same step at B = 2^20 where the GPU is full), `clocks`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "humanoid-navigation-using-mpc-ldcbf_b200"))

import numpy as np  # noqa: E402

METRIC = "ldcbf_mpc_qp_solves_per_sec"
UNIT = "solves/s"
N_HORIZON = 3
WORKLOAD = "config2: batched basic simulation, 4096 randomized start/goal poses x 3 circle obstacles, N=3, T=0.4"
# algorithmic flop model of the fused step kernel at (N, n_obs) = (3, 3), DESIGN.md §6
FLOP_PER_ITER = 470.0
FLOP_SETUP = 500.0
BYTES_K1 = 984.0        # 16*E + 8*7 + 32*n_obs at E = 52, n_obs = 3
BYTES_STEP = 1232.0     # inputs 984 B + outputs 248 B


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port): the only place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_init():
    # one BLAS thread per worker process: the port is a scalar loop, oversubscription only slows it down
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    sys.path.insert(0, ROOT)


def _cpu_worker(job):
    from oracle import mpc
    states, goals, foots, rings = job
    n = 0
    for s, g, f, r in zip(states, goals, foots, rings):
        mpc.mpc_step(s, g, r, [int(v) for v in f], N=N_HORIZON, sampling_time=0.4)
        n += 1
    return n


class CpuPort:
    """The numpy oracle (faithful per-step control flow + exact NNLS solve) on `cores` worker processes."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)

    def rate(self, sc, foots, n_sample):
        idx = np.arange(n_sample) % len(sc["state"])
        jobs = [(sc["state"][c], sc["goal"][c], foots[c], [sc["rings"][i] for i in c])
                for c in np.array_split(idx, self.cores) if len(c)]
        t0 = time.perf_counter()
        n = sum(self.pool.map(_cpu_worker, jobs))
        wall = time.perf_counter() - t0
        return n / wall, n, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    """`--impl reference`: the reference's CPU path for this hot path.  CasADi/IPOPT are not installable offline
    (SURVEY.md §8c), so this times the oracle port — a numpy restatement of HumanoidMpc.py:380-455 with an exact
    QP solve — on all host cores, on the same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ldcbf_b200 import scenarios
    cores = os.cpu_count() or 1
    sc = scenarios.config2(512, seed=0)
    foots = scenarios.foot_window(sc["right_first"], 0, N_HORIZON)
    per_step = 64 * cores      # bounded sample of the 4096-scenario batch per "step"
    port = CpuPort(cores)
    for _ in range(min(args.warmup, 2)):
        port.rate(sc, foots, per_step)
    t_total, n_total = 0.0, 0
    steps = max(1, min(args.steps, 20))
    for _ in range(steps):
        _, n, wall = port.rate(sc, foots, per_step)
        t_total += wall
        n_total += n
    port.close()
    value = n_total / t_total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": args.batch, "horizon": N_HORIZON, "obstacles": 3},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} scenarios of the batch per step x {steps} steps; numpy oracle "
                                       "(reference restatement; CasADi/IPOPT unavailable offline)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append((float(f[0]), float(f[1]), f[2:6]))
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.samples[0][1],
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def timed_steps(fn, steps, flush, torch):
    """Per-step CUDA-event times (ms) with an L2 flush (write of a > L2 buffer) before every timed step."""
    ts = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts


def device_inputs(sc, foots, torch, rep=1):
    def cu(a, dt):
        a = np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1)))
        return torch.as_tensor(a, dtype=dt).cuda()
    return dict(x0=cu(sc["state"][:, :4], torch.float64), th=cu(sc["state"][:, 4], torch.float64),
                goal=cu(sc["goal"], torch.float64), foot=cu(foots, torch.int8), verts=cu(sc["verts"], torch.float64),
                nverts=cu(sc["nverts"], torch.int32), nobs=cu(sc["nobs"], torch.int32))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # fork the CPU-baseline workers before this process creates a CUDA context
    port = CpuPort(os.cpu_count() or 1) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.lib()
    B, N = args.batch, N_HORIZON
    # Weak scaling with exactly equal work per GPU: every rank solves its own copy of the same config-2 batch (the
    # step time of a latency-bound batch is set by its slowest scenario, so different random batches per rank would
    # measure the luck of the draw, not the scaling).  No data is shared between ranks.
    sc = scenarios.config2(B, seed=0)
    foots = scenarios.foot_window(sc["right_first"], 0, N)
    d = device_inputs(sc, foots, torch)
    prm = L.default_params(0.4)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    out = {}

    def step():
        L.mpc_step(prm, d["x0"], d["th"], d["goal"], d["foot"], d["verts"], d["nverts"], d["nobs"], out=out)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    with Clocks(local) as clk:
        sync_all()
        ts = timed_steps(step, args.steps, flush, torch)
        sync_all()
        total_ms = torch.tensor([sum(ts)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_ms = float(total_ms.item())
        value = B * args.steps * world / (total_ms * 1e-3)
        iters = out["iters"].double()
        status = torch.bincount(out["status"], minlength=4).tolist()

        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "horizon": N, "obstacles": 3,
                           "per_rank_data": "own copy of the same seeded batch on every rank (equal work per GPU)",
                           "timing": "CUDA events per step, L2 flushed (256 MB write) between timed steps",
                           "solver": "dual active set (Goldfarb-Idnani) in CoM-position space, fp64; geometric "
                                     "active-set guess; two racing pivot orders per scenario at this batch size"},
                "p50_step_us": 1e3 * statistics.median(ts), "gpu_launches": 2 * args.steps,
                "iters_mean": float(iters.mean().item()), "status_counts": status}

        # ---- end to end through the public API with host buffers (every rank; max over ranks)
        e2e_res = e2e(L, sc, foots, args, torch)
        e2e_ms = torch.tensor([e2e_res["ms_total"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e_ms = float(e2e_ms.item())
        line["e2e"] = {"value": B * args.steps * world / (e2e_ms * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": e2e_res["h2d"], "d2h_bytes_per_step": e2e_res["d2h"],
                       "ms_per_step": e2e_ms / args.steps,
                       "api": "BatchedHumanoidMPC.step_host -> ldcbf_mpc_step_packed_f64 (one pinned [B,6] state row in, one [B,10] result row out)"}
        if rank == 0:
            # ---- kernel-only timing of the dominant kernel for the roofline (same inputs, L2 flushed)
            t_qp = timed_steps(lambda: L.mpc_qp(prm, d["x0"], d["th"], d["goal"], d["foot"], out["c_eta"], d["nobs"],
                                                out=out), min(args.steps, 50), flush, torch)
            peak_fp64 = max(L.probe_fp64() for _ in range(3))
            flops = float((iters * FLOP_PER_ITER + FLOP_SETUP).sum().item())
            ach = flops / (statistics.mean(t_qp) * 1e-3) / 1e12
            line["roofline"] = {"bound": "fp64", "kernel": "mpc_qp_race_kernel<3,4,16>", "achieved": ach, "peak": peak_fp64,
                                "unit": "TFLOP/s", "frac": ach / peak_fp64, "traffic": profile_traffic(B),
                                "kernel_ms": statistics.mean(t_qp),
                                "peak_source": "FP64 FMA-chain probe measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                                "flop_model": f"sum over scenarios of iters*{FLOP_PER_ITER:.0f} + {FLOP_SETUP:.0f} (iters of the "
                                              "winning path, warm-start rounds counted as iterations; DESIGN.md §6)",
                                "note": "B=4096 is 512 half-warps on 592 SM sub-partitions: bound by the instruction "
                                        "latency of the slowest scenario's path (ncu: 3.9 cycles per issued instruction, "
                                        "1 warp per sub-partition); large_batch shows the solver with the GPU full"}
            # ---- large batch: the regime where the GPU is full
            line["large_batch"] = large_batch(L, sc, foots, prm, flush, peak_fp64, torch)
            # the memory-bound kernel of the path against the measured HBM peak (it needs a full GPU to mean anything)
            line["roofline_hbm"] = dict(line["large_batch"]["roofline_hbm"], batch=line["large_batch"]["batch"],
                                        algorithmic_bytes_per_scenario=BYTES_K1)
            # ---- the other rows of the hot path: closed-loop rollout kernel, LiDAR caster, single-scenario latency
            line["rollout"] = rollout_bench(L, sc, torch)
            line["rollout_margin_1e-6"] = rollout_bench(L, sc, torch, delta=1e-6)
            line["lidar"] = lidar_bench(L, flush, peak_fp64, torch)
            line["subgoal_rollout"] = subgoal_rollout_bench(L, torch)
            line["unknown_env"] = unknown_env_bench(L, flush, torch)
            line["latency_b1"] = latency_b1(L, torch)
            line["bounds_tuning"] = bounds_tuning_bench(torch)
            line["long_horizon"] = long_horizon_bench(L, torch)
            line["clearance_grid"] = clearance_bench(L, torch)
    if rank == 0:
        line["clocks"] = clk.summary()
        if port is not None:
            cores = port.cores
            port.rate(sc, foots, cores * 8)                          # warm the workers (imports)
            r0, _, _ = port.rate(sc, foots, cores * 64)              # calibration
            n_sample = int(max(cores * 64, min(r0 * 15.0, 2e6)))    # about 15 s of CPU work
            rate, n, wall = port.rate(sc, foots, n_sample)
            port.close()
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n} solves ({wall:.1f} s): the scenarios of the same batch, one MPC step each, repeated; numpy "
                                              "oracle (reference restatement; CasADi/IPOPT unavailable offline)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def profile_traffic(batch):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists for this batch."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "kernel_summary.json")))
        return prof["k2k3"]["dram_bytes_per_launch"].get(str(batch))
    except Exception:
        return None


def large_batch(L, sc, foots, prm, flush, peak_fp64, torch, B=1 << 20):
    rep = B // len(sc["state"])
    d = device_inputs(sc, foots, torch, rep)
    Bl = d["x0"].shape[0]
    out = {}
    step = lambda: L.mpc_step(prm, d["x0"], d["th"], d["goal"], d["foot"], d["verts"], d["nverts"], d["nobs"], out=out)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = timed_steps(step, 10, flush, torch)
    # K1 alone ([B,2] positions) and K2+K3 alone
    p2 = d["x0"][:, [0, 2]].contiguous()
    t_hp = timed_steps(lambda: L.half_planes(p2, d["verts"], d["nverts"], d["nobs"], c_eta=out["c_eta"]), 10, flush, torch)
    t_qp = timed_steps(lambda: L.mpc_qp(prm, d["x0"], d["th"], d["goal"], d["foot"], out["c_eta"], d["nobs"], out=out),
                       10, flush, torch)
    iters = out["iters"].double()
    flops = float((iters * FLOP_PER_ITER + FLOP_SETUP).sum().item())
    ach = flops / (statistics.mean(t_qp) * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    gbs_k1 = BYTES_K1 * Bl / (statistics.mean(t_hp) * 1e-3) / 1e9
    return {"batch": Bl, "value": Bl / (statistics.mean(ts) * 1e-3), "unit": UNIT, "ms_per_step": statistics.mean(ts),
            "roofline": {"bound": "fp64", "kernel": "mpc_qp_prepare_kernel<3,4,128> + mpc_qp_refill_kernel<3,4,128,resume>",
                         "achieved": ach, "peak": peak_fp64,
                         "unit": "TFLOP/s", "frac": ach / peak_fp64, "kernel_ms": statistics.mean(t_qp),
                         "traffic": profile_traffic(Bl)},
            "roofline_hbm": {"bound": "hbm", "kernel": "halfplane_kernel<exact>", "achieved": gbs_k1, "peak": hbm_peak,
                             "unit": "GB/s", "frac": gbs_k1 / hbm_peak, "kernel_ms": statistics.mean(t_hp),
                             "peak_source": hbm_src}}


def rollout_bench(L, sc, torch, T=150, delta=None):
    """Config 2 as SURVEY.md §8d states it: closed loop to the 0.05 stop or 150 steps, one kernel launch.
    `delta` = LDCBF margin (HumanoidMPCCustomLCBF.py:30-31).  With an exact solver an active LDCBF row puts the next
    CoM exactly on an obstacle edge, where the reference's normal (x-c)/||x-c|| is numerically undefined
    (ObstaclesUtils.py:98-104) and the run usually ends infeasible; IPOPT's barrier keeps ~1e-6 clearance
    implicitly, which delta = 1e-6 reproduces."""
    B = len(sc["state"])
    eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=N_HORIZON, sampling_time=0.4,
                               delta=None if delta is None else np.full(B, delta))
    rf = torch.as_tensor(sc["right_first"].astype(np.int8)).cuda()
    st0 = torch.as_tensor(sc["state"], dtype=torch.float64).cuda()
    for _ in range(2):
        r = eng.rollout(st0.clone(), rf, T, record=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        st = st0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = eng.rollout(st, rf, T, record=False)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    solves = int(r["total_solves"].item())
    ms = statistics.median(ts)
    return {"batch": B, "max_steps": T, "solves": solves, "ms": ms, "value": solves / (ms * 1e-3), "unit": UNIT,
            "us_per_step_of_batch": 1e3 * ms / max(1, int(r["steps"].max().item())), "delta": delta or 0.0,
            "runs_ending_by_stop_rule": int((r["status"] == 0).sum().item()),
            "runs_ending_infeasible": int((r["status"] == 2).sum().item())}


def subgoal_rollout_bench(L, torch, B=8192, per_goal=300):
    """Config 4 (the per-GPU share of 65536 scenarios over 8 GPUs): sequential sub-goal runs around a wall, a fresh
    run per way-point (HumanoidMPCWithRRT.py:153-181), all in one rollout launch."""
    from ldcbf_b200 import scenarios
    c4 = scenarios.config4(B, seed=0)
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    prm = L.default_params(0.4)
    goals, v, nv, no = cu(c4["goals"]), cu(c4["verts"]), cu(c4["nverts"], torch.int32), cu(c4["nobs"], torch.int32)
    rf = cu(c4["right_first"].astype(np.int8), torch.int8)
    delta = torch.full((B,), 1e-6, dtype=torch.float64, device="cuda")
    G = goals.shape[1]
    run = lambda: L.rollout(prm, cu(c4["state"]), goals, rf, v, nv, no, T=G * 120, N=N_HORIZON,
                            max_steps_per_goal=per_goal, delta=delta, record=False)
    for _ in range(2):
        r = run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = run()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    solves = int(r["total_solves"].item())
    gs = r["goal_steps"]
    return {"batch": B, "sub_goals": G, "solves": solves, "ms": ms, "value": solves / (ms * 1e-3), "unit": UNIT,
            "scenarios_reaching_last_goal": int((gs[:, -1] > 0).sum().item()),
            "mean_steps_per_scenario": float(r["steps"].double().mean().item())}


def lidar_bench(L, flush, peak_fp64, torch, B=16384):
    """Config 3 shape: 360 rays x 20 convex obstacles (~85 edges), lidar_range 1.5 (simulation_1.py:201-231)."""
    from ldcbf_b200 import scenarios
    c3 = scenarios.config3(B, seed=0)
    cu = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    pos, v = cu(c3["pos"], torch.float64), cu(c3["verts"], torch.float64)
    nv, no = cu(c3["nverts"], torch.int32), cu(c3["nobs"], torch.int32)
    rays = L.binding.ray_table(1.5, 360).cuda()
    fn = lambda: L.lidar_cast(pos, v, nv, no, 1.5, 360, rays=rays)
    for _ in range(3):
        ho, he, xy = fn()
    torch.cuda.synchronize()
    ts = timed_steps(fn, 10, flush, torch)
    ms = statistics.mean(ts)
    edges = float(c3["nverts"].sum())
    # the kernel drops, per scan, every obstacle whose bounding box is out of the LiDAR's reach; the flop model only
    # counts the ray-edge tests that remain (same rule, evaluated here with numpy)
    vv, nn = c3["verts"], c3["nverts"]
    valid = np.arange(vv.shape[2])[None, None, :] < nn[:, :, None]
    lo = np.where(valid[..., None], vv, np.inf).min(axis=2)
    hi = np.where(valid[..., None], vv, -np.inf).max(axis=2)
    pc = c3["pos"][:, None, :]
    dd = np.maximum(np.maximum(lo - pc, pc - hi), 0.0)
    reach = (np.arange(vv.shape[1])[None, :] < c3["nobs"][:, None]) & (nn > 0) & ((dd ** 2).sum(-1) <= (1.5 * (1 + 1e-9)) ** 2)
    edges_tested = float((nn * reach).sum())
    flops = 30.0 * 360 * edges_tested               # ~30 flop per ray-edge test
    ach = flops / (ms * 1e-3) / 1e12
    return {"batch": B, "rays": 360, "mean_edges": edges / B, "mean_edges_in_reach": edges_tested / B, "ms": ms,
            "scans_per_s": B / (ms * 1e-3),
            "hit_fraction": float((ho >= 0).float().mean().item()),
            "roofline": {"bound": "fp64", "kernel": "lidar_kernel", "achieved": ach, "peak": peak_fp64,
                         "unit": "TFLOP/s", "frac": ach / peak_fp64}}


def unknown_env_bench(L, flush, torch, B=16384):
    """Config 3 end to end on the device: LiDAR scan (K4) -> clusters + hulls (f1) -> half-planes of the inferred
    obstacles (K1) -> QP (K2+K3), noisy readings (sigma = 0.01 as range_finder `:161-172`, tensor injected)."""
    from ldcbf_b200 import scenarios
    c3 = scenarios.config3(B, seed=0)
    foots = scenarios.foot_window(np.ones(B, bool), 0, N_HORIZON)
    eng = L.BatchedUnknownEnvMPC(c3["goal"], c3["verts"], c3["nverts"], c3["nobs"], lidar_range=1.5, sampling_time=0.4,
                                 N_horizon=N_HORIZON)
    cu = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    x0, th, ft = cu(c3["state"][:, :4], torch.float64), cu(c3["state"][:, 4], torch.float64), cu(foots, torch.int8)
    noise = torch.randn((B, 360, 2), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0)) * 0.01
    for _ in range(3):
        out = eng.step(x0, th, ft, noise=noise)
    torch.cuda.synchronize()
    ts = timed_steps(lambda: eng.step(x0, th, ft, noise=noise), 10, flush, torch)
    xy = out["sensed"]["hit_xy"]
    t_f1 = timed_steps(lambda: L.lidar_clusters(xy, noise=noise), 10, flush, torch)
    ms = statistics.mean(ts)
    return {"batch": B, "ms_per_step": ms, "value": B / (ms * 1e-3), "unit": UNIT, "f1_kernel_ms": statistics.mean(t_f1),
            "mean_inferred_obstacles": float(out["sensed"]["nobs"].double().mean().item()),
            "overflow": int(out["sensed"]["overflow"].sum().item()),
            "status_counts": torch.bincount(out["status"], minlength=4).tolist(), "gpu_launches_per_step": 4}


def latency_b1(L, torch, n=200):
    """p50 latency of one MPC step for ONE scenario (config 1: the reference's basic simulation), device-resident
    inputs: (a) one call of ldcbf_mpc_step_f64 through the binding, (b) the same two launches replayed as a CUDA graph
    (BatchedHumanoidMPC.step(graph=True)); CUDA events around the call, and wall clock including the synchronise."""
    from ldcbf_b200 import scenarios
    rings = scenarios.circle_rings()
    verts, nverts, nobs = scenarios.pack_rings([rings])
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    x0, th, g = cu([[0.0, 0, 3, 0]]), cu([0.0]), cu([[6.0, -3.0]])
    ft, v, nv, no = cu([[1, -1, 1, -1]], torch.int8), cu(verts), cu(nverts, torch.int32), cu(nobs, torch.int32)
    prm = L.default_params(0.4)
    eng = L.BatchedHumanoidMPC(g, v, nv, no, N_horizon=N_HORIZON, sampling_time=0.4)
    out = {}

    def timed(fn):
        ts, tw = [], []
        for i in range(n + 20):
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            if i >= 20:
                ts.append(e0.elapsed_time(e1) * 1e3)
                tw.append((time.perf_counter() - t0) * 1e6)
        return statistics.median(ts), statistics.median(tw)

    d0, w0 = timed(lambda: L.mpc_step(prm, x0, th, g, ft, v, nv, no, out=out))
    d1, w1 = timed(lambda: eng.step(x0, th, ft, graph=True))
    same = bool(torch.equal(out["U"], eng._out["U"]))
    return {"p50_device_us": d1, "p50_wall_us": w1, "plain_call_p50_device_us": d0, "plain_call_p50_wall_us": w0,
            "graph_equals_plain": same,
            "note": "B=1, config 1 step 0; headline numbers are the CUDA-graph replay of the two launches; reference: "
                    "CasADi/IPOPT per step, not measurable offline"}


def long_horizon_bench(L, torch):
    """Config 5 (scaling sweep) samples: one open-loop MPC step per scenario at horizon 10 / 20 / 40 with 8 / 16 / 64
    octagonal obstacles, solved by the block-per-scenario kernel (csrc/mpc_long.cu)."""
    from ldcbf_b200 import scenarios
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    prm = L.default_params(0.4)
    rows = []
    for N, n_obs, B in ((10, 8, 8192), (20, 16, 4096), (40, 64, 1184)):
        sc = scenarios.config5(B, n_obs, seed=0)
        foots = scenarios.foot_window(sc["right_first"], 0, N)
        args = (prm, cu(sc["state"][:, :4]), cu(sc["state"][:, 4]), cu(sc["goal"]), cu(foots, torch.int8),
                cu(sc["verts"]), cu(sc["nverts"], torch.int32), cu(sc["nobs"], torch.int32))
        out = L.mpc_step(*args)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = L.mpc_step(*args, out=out)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        st = out["status"]
        rows.append({"horizon": N, "obstacles": n_obs, "batch": B, "ms": ms, "value": B / (ms * 1e-3), "unit": UNIT,
                     "solved": int((st == 0).sum().item()), "infeasible": int((st == 2).sum().item()),
                     "iteration_cap": int((st == 1).sum().item()),
                     "mean_iterations": float(out["iters"].double().mean().item())})
    return rows


def clearance_bench(L, torch, B=1024):
    """f3: occupancy grid + exact distance transform + clearance cost of the sub-goal planner's front-end
    (HumanoidMPCWithRRT.py:21-88,103-108) for B jittered copies of the config-4 wall map, 251 x 274 cells each."""
    from ldcbf_b200 import scenarios
    rng = np.random.default_rng(0)
    wall = np.array([[2.0, -3.0], [3.0, -3.0], [3.0, 3.0], [2.0, 3.0]])
    rings = [[wall + rng.uniform(-0.2, 0.2, 2)] for _ in range(B)]
    verts, nverts, nobs = scenarios.pack_rings(rings)
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    args = (cu(np.tile([5.0, 0.0], (B, 1))), cu(verts), cu(nverts, torch.int32), cu(nobs, torch.int32))
    r = L.clearance_grid(*args, h_cap=288)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = L.clearance_grid(*args, h_cap=288)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    cells = float((r["meta"][:, 4] + 1).sum().item()) * 251
    return {"batch": B, "ms": ms, "value": B / (ms * 1e-3), "unit": "maps/s", "cells_per_s": cells / (ms * 1e-3),
            "note": "timed through the binding: includes the output allocations (zero-filled) of the call"}


def bounds_tuning_bench(torch):
    """The reference's own batch workload (report_simulations/bounds_tuning.py: 26 880 closed-loop simulations run one
    after the other) as ONE rollout launch with per-scenario limits."""
    from HumanoidNavigation.report_simulations.bounds_tuning import bounds_tuning
    bounds_tuning()                                   # warm-up (allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    best, res, score, steps = bounds_tuning(return_all=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"simulations": int(len(score)), "seconds": dt, "mpc_solves": int(steps.sum() // 4),
            "best_combination": [float(v) for v in best], "best_res": res}


def e2e(L, sc, foots, args, torch):
    eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=N_HORIZON, sampling_time=0.4)
    state6 = np.column_stack((sc["state"], foots[:, 0].astype(np.float64)))      # (..., theta, first stance foot)
    state_h = torch.as_tensor(state6, dtype=torch.float64).pin_memory()
    for _ in range(3):
        eng.step_host(state_h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = eng.step_host(state_h)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    assert int((res[:, 9] == 0).sum()) > 0
    return {"ms_total": max(e0.elapsed_time(e1), wall * 1e3), "h2d": eng.h2d_bytes_per_step,
            "d2h": eng.d2h_bytes_per_step}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
