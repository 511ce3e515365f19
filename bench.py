#!/usr/bin/env python
"""bench.py — LDCBF-MPC QP solves/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

`value` is SURVEY.md §8d's config 2: ONE batch of 4096 x N randomised basic simulations (`scenarios.config2`, seed 0),
sharded contiguously over the N ranks (`sharding.shard_bounds`, 4096 per GPU, no data-path collective), each run
CLOSED LOOP to the reference's stop rule (previous objective < 0.05, HumanoidMpc.py:392), to a failed solve
(:419-429) or to 150 steps; a "step" of the bench is one such pass over the batch (one `ldcbf_rollout_f64` launch per
rank), and the number reported is (sum of executed MPC steps over all scenarios and ranks) / time, the time of a
pass measured with CUDA events on the launching stream, L2 flushed between passes, max over ranks.  LDCBF margin
delta = 1e-6 m on both arms.  After the timed region the per-scenario results are gathered over NCCL
(`sharding.gather_results`) and that exchange is timed and check-summed separately (`gather`).

Extra keys on the JSON line: `step0` (the open-loop first step of the same batch: K1 + K2+K3, the number round 1
reported as `value`), `roofline` (dominant kernel, FP64-pipe bound, peak measured live with an FMA-chain probe),
`roofline_hbm` (half-plane builder vs MEASURED_PEAKS.json), `cpu_baseline` (the numpy oracle port on the host
cores; the LiDAR and half-plane blocks carry a `cpu_baseline` of kind "reference": the reference's own functions
from baseline/_ref), `e2e` (pinned host buffers -> H2D -> closed loops -> D2H through
BatchedHumanoidMPC.rollout_host), `large_batch` (open-loop step at B = 2^20 where the GPU is full),
`config4_sharded` (65536-scenario sub-goal rollouts, 8192 per GPU), `per_rank`, `clocks`.
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
MAX_STEPS = 150         # closed loop to the stop rule or 150 steps
MARGIN = 1e-6           # LDCBF margin of both arms (the clearance an interior-point iterate keeps)
WORKLOAD = ("config2: batched basic simulation, 4096 randomized start/goal poses x 3 circle obstacles per GPU, N=3, "
            "T=0.4, closed loop to the 0.05 stop rule / failed solve / 150 steps, executed MPC steps counted")
BYTES_ROLLOUT_STEP = 832.0 + 64.0   # a closed-loop step re-reads the scenario's rings (L1/L2) + trajectory row if recorded
FLOP_K1 = 28.0 * 52                 # ~28 flop per edge x 52 edges — the K1 share of a fused loop step


def make_config(batch_per_gpu):
    """The `config` object of the JSON line — identical on both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": batch_per_gpu, "horizon": N_HORIZON, "obstacles": 3,
            "max_steps": MAX_STEPS, "ldcbf_margin": MARGIN, "seed": 0,
            "sharding": "one seeded batch of batch_per_gpu x n_gpus scenarios (block i = generator seed i), contiguous shard per rank",
            "l2": "flushed (256 MB write) between timed passes on the GPU arm"}
# algorithmic flop model of the fused step kernel at (N, n_obs) = (3, 3), DESIGN.md §6
FLOP_PER_ITER = 470.0
FLOP_SETUP = 500.0
BYTES_K1 = 984.0        # 16*E + 8*7 + 32*n_obs at E = 52, n_obs = 3
BYTES_STEP = 1232.0     # inputs 984 B + outputs 248 B


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline, e2e, step0 and the sharded arms")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the only place bench.py executes oracle/ (the numpy port) and baseline/_ref (the reference's own functions)
# ---------------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
STUB_DIR = os.path.join(ROOT, "tools", "mpl_stub")


def _cpu_init():
    # one BLAS thread per worker process: the port is a scalar loop, oversubscription only slows it down
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    sys.path.insert(0, ROOT)


def _cpu_worker(job):
    """Closed loops of the oracle port (oracle.mpc.run_simulation restates HumanoidMpc.py:380-459) -> executed MPC
    solves (the failed solve that ends a run included, as in the GPU arm's counter, csrc/rollout.cu)."""
    from oracle import mpc
    n = 0
    for state, goal, right_first, rings in job:
        info = {}
        mpc.run_simulation(goal, rings, state, N_horizon=N_HORIZON, N_mpc_timesteps=MAX_STEPS, sampling_time=0.4,
                           start_with_right_foot=bool(right_first), delta=MARGIN, info=info)
        n += info["solves"]
    return n


class CpuPort:
    """The numpy oracle (faithful per-step control flow + exact NNLS solve) on `cores` worker processes."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)

    def rate(self, sc, first, n_scen):
        """Closed loops of scenarios first .. first + n_scen - 1 (mod batch) -> (solves/s, solves, wall seconds)."""
        idx = (first + np.arange(n_scen)) % len(sc["state"])
        jobs = [[(sc["state"][i], sc["goal"][i], sc["right_first"][i], sc["rings"][i]) for i in c]
                for c in np.array_split(idx, min(self.cores, n_scen)) if len(c)]
        t0 = time.perf_counter()
        n = sum(self.pool.map(_cpu_worker, jobs))
        wall = time.perf_counter() - t0
        return n / wall, n, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def _ref_init():
    """Worker of the reference-function legs: the UNMODIFIED reference package from baseline/_ref behind the matplotlib
    stand-in (the reference imports matplotlib at module level for its plots)."""
    sys.path[:0] = [STUB_DIR, REF_DIR]
    for m in [m for m in sys.modules if m == "HumanoidNavigation" or m.startswith("HumanoidNavigation.")]:
        del sys.modules[m]


def _ref_lidar_worker(job):
    """compute_lidar_readings (range_finder_wth_polygons_dbscan.py:26-63) on (position, obstacle vertex arrays) pairs."""
    from HumanoidNavigation.RangeFinder.range_finder_wth_polygons_dbscan import compute_lidar_readings
    n = 0
    for pos, obstacles, rng_, res in job:
        compute_lidar_readings(pos, obstacles, rng_, res)
        n += 1
    return n


def _ref_halfplane_worker(job):
    """get_closest_point_and_normal_vector_from_obs (ObstaclesUtils.py:60-109) for every obstacle of every scenario,
    called as HumanoidMpc.py:311-317 does."""
    from scipy.spatial import ConvexHull
    from HumanoidNavigation.Utils.ObstaclesUtils import ObstaclesUtils
    hulls = {}
    n = 0
    for pos, rings, key in job:
        if key not in hulls:
            hulls[key] = [ConvexHull(r) for r in rings]
        for h in hulls[key]:
            ObstaclesUtils.get_closest_point_and_normal_vector_from_obs(x=pos, polygon=h, unitary_normal_vector=True)
        n += 1
    return n


def reference_function_rate(worker, items, cores, seconds=6.0):
    """units/s of one of the reference's own functions, one process per core, on a bounded sample of `items`
    (repeated until about `seconds` of wall time).  None when baseline/_ref is not there."""
    if not os.path.isdir(os.path.join(REF_DIR, "HumanoidNavigation")):
        return None
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores, initializer=_ref_init) as pool:
        chunk = [items[i::cores] for i in range(cores)]
        chunk = [c for c in chunk if c]
        pool.map(worker, [c[:1] for c in chunk])                 # imports
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            n += sum(pool.map(worker, chunk))
        wall = time.perf_counter() - t0
    return {"value": n / wall, "cores": len(chunk), "kind": "reference", "n": n, "wall_s": wall}


def run_reference(args):
    """`--impl reference`: the reference's CPU path for this hot path.  CasADi/IPOPT are not installable offline
    (SURVEY.md §8c), so this times the oracle port — a numpy restatement of HumanoidMpc.py:380-455 with an exact
    QP solve — on all host cores, on the same config/metric: closed loops of scenarios of the SAME seeded batch, one
    scenario per core per "step" (a bounded sample of the 4096 x n_gpus closed loops of a pass)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ldcbf_b200 import scenarios
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 20))
    warm = min(args.warmup, 1)
    per = 4 * cores                                   # closed loops per "step": four per core
    sc = scenarios.config2(min(args.batch, per * (steps + warm)), seed=0)    # a prefix of the seed-0 batch
    port = CpuPort(cores)
    for w in range(warm):
        port.rate(sc, w * per, per)
    t_total, n_total = 0.0, 0
    for i in range(steps):
        _, n, wall = port.rate(sc, (warm + i) * per, per)
        t_total += wall
        n_total += n
    port.close()
    value = n_total / t_total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.batch),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per} closed loops (scenarios {warm * per}.. of the seed-0 batch, four per "
                                       f"core) per step x {steps} steps = {n_total} MPC solves; numpy oracle "
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


def shard_of(sc, lo, hi):
    out = {}
    for k, v in sc.items():
        out[k] = v[lo:hi] if isinstance(v, (np.ndarray, list)) and len(v) == len(sc["state"]) else v
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ldcbf_b200 as L
    from ldcbf_b200 import scenarios, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON result: everything else any library writes to file descriptor 1 (NCCL
    # prints its version banner there) goes to stderr; the result is written to the saved descriptor at the end
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    # fork the CPU-baseline workers before this process creates a CUDA context
    port = CpuPort(os.cpu_count() or 1) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.lib()
    Bg, N = args.batch, N_HORIZON
    B_total = Bg * world
    # ONE seeded batch for the whole job, contiguous shard per rank: block i of the batch is
    # config2(batch, seed = i), so a rank draws only the blocks its shard touches and no data moves between ranks
    lo, hi = sharding.shard_bounds(B_total, world, rank)
    sc = scenarios.config2_sharded(B_total, lo, hi, seed=0, block=Bg)
    B = hi - lo
    foots = scenarios.foot_window(sc["right_first"], 0, N)
    d = device_inputs(sc, foots, torch)
    prm = L.default_params(0.4)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=N, sampling_time=0.4,
                               delta=np.full(B, MARGIN))
    st0 = torch.as_tensor(sc["state"], dtype=torch.float64).cuda()
    rf = torch.as_tensor(sc["right_first"].astype(np.int8)).cuda()
    state = st0.clone()
    res = {}

    def closed_loops():
        res["r"] = eng.rollout(state, rf, MAX_STEPS, record=False)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(x, op):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    W = max(3, args.warmup)
    for _ in range(W):
        state.copy_(st0)
        closed_loops()
    sync_all()
    clk = Clocks(local) if rank == 0 else None          # ONE nvidia-smi poller per job, not one per rank
    if clk:
        clk.__enter__()
    sync_all()
    ts = []
    for _ in range(args.steps):
        state.copy_(st0)                                 # untimed: back to the initial states
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        closed_loops()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    sync_all()
    r = res["r"]
    solves_rank = float(r["total_solves"].item())
    iters_rank = float(r["total_iters"].item())
    total_ms = allreduce(sum(ts), dist.ReduceOp.MAX if world > 1 else None)
    solves_job = allreduce(solves_rank, dist.ReduceOp.SUM if world > 1 else None)
    value = solves_job * args.steps / (total_ms * 1e-3)
    ends = torch.bincount(r["end_code"], minlength=len(L.binding.END_NAMES)).double()
    if world > 1:
        dist.all_reduce(ends)
    per_rank = torch.tensor([statistics.median(ts), max(ts), sum(ts), solves_rank], dtype=torch.float64, device="cuda")
    ranks = [torch.empty_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(ranks, per_rank)
    else:
        ranks = [per_rank]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": make_config(Bg),
            "value_is": "config 2 closed loop (SURVEY.md 8d): executed MPC steps of all scenarios / time of the pass; "
                        "round 1 reported the open-loop first step here - that number is now `step0`",
            "solver": "dual active set (Goldfarb-Idnani) in CoM-position space, fp64, warm-started from the previous "
                      "step's shifted active set; K1 + K2+K3 fused in one persistent launch per pass (rollout_kernel)",
            "gpu_launches": args.steps, "solves_per_pass": solves_job, "mean_steps_per_scenario": solves_job / B_total,
            "iters_mean": iters_rank / max(1.0, solves_rank), "p50_step_us": 1e3 * statistics.median(ts),
            "p50_solve_step_us": 1e3 * statistics.median(ts) / float(r["steps"].max().item()),
            "endings": {n: int(v) for n, v in zip(L.binding.END_NAMES, ends.tolist())},
            "endings_note": "infeasible_future_rows: LDCBF rows of the predicted stages conflict with the kinematic rows "
                            "(the reference formulation is not recursively feasible; IPOPT raises there too, "
                            "HumanoidMpc.py:419-429); infeasible_k0_row: current CoM violates a half-plane by > 1e-6",
            "per_rank": [{"rank": i, "p50_ms": float(t[0]), "max_ms": float(t[1]), "sum_ms": float(t[2]),
                          "solves_per_pass": float(t[3])} for i, t in enumerate(x.tolist() for x in ranks)]}

    # ---- C1: gather of the fixed-size per-scenario results over NCCL, after the timed region
    rows = torch.cat((state, r["steps"].double()[:, None], r["status"].double()[:, None],
                      r["end_code"].double()[:, None]), dim=1).contiguous()
    if world > 1:
        for _ in range(2):
            allrows = sharding.gather_results(rows, B_total)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        allrows = sharding.gather_results(rows, B_total)
        e1.record()
        e1.synchronize()
        g_ms = allreduce(e0.elapsed_time(e1), dist.ReduceOp.MAX)
    else:
        allrows, g_ms = rows, 0.0
    w = torch.arange(1, allrows.shape[0] + 1, dtype=torch.float64, device="cuda")
    checksum = float((allrows[:, 5] * w).sum().item())          # step counts weighted by the global scenario index
    cs_lo = allreduce(checksum, dist.ReduceOp.MIN if world > 1 else None)
    cs_hi = allreduce(checksum, dist.ReduceOp.MAX if world > 1 else None)
    line["gather"] = {"collective": "all_gather (NCCL)" if world > 1 else "none (1 rank)", "ms": g_ms,
                      "bytes_per_rank_in": rows.numel() * 8, "bytes_out": allrows.numel() * 8,
                      "rows": int(allrows.shape[0]), "checksum": checksum, "checksum_equal_on_all_ranks": cs_lo == cs_hi,
                      "solves_from_gathered_rows": float(allrows[:, 5].sum().item())}

    # ---- end to end through the public API with host buffers (every rank; max over ranks)
    e2e_res = e2e(eng, sc, args, torch)
    e2e_ms = allreduce(e2e_res["ms_total"], dist.ReduceOp.MAX if world > 1 else None)
    line["e2e"] = {"value": solves_job * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": e2e_res["h2d"], "d2h_bytes_per_step": e2e_res["d2h"],
                   "ms_per_step": e2e_ms / args.steps,
                   "api": "BatchedHumanoidMPC.rollout_host -> ldcbf_rollout_f64 (pinned [B,5] states + [B] first foot in, "
                          "pinned [B,8] result rows out, stream synchronise, every pass)"}

    # ---- config 4 sharded: ONE batch of 8192 x n_gpus sub-goal scenarios, shard per rank, gather afterwards
    line["config4_sharded"] = config4_sharded(L, sharding, world, rank, torch, dist, allreduce, sync_all)

    # ---- the open-loop first step of the same batch (round 1's `value`), per rank, max over ranks
    out = {}
    step = lambda: L.mpc_step(prm, d["x0"], d["th"], d["goal"], d["foot"], d["verts"], d["nverts"], d["nobs"],
                              delta=eng.delta, out=out)
    for _ in range(5):
        step()
    sync_all()
    t0s = timed_steps(step, max(args.steps, 50), flush, torch)
    sync_all()
    s0_ms = allreduce(sum(t0s), dist.ReduceOp.MAX if world > 1 else None)
    iters = out["iters"].double()
    line["step0"] = {"value": B_total * len(t0s) / (s0_ms * 1e-3), "unit": UNIT, "ms_per_step": s0_ms / len(t0s),
                     "p50_step_us": 1e3 * statistics.median(t0s), "iters_mean": float(iters.mean().item()),
                     "status_counts": torch.bincount(out["status"], minlength=4).tolist(), "gpu_launches_per_step": 2,
                     "what": "one batched open-loop MPC step (K1 + K2+K3) over the rank's 4096 initial states, L2 flushed"}

    if rank == 0:
        peak_fp64 = max(L.probe_fp64() for _ in range(3))
        # dominant kernel of the headline: rollout_kernel (K1 + K2+K3 fused, one thread x 4 lanes per scenario)
        flops = iters_rank * FLOP_PER_ITER + solves_rank * (FLOP_SETUP + FLOP_K1)
        k_ms = statistics.mean(ts)
        ach = flops / (k_ms * 1e-3) / 1e12
        lanes = 4 if B < 3072 else 2 if B < 7168 else 1             # csrc/rollout.cu: launch_rollout
        line["roofline"] = {"bound": "fp64", "kernel": f"rollout_kernel<3,4,exact,32,{lanes}>", "achieved": ach, "peak": peak_fp64,
                            "unit": "TFLOP/s", "frac": ach / peak_fp64, "traffic": profile_traffic("rollout"),
                            "kernel_ms": k_ms,
                            "peak_source": "FP64 FMA-chain probe measured in this run (MEASURED_PEAKS.json has no fp64 "
                                           "entry); profiles/ holds the probe's own ncu counters",
                            "flop_model": f"iterations*{FLOP_PER_ITER:.0f} + solves*({FLOP_SETUP:.0f} + {FLOP_K1:.0f} for the "
                                          "52 ring edges) (DESIGN.md §6)",
                            "note": "4096 closed loops are 4096 sequential chains of <= 150 dependent solves: 256 warps on "
                                    "592 SM sub-partitions, bound by the instruction latency of the longest chain (ncu: "
                                    "4.1 cycles per issued instruction, 2.0 of them fixed-latency dependencies); "
                                    "large_batch shows the same solver with the GPU full"}
        # kernel-only timing of the open-loop solve for step0's roofline (same inputs, L2 flushed)
        t_qp = timed_steps(lambda: L.mpc_qp(prm, d["x0"], d["th"], d["goal"], d["foot"], out["c_eta"], d["nobs"],
                                            delta=eng.delta, out=out), 50, flush, torch)
        f0 = float((iters * FLOP_PER_ITER + FLOP_SETUP).sum().item())
        a0 = f0 / (statistics.mean(t_qp) * 1e-3) / 1e12
        line["step0"]["roofline"] = {"bound": "fp64", "kernel": "mpc_qp_race_kernel<3,4,16>", "achieved": a0,
                                     "peak": peak_fp64, "unit": "TFLOP/s", "frac": a0 / peak_fp64,
                                     "kernel_ms": statistics.mean(t_qp), "traffic": profile_traffic(B)}
        if world == 1 and not args.no_extras:
            line["large_batch"] = large_batch(L, sc, foots, prm, flush, peak_fp64, torch)
            # the memory-bound kernel of the path against the measured HBM peak (it needs a full GPU to mean anything)
            line["roofline_hbm"] = dict(line["large_batch"]["roofline_hbm"], batch=line["large_batch"]["batch"],
                                        algorithmic_bytes_per_scenario=BYTES_K1)
            # the other rows of the hot path
            line["rollout_margin_0"] = rollout_bench(L, sc, torch, delta=0.0)
            line["rollout_cold_start"] = rollout_bench(L, sc, torch, delta=MARGIN, cold=True)
            line["lidar"] = lidar_bench(L, flush, peak_fp64, torch)
            line["subgoal_rollout"] = subgoal_rollout_bench(L, torch)
            line["unknown_env"] = unknown_env_bench(L, flush, torch)
            line["unknown_env_rollout"] = unknown_env_rollout_bench(L, torch)
            line["latency_b1"] = latency_b1(L, torch)
            line["bounds_tuning"] = bounds_tuning_bench(torch)
            line["long_horizon"] = long_horizon_bench(L, torch, peak_fp64)
            line["clearance_grid"] = clearance_bench(L, torch)
    if clk:
        clk.__exit__()
    if rank == 0:
        line["clocks"] = clk.summary()
        if port is not None:
            cores = port.cores
            port.rate(sc, 0, cores)                                   # warm the workers (imports)
            r0, _, w0 = port.rate(sc, cores, cores)                   # calibration: one closed loop per core
            n_scen = int(max(cores, min(cores * round(15.0 / max(w0, 1e-3)), 4096)))    # about 15 s of CPU work
            rate, n, wall = port.rate(sc, 2 * cores, n_scen)
            port.close()
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_scen} closed loops of the same seed-0 batch (scenarios {2 * cores}..) = {n} "
                                              f"MPC solves in {wall:.1f} s; numpy oracle (reference restatement; CasADi/IPOPT "
                                              "unavailable offline)"}
            if not args.no_extras and "lidar" in line:
                line["lidar"]["cpu_baseline"] = lidar_reference_baseline(cores)
                line["roofline_hbm"]["cpu_baseline"] = halfplane_reference_baseline(sc, cores)
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config4_sharded(L, sharding, world, rank, torch, dist, allreduce, sync_all, per_gpu=8192, per_goal=300):
    """BASELINE.json config 4: 8192 x n_gpus sub-goal scenarios (65536 on 8 GPUs) as ONE seeded batch, contiguous shard
    per rank, sequential sub-goal runs in one rollout launch per rank (HumanoidMPCWithRRT.py:153-181), then the NCCL
    gather of the per-scenario result rows (timed separately)."""
    from ldcbf_b200 import scenarios
    B_total = per_gpu * world
    lo, hi = sharding.shard_bounds(B_total, world, rank)
    c4 = shard_of(scenarios.config4(B_total, seed=0), lo, hi)
    B = hi - lo
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    prm = L.default_params(0.4)
    goals, v, nv, no = cu(c4["goals"]), cu(c4["verts"]), cu(c4["nverts"], torch.int32), cu(c4["nobs"], torch.int32)
    rf = cu(c4["right_first"].astype(np.int8), torch.int8)
    delta = torch.full((B,), MARGIN, dtype=torch.float64, device="cuda")
    G = goals.shape[1]
    st0 = cu(c4["state"])
    state = st0.clone()
    run = lambda: L.rollout(prm, state, goals, rf, v, nv, no, T=G * 120, N=N_HORIZON, max_steps_per_goal=per_goal,
                            delta=delta, record=False)
    for _ in range(2):
        state.copy_(st0)
        r = run()
    sync_all()
    ts = []
    for _ in range(3):
        state.copy_(st0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = run()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    sync_all()
    op_max = dist.ReduceOp.MAX if world > 1 else None
    op_sum = dist.ReduceOp.SUM if world > 1 else None
    ms = allreduce(statistics.median(ts), op_max)
    solves = allreduce(float(r["total_solves"].item()), op_sum)
    reached = allreduce(float((r["goal_steps"][:, -1] > 0).sum().item()), op_sum)
    rows = torch.cat((state, r["steps"].double()[:, None], r["status"].double()[:, None],
                      r["goal_steps"].double()), dim=1).contiguous()
    if world > 1:
        for _ in range(2):
            allrows = sharding.gather_results(rows, B_total)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        allrows = sharding.gather_results(rows, B_total)
        e1.record()
        e1.synchronize()
        g_ms = allreduce(e0.elapsed_time(e1), op_max)
    else:
        allrows, g_ms = rows, 0.0
    w = torch.arange(1, allrows.shape[0] + 1, dtype=torch.float64, device="cuda")
    cs = float((allrows[:, 5] * w).sum().item())
    equal = allreduce(cs, dist.ReduceOp.MIN if world > 1 else None) == allreduce(cs, op_max)
    return {"batch_total": B_total, "batch_per_gpu": per_gpu, "sub_goals": G, "solves": solves, "ms": ms,
            "value": solves / (ms * 1e-3), "unit": UNIT, "scenarios_reaching_last_goal": reached,
            "gather_ms": g_ms, "gather_bytes_out": allrows.numel() * 8, "gather_checksum": cs,
            "checksum_equal_on_all_ranks": equal}


def lidar_reference_baseline(cores, n_scans=64):
    """The reference's own compute_lidar_readings on config-3 scans (same generator, same shapes), one process per core."""
    from ldcbf_b200 import scenarios
    c3 = scenarios.config3(n_scans, seed=0)
    items = [(tuple(c3["pos"][i]), [np.asarray(r) for r in c3["rings"][c3["map_index"][i]]], 1.5, 360)
             for i in range(n_scans)]
    res = reference_function_rate(_ref_lidar_worker, items, cores)
    if res is None:
        return {"value": None, "kind": "reference", "note": "baseline/_ref not present on this box"}
    return {"value": res["value"], "unit": "scans/s", "cores": res["cores"], "kind": "reference",
            "sample": f"{res['n']} scans of 360 rays x 20 obstacles in {res['wall_s']:.1f} s: the reference's own "
                      "compute_lidar_readings (range_finder_wth_polygons_dbscan.py:26-63) from baseline/_ref"}


def halfplane_reference_baseline(sc, cores, n_scen=512):
    """The reference's own get_closest_point_and_normal_vector_from_obs on config-2 scenarios (3 obstacles each)."""
    n_scen = min(n_scen, len(sc["state"]))
    items = [(sc["state"][i][[0, 2]].copy(), [np.asarray(r) for r in sc["rings"][i]], i) for i in range(n_scen)]
    res = reference_function_rate(_ref_halfplane_worker, items, cores)
    if res is None:
        return {"value": None, "kind": "reference", "note": "baseline/_ref not present on this box"}
    return {"value": res["value"], "unit": "scenario-steps/s", "cores": res["cores"], "kind": "reference",
            "sample": f"{res['n']} scenario-steps (3 obstacles, 52 edges) in {res['wall_s']:.1f} s: the reference's own "
                      "get_closest_point_and_normal_vector_from_obs (ObstaclesUtils.py:60-109) from baseline/_ref; the "
                      "ConvexHull objects are built once per scenario outside the count"}


def profile_traffic(key):
    """dram bytes per launch of a kernel from the committed ncu capture (profiles/kernel_summary.json, written by
    tools/ncu_traffic.py): key = "rollout" (headline pass) or the batch size of the open-loop K2+K3 solve."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "kernel_summary.json")))
        if key == "rollout":
            return prof["rollout"]["dram_bytes_per_launch"]
        return prof["k2k3"]["dram_bytes_per_launch"].get(str(key))
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


def rollout_bench(L, sc, torch, T=MAX_STEPS, delta=0.0, cold=False):
    """Variants of the headline pass for comparison (same scenarios, one launch): `delta` = LDCBF margin
    (HumanoidMPCCustomLCBF.py:30-31; the headline uses 1e-6), `cold` = LDCBF_FLAG_COLD_START (no warm
    start from the previous step's active set)."""
    B = len(sc["state"])
    kw = {"flags": L.binding.FLAG_COLD_START} if cold else {}
    eng = L.BatchedHumanoidMPC(sc["goal"], sc["verts"], sc["nverts"], sc["nobs"], N_horizon=N_HORIZON, sampling_time=0.4,
                               delta=np.full(B, delta), **kw)
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
    ends = torch.bincount(r["end_code"], minlength=len(L.binding.END_NAMES)).tolist()
    return {"batch": B, "max_steps": T, "solves": solves, "ms": ms, "value": solves / (ms * 1e-3), "unit": UNIT,
            "us_per_step_of_batch": 1e3 * ms / max(1, int(r["steps"].max().item())), "delta": delta, "cold_start": cold,
            "iters_mean": float(r["total_iters"].item()) / max(1, solves),
            "endings": dict(zip(L.binding.END_NAMES, ends))}


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


def unknown_env_rollout_bench(L, torch, B=16384, T=60):
    """Config 3 as a closed loop on the device (`ldcbf_rollout_unknown_f64`): every step K4 -> f1 -> K1 -> K2+K3 ->
    advance for all scenarios, no host round trip; starts jittered around the reference's (0, 0, pi/2), noisy readings."""
    from ldcbf_b200 import scenarios
    c3 = scenarios.config3(B, seed=0)
    rs = np.random.default_rng(0)
    st0 = np.zeros((B, 5))
    st0[:, 0], st0[:, 2], st0[:, 4] = rs.uniform(-0.2, 0.2, B), rs.uniform(-0.2, 0.2, B), np.pi / 2
    eng = L.BatchedUnknownEnvMPC(c3["goal"], c3["verts"], c3["nverts"], c3["nobs"], lidar_range=1.5, sampling_time=0.4,
                                 N_horizon=N_HORIZON, delta=np.full(B, MARGIN))
    cu = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    state0, rf = cu(st0, torch.float64), cu(np.ones(B, np.int8), torch.int8)
    noise = torch.randn((B, 360, 2), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0)) * 0.01
    r = eng.rollout(state0.clone(), rf, T, noise=noise, record=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        st = state0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = eng.rollout(st, rf, T, noise=noise, record=False)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    solves = int(r["total_solves"].item())
    ends = torch.bincount(r["end_code"], minlength=len(L.binding.END_NAMES)).tolist()
    return {"batch": B, "max_steps": T, "solves": solves, "ms": ms, "value": solves / (ms * 1e-3), "unit": UNIT,
            "ms_per_step_of_batch": ms / T, "gpu_launches_per_step": 5, "mean_steps": float(r["steps"].double().mean().item()),
            "endings": dict(zip(L.binding.END_NAMES, ends))}


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


def long_horizon_bench(L, torch, peak_fp64=None):
    """Config 5 (scaling sweep): one open-loop MPC step per scenario over the grid horizon {10, 20, 40} x obstacles
    {8, 16, 32, 64}, solved by the block-per-scenario kernel (csrc/mpc_long.cu); batch sized to one wave of resident
    blocks and above.  Flop model of one active-set iteration with n = 2N unknowns and q active rows:
    d = J^T n+ (2n * 2k nonzeros <= 2n^2/2), r = R^-1 d (q^2), z = J2 d2 (2n(n-q)), Householder on the free columns
    (4n(n-q)) or a Givens chain (6n(q-l)), scan of 4N + N*n_obs rows (~12 flop each): about 5n^2 + 12(4N + N n_obs)."""
    from ldcbf_b200 import scenarios
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    prm = L.default_params(0.4)
    rows = []
    grid = [(N, n_obs) for N in (10, 20, 40) for n_obs in (8, 16, 32, 64)]
    batch = {10: 8192, 20: 4096, 40: 1184}
    for N, n_obs in grid:
        B = batch[N]
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
        it = float(out["iters"].double().sum().item())
        n = 2 * N
        flops = it * (5.0 * n * n + 12.0 * (4 * N + N * n_obs))
        row = {"horizon": N, "obstacles": n_obs, "batch": B, "ms": ms, "value": B / (ms * 1e-3), "unit": UNIT,
               "solved": int((st == 0).sum().item()), "infeasible": int((st == 2).sum().item()),
               "iteration_cap": int((st == 1).sum().item()), "mean_iterations": it / B}
        if peak_fp64:
            ach = flops / (ms * 1e-3) / 1e12
            row["roofline"] = {"bound": "fp64", "kernel": "mpc_long_kernel", "achieved": ach, "peak": peak_fp64,
                               "unit": "TFLOP/s", "frac": ach / peak_fp64}
        rows.append(row)
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
        r = L.clearance_grid(*args, h_cap=288, out=r, check=False)       # buffers reused, no read-back inside the window
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    cells = float((r["meta"][:, 4] + 1).sum().item()) * 251
    return {"batch": B, "ms": ms, "value": B / (ms * 1e-3), "unit": "maps/s", "cells_per_s": cells / (ms * 1e-3),
            "note": "timed through the binding with the result buffers of the previous call reused (out=, check=False)"}


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


def e2e(eng, sc, args, torch):
    """Same closed loops through the public API with HOST buffers: pinned states in, pinned result rows out."""
    state_h = torch.as_tensor(sc["state"], dtype=torch.float64).pin_memory()
    rf_h = torch.as_tensor(sc["right_first"].astype(np.int8)).pin_memory()
    for _ in range(3):
        eng.rollout_host(state_h, rf_h, MAX_STEPS)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = eng.rollout_host(state_h, rf_h, MAX_STEPS)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    assert float(res[:, 5].sum()) > 0
    return {"ms_total": max(e0.elapsed_time(e1), wall * 1e3), "h2d": eng.rollout_h2d_bytes, "d2h": eng.rollout_d2h_bytes}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
