"""N > 1 host logic on CPU: contiguous sharding + result gather over a world_size-2 gloo group."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, B, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "humanoid-navigation-using-mpc-ldcbf_b200"))
    from ldcbf_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(B, world, rank)
    full = torch.arange(B * 3, dtype=torch.float64).reshape(B, 3)
    got = sharding.gather_results(full[lo:hi].clone(), B)
    # bench-style timing reduction: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ret[rank] = (bool(torch.equal(got, full)), float(t.item()), lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world, B = 2, 4097
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29531 + os.getpid() % 200, B, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == 2.0 and ret[1][1] == 2.0
    assert (ret[0][2], ret[0][3], ret[1][2], ret[1][3]) == (0, 2049, 2049, 4097)


def test_shard_bounds_cover_everything():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "humanoid-navigation-using-mpc-ldcbf_b200"))
    from ldcbf_b200 import sharding
    for B in (0, 1, 7, 4096, 65537):
        for w in (1, 2, 4, 8):
            b = [sharding.shard_bounds(B, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == B and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _worker_batch(rank, world, port, B, block, ret):
    """bench.py's multi-GPU data path on CPU: every rank draws only ITS shard of the one seeded config-2 batch, runs the
    oracle's first step on it (stand-in for the GPU pass), and the per-scenario rows are gathered on every rank."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "humanoid-navigation-using-mpc-ldcbf_b200")]
    from ldcbf_b200 import scenarios, sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(B, world, rank)
    sc = scenarios.config2_sharded(B, lo, hi, seed=0, block=block)
    rows = torch.as_tensor(np.column_stack((sc["state"], sc["goal"], sc["right_first"].astype(np.float64))))
    got = sharding.gather_results(rows, B)
    w = torch.arange(1, B + 1, dtype=torch.float64)
    cs = torch.tensor([float((got[:, 0] * w).sum())])
    lo_cs, hi_cs = cs.clone(), cs.clone()
    dist.all_reduce(lo_cs, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_cs, op=dist.ReduceOp.MAX)
    ret[rank] = (got.numpy(), float(lo_cs), float(hi_cs))
    dist.barrier()
    dist.destroy_process_group()


def test_one_seeded_batch_sharded_over_two_ranks_equals_the_unsharded_batch():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "humanoid-navigation-using-mpc-ldcbf_b200")]
    from ldcbf_b200 import scenarios
    world, B, block = 2, 80, 32                     # shards cut through the generator blocks (0..40 | 40..80)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_batch, args=(world, 29731 + os.getpid() % 200, B, block, ret), nprocs=world, join=True)
    full = scenarios.config2_sharded(B, 0, B, seed=0, block=block)
    ref = np.column_stack((full["state"], full["goal"], full["right_first"].astype(np.float64)))
    assert np.array_equal(ret[0][0], ref) and np.array_equal(ret[1][0], ref)
    assert ret[0][1] == ret[0][2] == ret[1][1] == ret[1][2]                   # the checksum bench.py prints
    first = scenarios.config2(block, seed=0)
    assert np.array_equal(ref[:block, :5], first["state"])                    # block 0 is config2(block, seed 0)
