"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, 'compute' per-scenario results with the oracle
(standing in for the kernel, which needs a GPU) and gather them; the concatenation must equal the single-rank result
bit for bit, including ragged splits."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import multi_robot_fabrics_b200 as m
    from multi_robot_fabrics_b200.sharding import gather_results, shard_range
    from oracle import o2
    rec = m.scenarios.generate(B, 2, seed=77)
    lo, hi = shard_range(B, rank, world)
    avg, _ = o2.rollout_jointspace_avg(o2.default_config(2), rec[lo:hi], 3)
    local = torch.from_numpy(np.ascontiguousarray(avg.T))            # (R, B_local), scenario last like the kernels
    full = gather_results(local, B)
    # the preallocated, equal-shard form bench.py uses for the sweep's one final exchange ([K][R+1][B] per rank)
    from multi_robot_fabrics_b200.sharding import gather_into
    mine = torch.full((3, 2, 4), float(rank))
    out = torch.empty((world, 3, 2, 4))
    gather_into(out, mine)
    assert all(bool((out[r] == r).all()) for r in range(world))
    try:
        gather_into(torch.empty((world, 3, 2, 5)), mine)
        raise AssertionError("shape mismatch not caught")
    except ValueError:
        pass
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from multi_robot_fabrics_b200.sharding import shard_range
    for B in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            parts = [shard_range(B, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


@pytest.mark.parametrize("B", [10, 9])
def test_two_rank_gather_equals_single_rank(built, B):
    import multi_robot_fabrics_b200 as m
    from oracle import o2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rec = m.scenarios.generate(B, 2, seed=77)
    avg, _ = o2.rollout_jointspace_avg(o2.default_config(2), rec, 3)
    assert np.array_equal(got, avg.T)
