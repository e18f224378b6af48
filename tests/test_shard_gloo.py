"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: cost-balanced sharding and the final
consensus gather.  The per-rank compute is stubbed with the CPU oracle here (test infrastructure);
on GPUs the same code path runs native.consensus per rank (bench.py / tests -m gpu)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from hypo_b200.shard import shard_indices, window_cost
from hypo_b200.synth import random_batch


def test_shards_partition_and_balance():
    b = random_batch(21, 101, kind="mixed", length=60, n_arms=12)
    for world in (1, 2, 4, 8):
        parts = [shard_indices(b, world, r) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert allidx.tolist() == list(range(b.n_win))
        cost = window_cost(b)
        loads = [int(cost[p].sum()) for p in parts]
        assert max(loads) <= 1.25 * (sum(loads) / world) + cost.max()
        assert all((np.diff(p) > 0).all() for p in parts if len(p) > 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hypo_b200.shard import polish_sharded
    from tests.oracle_util import oracle_consensus
    b = random_batch(22, 37, kind="mixed", length=40, n_arms=10)
    res = polish_sharded(b, world, rank, lambda sub: oracle_consensus(sub, threads=1)[0])
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_polish_equals_single_rank_gloo():
    from tests.oracle_util import oracle_consensus
    b = random_batch(22, 37, kind="mixed", length=40, n_arms=10)
    want, _ = oracle_consensus(b)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want
