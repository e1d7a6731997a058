"""world_size-2 gloo test (CPU) of the N>1 host logic: stream sharding, max-over-ranks timing, output gather."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from deepmusicgeneration_b200 import sharding


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    r, lr, w = sharding.init_distributed(backend='gloo')
    lo, hi = sharding.shard_range(total, r, w)
    local = torch.arange(lo, hi, dtype=torch.int32)[None].repeat(3, 1) + 1000 * torch.arange(3, dtype=torch.int32)[:, None]
    sharding.barrier()
    slowest = sharding.max_over_ranks(1.0 + r)
    total_tok = sharding.sum_over_ranks(float(hi - lo))
    full = sharding.gather_streams(local, total)
    q.put((r, slowest, total_tok, None if full is None else full.tolist()))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_shard_gather():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    total = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for r, slowest, total_tok, full in res:
        assert slowest == 2.0 and total_tok == float(total)
    full = res[0][3]
    assert res[1][3] is None
    assert full == [[1000 * i + j for j in range(total)] for i in range(3)]
