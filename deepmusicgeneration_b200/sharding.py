"""Batch-dimension sharding of independent generation streams across the GPUs of one box.

Generation streams never exchange data: each rank owns a contiguous slice of the streams, replicated weights and
private K/V rings - there is NO collective on the data path.  torch.distributed is used only for the launch
plumbing (rendezvous, barriers, max-over-ranks of the measured time, gathering the token ids on rank 0).
"""
import os

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    "Contiguous, balanced [lo, hi) slice of `total` streams for `rank` (first `total % world` ranks get one more)."
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_distributed(backend=None):
    "-> (rank, local_rank, world).  Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* as set by torchrun."
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
        import atexit
        atexit.register(finalize)
    return rank, local_rank, world


def finalize():
    "destroy the process group (registered at exit by init_distributed)"
    if dist.is_available() and dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device='cpu'):
    "Max of a python float over all ranks (the timing rule: a multi-GPU number is the slowest rank's)."
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device='cpu'):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_streams(local_tokens, total, device='cpu'):
    """Rank 0 receives the [n_words, total] token matrix assembled from every rank's [n_words, hi-lo] slice
    (host-side gather of outputs only; other ranks get None)."""
    if not dist.is_initialized():
        return local_tokens
    world, rank = dist.get_world_size(), dist.get_rank()
    n_words = local_tokens.shape[0]
    widest = -(-total // world)
    pad = torch.full((n_words, widest), -1, dtype=local_tokens.dtype, device=device)
    pad[:, :local_tokens.shape[1]] = local_tokens.to(device)
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, parts, dst=0)
    if rank != 0:
        return None
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(total, r, world)
        out.append(p[:, :hi - lo])
    return torch.cat(out, dim=1)
