"""Scenario sharding across ranks (one process per GPU).  A scenario is the atomic unit: batches split
contiguously, every rank runs the rollout kernel on its shard, and the only exchange is one final gather of the small
per-scenario results.  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_scenarios: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split [lo, hi) of n_scenarios over `world` ranks; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_scenarios, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_into(out, local, group=None, async_op: bool = False):
    """all_gather of equally sized shards into a preallocated `out` of shape (world,) + local.shape, issued on the
    CURRENT stream (run it under `torch.cuda.stream(side)` to take the exchange off the compute stream: the sweep's
    next rollout then overlaps it).  Returns the work handle when async_op, else None."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if tuple(out.shape) != (world,) + tuple(local.shape) or out.dtype != local.dtype:
        raise ValueError(f"out must be {(world,) + tuple(local.shape)} of {local.dtype}")
    if not (out.is_contiguous() and local.is_contiguous()):
        raise ValueError("gather_into needs contiguous tensors")
    return dist.all_gather_into_tensor(out.view(-1), local.view(-1), group=group, async_op=async_op)


def gather_results(local, n_scenarios: int, group=None):
    """all_gather of per-scenario results laid out (..., B_local) (scenario index last, as the kernels write them);
    returns (..., n_scenarios) in global scenario order on every rank.  Ragged shards are padded to the largest."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_range(n_scenarios, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[-1] != bmax:
        pad = torch.zeros(local.shape[:-1] + (bmax,), dtype=local.dtype, device=local.device)
        pad[..., : local.shape[-1]] = local
    pad = pad.contiguous()
    out = torch.empty((world,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
    gather_into(out, pad, group=group)                      # 1-D concatenated form: accepted by NCCL and gloo
    return torch.cat([out[r][..., : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=-1)
