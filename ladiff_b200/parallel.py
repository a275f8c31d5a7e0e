"""Multi-GPU plan: prompts are independent, so rank r samples the contiguous shard
``[r*N/G, (r+1)*N/G)`` with its own replica of the weights and its own CUDA graph -- no collective on the data path --
and ONE all-gather returns the padded motions (NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no
inference-time collective (single test device, ``src/test.py:100``)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one) split of n prompts."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_motions(local: torch.Tensor, local_lengths: Sequence[int], n_total: int) -> Tuple[torch.Tensor, List[int]]:
    """local [n_r, max_len, F] (padded frames zero) from every rank -> ([n_total, max_len, F], lengths) on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local, list(local_lengths)
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    buf = local.new_zeros((cap,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    lens = torch.zeros(cap, dtype=torch.int32, device=local.device)
    lens[: len(local_lengths)] = torch.as_tensor(list(local_lengths), dtype=torch.int32, device=local.device)
    out = local.new_empty((world * cap,) + tuple(local.shape[1:]))
    out_l = torch.empty(world * cap, dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(out, buf) if out.is_cuda else dist.all_gather(list(out.view(world, cap, *local.shape[1:]).unbind(0)), buf)
    dist.all_gather_into_tensor(out_l, lens) if out_l.is_cuda else dist.all_gather(list(out_l.view(world, cap).unbind(0)), lens)
    keep = []
    for r in range(world):
        s, e = shard_range(n_total, r, world)
        keep.extend(range(r * cap, r * cap + (e - s)))
    idx = torch.as_tensor(keep, device=local.device)
    return out.index_select(0, idx), out_l.index_select(0, idx).tolist()
