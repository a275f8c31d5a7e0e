"""Multi-GPU plan: prompts are independent, so rank r samples the contiguous shard
``[r*N/G, (r+1)*N/G)`` with its own replica of the weights and its own CUDA graph -- no collective on the data path --
and ONE all-gather returns the padded motions (NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no
inference-time collective (single test device, ``src/test.py:100``)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one) split of n prompts."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def agree_max_len(local_max_len: int, device=None) -> int:
    """max over ranks of the local ``max(lengths)``: ``LADiffVae.decode`` pads to the LOCAL maximum (reference
    ladiff_vae.py:360), which differs between ranks for ragged shards, and an all-gather needs one shape."""
    if not _is_dist():
        return int(local_max_len)
    t = torch.tensor([int(local_max_len)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def gather_motions(local: torch.Tensor, local_lengths: Sequence[int], n_total: int,
                   max_len: Optional[int] = None) -> Tuple[torch.Tensor, List[int]]:
    """local [n_r, max_len_r, F] (padded frames zero) from every rank -> ([n_total, max_len, F], lengths) on every rank.

    ``max_len_r`` may differ per rank (ragged shards decoded with the reference's ``max(lengths)`` padding): every rank
    pads to the agreed global maximum (``max_len`` if given -- e.g. ``cfg`` max_frames, no collective needed -- else an
    all-reduce(MAX) of the local maxima) before the single all-gather."""
    if not _is_dist():
        return local, list(local_lengths)
    world = dist.get_world_size()
    if max_len is None:
        max_len = agree_max_len(local.shape[1], local.device)
    if local.shape[1] > max_len:
        raise ValueError(f"local motions have {local.shape[1]} frames, more than the agreed max_len {max_len}")
    cap = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    tail = tuple(local.shape[2:])
    buf = local.new_zeros((cap, max_len) + tail)
    buf[: local.shape[0], : local.shape[1]] = local
    lens = torch.zeros(cap, dtype=torch.int32, device=local.device)
    lens[: len(local_lengths)] = torch.as_tensor(list(local_lengths), dtype=torch.int32, device=local.device)
    out = local.new_empty((world * cap, max_len) + tail)
    out_l = torch.empty(world * cap, dtype=torch.int32, device=local.device)
    if out.is_cuda:
        dist.all_gather_into_tensor(out, buf)
        dist.all_gather_into_tensor(out_l, lens)
    else:   # gloo has no all_gather_into_tensor: views of the contiguous output
        dist.all_gather(list(out.view(world, cap, max_len, *tail).unbind(0)), buf)
        dist.all_gather(list(out_l.view(world, cap).unbind(0)), lens)
    keep = []
    for r in range(world):
        s, e = shard_range(n_total, r, world)
        keep.extend(range(r * cap, r * cap + (e - s)))
    if len(keep) == world * cap:
        return out, out_l.tolist()
    idx = torch.as_tensor(keep, device=local.device)
    return out.index_select(0, idx), out_l.index_select(0, idx).tolist()
