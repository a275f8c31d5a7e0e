"""CPU, world_size 2 over gloo: the multi-GPU plan of SURVEY.md 8e -- prompts are split into contiguous per-rank shards,
every rank samples only its shard (no data-path collective) and ONE all-gather returns the motions."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ladiff_b200.parallel import gather_motions, shard_range


def test_shard_range_covers_everything_once():
    for n in (1, 7, 128, 8192):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def _worker(rank, world, port, n, q, ragged_pad=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(n, rank, world)
    lengths = [40 + 4 * (i % 40) for i in range(n)]
    # stand-in for the CUDA sampling of this rank's shard: a deterministic function of the global prompt index.
    # ragged_pad: padded to the LOCAL max(lengths) like LADiffVae.decode (ladiff_vae.py:360) -> differs between ranks
    pad = max(lengths[s:e]) if ragged_pad else 196
    local = torch.zeros((e - s, pad, 5))
    for j, i in enumerate(range(s, e)):
        local[j, :lengths[i]] = float(i + 1)
    motions, all_len = gather_motions(local, lengths[s:e], n)
    ok = motions.shape == (n, max(lengths) if ragged_pad else 196, 5) and all_len == lengths
    for i in range(n):
        ok = ok and bool((motions[i, :lengths[i]] == i + 1).all()) and bool((motions[i, lengths[i]:] == 0).all())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged_pad", [False, True])
def test_all_gather_of_motions_world2(ragged_pad):
    """ragged_pad=True: each rank pads to its own max(lengths) (64 vs 88 frames here) -- gather_motions must agree on a
    global max_len first (all-reduce MAX) instead of handing mismatched shapes to the collective."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 13, q, ragged_pad)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
