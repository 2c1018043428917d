"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU path (sharding, uneven all-gather,
bucketed gradient all-reduce, top-k merge)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mvptr_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeArena:
    def __init__(self, n, rank):
        g = torch.Generator().manual_seed(100 + rank)
        self.grad = torch.randn(n, generator=g)
        self.numel = n
        self.device = torch.device("cpu")


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvp_pytorch_b200 import parallel as P
    # shards tile the range exactly
    lo, hi = P.shard_range(11, rank, world)
    # uneven all-gather
    mine = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
    full = P.all_gather_rows(mine)
    assert torch.equal(full[:, 0], torch.arange(11, dtype=torch.float32))
    # bucketed all-reduce == mean of the replicas, whatever the bucket order / coverage
    arena = _FakeArena(1000, rank)
    expect = (_FakeArena(1000, 0).grad + _FakeArena(1000, 1).grad) / 2
    sync = P.GradientSync(arena, min_bucket=0)
    assert sync.enabled and sync.reduce_dtype == torch.float32  # CPU / gloo: fp32 in place
    sync.layer_done(600, 800)
    sync.will_write(650, 700)  # second pass over a range already handed to the collective: must not deadlock
    sync.layer_done(100, 300)
    sync.finish()
    assert torch.allclose(arena.grad, expect, atol=1e-6)
    # small ranges (per-layer bias slices) are deferred to finish() instead of being sent on their own
    arena2 = _FakeArena(1000, rank)
    sync2 = P.GradientSync(arena2, min_bucket=500)
    sync2.layer_done(600, 800)
    assert sync2.done == []
    sync2.layer_done(0, 600)
    sync2.finish()
    assert torch.allclose(arena2.grad, expect, atol=1e-6)
    # sharded scoring: every rank scores its shard of the pairs, results gathered in pair order
    scores = torch.arange(23, dtype=torch.float32) * 0.5
    plo, phi = P.shard_range(23, rank, world)
    assert torch.equal(P.all_gather_rows(scores[plo:phi].clone()), scores)
    if rank == 0:
        out.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_shard_range_partitions():
    from mvp_pytorch_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 25000, 2240000):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def test_merge_topk_equals_global_ranking():
    from mvp_pytorch_b200.parallel import merge_topk
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(5, 40, generator=g)
    scores[:, 7] = scores[:, 3]  # ties
    k = 6
    # two candidate shards, each with its local top-k (global indexes)
    parts_s, parts_i = [], []
    for lo, hi in ((0, 20), (20, 40)):
        idx = O.topk_desc(scores[:, lo:hi], k) + lo
        parts_i.append(idx)
        parts_s.append(torch.gather(scores, 1, idx))
    ms, mi = merge_topk(torch.cat(parts_s, 1), torch.cat(parts_i, 1), k)
    assert torch.equal(mi, O.topk_desc(scores, k))
