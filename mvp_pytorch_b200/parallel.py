"""One-process-per-GPU data parallelism for the hot path.

The reference trains with DeepSpeed ZeRO-2 / DDP (run_pretrain_ml.py:226-227, 406-418) and evaluates
with single-process nn.DataParallel (run_retrieval.py:577-578).  Here every rank owns one GPU and
a full replica; the only data-path exchange of a training step is the all-reduce of the flat gradient
arena (bf16 on the wire, fp32 in the arena), issued bucket by bucket on a side stream while backward is still
running; retrieval
shards its independent units (captions, images, pairs) and all-gathers only embeddings / scores.

Everything in this file is host logic over torch.distributed and runs on the gloo backend too
(tests/test_parallel.py, world_size 2 on CPU).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous [lo, hi) slice of n independent units for this rank (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(t, group=None):
    """Concatenate per-rank tensors that differ in dim 0 (NCCL / gloo all_gather needs equal shapes)."""
    rank, ws = world()
    if ws == 1:
        return t
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    pad = t
    if t.shape[0] < mx:
        pad = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))], 0)
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def merge_topk(scores, indexes, k):
    """Merge per-shard top-k candidate lists [rows, shards*k] into the global top-k with the
    reference ranking order (descending score, ties -> larger global index first)."""
    order = torch.arange(scores.shape[1], device=scores.device).expand_as(scores)
    # sort by (score desc, index desc): stable sort by index desc first, then by score desc
    by_idx = torch.sort(indexes, dim=1, descending=True, stable=True)[1]
    s1, i1 = torch.gather(scores, 1, by_idx), torch.gather(indexes, 1, by_idx)
    by_score = torch.sort(s1, dim=1, descending=True, stable=True)[1]
    del order
    return torch.gather(s1, 1, by_score)[:, :k], torch.gather(i1, 1, by_score)[:, :k]


class GradientSync:
    """Bucketed all-reduce(avg) of the flat gradient arena, overlapped with backward.

    ``layer_done(lo, hi)`` is called by the engine as soon as the gradients of arena range
    [lo, hi) are final (after each encoder layer's backward); the range is all-reduced on a
    communication stream.  ``finish()`` reduces whatever was not covered and joins the streams.
    Every rank runs the same model, so the collective order is identical everywhere.

    ``reduce_dtype``: ``torch.bfloat16`` halves the bytes on NVLink -- what SURVEY 2b specifies and what the
    reference's DeepSpeed fp16 configuration does (tmp_config.json: fp16 gradients): on the communication stream a
    range is narrowed into a bf16 staging arena (mvptr_cast_f32_bf16), all-reduced there, and widened back into the
    fp32 gradient arena (mvptr_cast_bf16_f32), so ``p.grad``, clip_grad_norm_ and the fused AdamW keep seeing fp32
    averaged gradients.  ``torch.float32`` reduces the arena in place.  The default on CUDA is ``"tail-bf16"``, the
    measured best of both (profiles/r2_dp_ab_n2.txt): the per-layer buckets, which run UNDER the backward GEMMs, go
    in place as fp32 -- their wire time is hidden anyway and the two cast kernels per bucket only took SM time from
    the GEMMs (N = 2: 32.5 ms fp32 vs 33.0 ms bf16 per step) -- while what is left for ``finish()``, above all the
    264 MB word-embedding gradient that is final only when backward ends and whose transfer is therefore EXPOSED,
    travels as bf16.

    Ranges below ``min_bucket`` elements (the per-layer bias / LayerNorm slices, ~10 k elements each) are not sent
    on their own -- a collective of a few KB is pure launch latency -- but left to ``finish()``, which sends all
    remaining ranges as few contiguous collectives."""

    def __init__(self, arena, group=None, reduce_dtype=None, min_bucket=1 << 16):
        self.arena, self.group = arena, group
        self.done = []
        self.stream = torch.cuda.Stream() if arena.device.type == "cuda" else None
        self.enabled = world()[1] > 1
        if reduce_dtype is None:
            reduce_dtype = "tail-bf16" if arena.device.type == "cuda" else torch.float32
        if reduce_dtype not in (torch.bfloat16, torch.float32, "tail-bf16"):
            raise ValueError("reduce_dtype must be torch.bfloat16, torch.float32 or 'tail-bf16'")
        self.reduce_dtype = reduce_dtype
        self._in_finish = False
        self.min_bucket = int(min_bucket)
        self.staging = None  # bf16 arena-shaped staging buffer, allocated on first use
        # grid cap of the widening cast (0 = none).  A narrow grid looked polite but is not: capped at 32 CTAs each
        # cast took 0.21 ms (20 per step = 4.2 ms of comm-stream time next to the backward GEMMs, which ran 4-10 %
        # slower under it; profiles/r2_dp_trace_n2_capped_casts.txt) -- at full width it is a 20 us kernel
        self.cast_ctas = int(__import__("os").environ.get("MVPTR_DP_CAST_CTAS", "0"))

    def _reduce(self, lo, hi):
        g = self.arena.grad[lo:hi]
        if self.stream is None:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            g.div_(world()[1])
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            if self.reduce_dtype == torch.float32 or (self.reduce_dtype == "tail-bf16" and not self._in_finish):
                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
                return
            from . import _lib
            if self.staging is None:
                self.staging = torch.empty(self.arena.numel, device=self.arena.device, dtype=torch.bfloat16)
            s = self.staging[lo:hi]
            _lib.call("mvptr_cast_f32_bf16", g, s, hi - lo)
            dist.all_reduce(s, op=dist.ReduceOp.AVG, group=self.group)
            _lib.call("mvptr_cast_bf16_f32", s, g, hi - lo, self.cast_ctas)

    def will_write(self, lo, hi):
        """Called before backward kernels accumulate into arena range [lo, hi).  If that range was already
        handed to the communication stream since the last finish() (a second backward before the optimizer
        step: gradient accumulation, or the same layers run twice in one forward), the compute stream first
        waits for that collective -- writing under an in-flight all-reduce would race.  The result stays
        right by linearity: avg(avg(g1) + g2) = avg(g1) + avg(g2)."""
        if self.enabled and self.stream is not None and any(l < hi and lo < h for l, h in self.done):
            torch.cuda.current_stream().wait_stream(self.stream)

    def layer_done(self, lo, hi):
        if not self.enabled or hi <= lo or hi - lo < self.min_bucket:
            return
        self._reduce(lo, hi)
        self.done.append((lo, hi))

    def finish(self):
        if not self.enabled:
            return
        pos = 0
        self._in_finish = True
        try:
            for lo, hi in sorted(self.done):
                if lo > pos:
                    self._reduce(pos, lo)
                pos = max(pos, hi)
            if pos < self.arena.numel:
                self._reduce(pos, self.arena.numel)
        finally:
            self._in_finish = False
        self.done = []
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


def allreduce_gradients(model, group=None):
    """Average the gradients of all replicas (call after backward(), before optimizer.step()).
    Uses the overlapped bucket schedule when ``enable_overlapped_allreduce(model)`` was called
    before backward; otherwise reduces the whole arena in one collective."""
    rt = model.runtime()
    sync = getattr(rt, "grad_sync", None)
    if sync is None:
        sync = GradientSync(rt.arena, group)
    sync.finish()


def enable_overlapped_allreduce(model, group=None, reduce_dtype=None, min_bucket=1 << 16):
    """Register the bucketed, backward-overlapped gradient all-reduce on a model (see GradientSync)."""
    rt = model.runtime()
    rt.arena.ensure_grad()
    rt.grad_sync = GradientSync(rt.arena, group, reduce_dtype=reduce_dtype, min_bucket=min_bucket)
    return rt.grad_sync
