"""Host -> device staging of input batches.

The reference moves every batch with `tuple(t.to(args.device) for t in batch)` on the compute stream
(run_pretrain_ml.py:520-526): 53 MB of region features per step copied synchronously from pageable memory
in front of the forward pass.  `PinnedPrefetcher` wraps any iterable of CPU batches (dicts of tensors, e.g.
the reference's DataLoader with a dict collate) and hands out DEVICE batches whose copies ran ahead on a
separate stream from pinned staging buffers:

    for batch in PinnedPrefetcher(loader, device):     # batch: dict of device tensors
        losses = step(batch)

Ordering rules (all enforced with CUDA events, no host synchronisation in steady state):
* the consumer's stream waits for the copy of the batch it receives;
* a slot's device buffers are overwritten only after the work that consumed them has finished (the consumer's
  stream position at its NEXT request);
* a slot's pinned buffers are overwritten by the host only after their previous host->device copy completed.

TSV decoding, tokenisation and masking stay the reference's (oscar_tsv4.py) -- out of scope here.
"""
import torch


class PinnedPrefetcher:
    def __init__(self, source, device, depth=2):
        self.source = source
        self.device = torch.device(device)
        self.depth = max(2, int(depth))
        self.cuda = self.device.type == "cuda"
        self.h2d_bytes = 0  # bytes copied host -> device so far

    def __iter__(self):
        if not self.cuda:  # CPU / debugging: nothing to overlap
            for b in self.source:
                yield {k: v.to(self.device) for k, v in b.items()}
            return
        it = iter(self.source)
        copy = torch.cuda.Stream(self.device)
        slots = [dict(host={}, dev={}, ready=torch.cuda.Event(), consumed=None, copied=None) for _ in range(self.depth)]

        def stage(slot, batch):
            s = slots[slot]
            if s["copied"] is not None:  # the previous H2D out of these pinned buffers must be over
                s["copied"].synchronize()
            for k, v in batch.items():
                if v.is_pinned():
                    s["host"][k] = v  # already page-locked (pin_memory=True loaders): no staging copy
                else:
                    h = s["host"].get(k)
                    if h is None or h.shape != v.shape or h.dtype != v.dtype or not h.is_pinned():
                        h = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                        s["host"][k] = h
                    h.copy_(v)
            for k in list(s["host"]):
                if k not in batch:
                    del s["host"][k]
            with torch.cuda.stream(copy):
                if s["consumed"] is not None:  # the step that read this slot's device buffers has finished
                    copy.wait_event(s["consumed"])
                for k, h in s["host"].items():
                    d = s["dev"].get(k)
                    if d is None or d.shape != h.shape or d.dtype != h.dtype:
                        d = torch.empty(h.shape, dtype=h.dtype, device=self.device)
                        s["dev"][k] = d
                    d.copy_(h, non_blocking=True)
                    self.h2d_bytes += h.numel() * h.element_size()
                for k in list(s["dev"]):
                    if k not in s["host"]:
                        del s["dev"][k]
                s["ready"].record(copy)
                s["copied"] = torch.cuda.Event()
                s["copied"].record(copy)

        staged = []  # slots holding a staged batch, oldest first
        nxt = 0
        for _ in range(self.depth - 1):
            b = next(it, None)
            if b is None:
                break
            stage(nxt, b)
            staged.append(nxt)
            nxt = (nxt + 1) % self.depth
        prev = None
        while staged:
            cur = staged.pop(0)
            if prev is not None:  # everything queued on the consumer's stream so far used the previous batch
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                slots[prev]["consumed"] = ev
            b = next(it, None)
            if b is not None:  # refill the free slot while the consumer works on `cur`
                stage(nxt, b)
                staged.append(nxt)
                nxt = (nxt + 1) % self.depth
            torch.cuda.current_stream(self.device).wait_event(slots[cur]["ready"])
            prev = cur
            yield slots[cur]["dev"]
