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

The per-sample byte / integer work of the reference's datasets also has a device-side form (csrc/data_prep.cu):
`decode_features` turns the base64 text of a batch of TSV feature columns into the padded [B, R, K] feature tensor
(oscar_tsv4.py:696-727), `mask_tokens` applies BERT token masking and phrase masking to token ids
(oscar_tsv4.py:782-850).  Tokenisation and TSV indexing stay the reference's.
"""
import torch

from . import _lib


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


def decode_features(b64_texts, num_boxes, max_regions, feature_dim, device, dtype=torch.bfloat16):
    """The reference's `get_img_feature` + zero padding for a whole batch, on the GPU (oscar_tsv4.py:696-727):
    b64_texts = the base64 column of each sample's TSV row (bytes), num_boxes = its box-count column.  Returns
    [B, max_regions, feature_dim] in `dtype` (bfloat16 or float32); raises on malformed text."""
    if dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("dtype must be torch.bfloat16 or torch.float32")
    B = len(b64_texts)
    sizes = [len(t) for t in b64_texts]
    offsets = torch.zeros(B + 1, dtype=torch.int64)
    offsets[1:] = torch.tensor(sizes, dtype=torch.int64).cumsum(0)
    host = torch.empty(int(offsets[-1]) + 16, dtype=torch.uint8).pin_memory()
    view = memoryview(host.numpy())
    pos = 0
    for t in b64_texts:
        view[pos:pos + len(t)] = t
        pos += len(t)
    nb = torch.as_tensor(list(num_boxes), dtype=torch.int32)
    d_src = host.to(device, non_blocking=True)
    d_off, d_nb = offsets.to(device, non_blocking=True), nb.to(device, non_blocking=True)
    out = torch.empty(B, max_regions, feature_dim, device=device, dtype=dtype)
    err = torch.zeros(1, device=device, dtype=torch.int32)
    _lib.call("mvptr_b64_decode_features", d_src, d_off, d_nb, out, int(dtype == torch.float32), B, max_regions,
              feature_dim, feature_dim, max(1, int(nb.max())), err)
    code = int(err)
    if code:
        what = [m for bit, m in ((1, "text length does not match num_boxes x feature_dim float32 values"),
                                 (2, "invalid base64 character"), (4, "padding character before the end")) if code & bit]
        raise _lib.MvptrError("decode_features: " + "; ".join(what))
    return out


def mask_tokens(ids, tok_first, tok_count, mask_id, word_vocab, phr_first=None, phr_count=None, links=None,
                phrase_vocab=0, vocab_size=0, uniforms=None, draws=None, seed=0):
    """BERT masking of oscar_tsv4.py:782-850 on device token ids [B, L] (modified IN PLACE): positions
    [tok_first, tok_first + tok_count) are caption / tag tokens (random_word), [phr_first, phr_first + phr_count)
    phrase concepts (random_phrases; links[b, i, :] = phrase indexes tied to caption token i).  Returns the MLM
    labels [B, L] (-1 = ignore).  `uniforms` / `draws` replay recorded random numbers; otherwise `seed` drives a
    stateless hash."""
    B, L = ids.shape
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    labels = torch.empty_like(ids)
    i32 = lambda t: None if t is None else t.to(device=ids.device, dtype=torch.int32).contiguous()
    lk = i32(links)
    _lib.call("mvptr_mlm_mask", ids, labels, i32(tok_first), i32(tok_count), i32(phr_first), i32(phr_count), lk,
              0 if lk is None else lk.shape[2], None if uniforms is None else uniforms.float().contiguous(),
              None if draws is None else draws.to(torch.int64).contiguous(), B, L, int(mask_id), int(word_vocab),
              int(phrase_vocab), int(vocab_size), int(seed) & 0xFFFFFFFF)
    return labels
