"""Whole training step as ONE CUDA graph.

`GraphedTrainStep` captures zero_grad + forward + backward + gradient all-reduce + fused AdamW of a
model built from this package into a single `torch.cuda.CUDAGraph`.  A replay costs one launch, so
the ~500 kernel launches of a step no longer depend on how fast (or how often interrupted) the
host thread is, and the launch gaps between them disappear.

What makes the step capturable:
* dropout: the per-site seeds are baked into the graph; a pinned host counter is copied into the
  library's dropout epoch word at the start of every replay (`mvptr_set_dropout_epoch`), so each
  replay draws fresh masks;
* AdamW: learning rate and step count travel the same way (pinned host -> device float[2]), so LR
  schedulers keep working (`optimizer.param_groups[0]['lr']` is read before every replay);
* masked-LM row selection: fixed-capacity `torch.nonzero_static` instead of the synchronising
  `torch.nonzero`; unused slots carry label -1.  More labels than slots raise `overflow`
  (checked by `check_overflow()`), never silently dropped;
* inputs live in static device buffers that `__call__` refills.
"""
import torch

from . import _lib


def _round_up(x, m):
    return (int(x) + m - 1) // m * m


class GraphedTrainStep:
    def __init__(self, model, optimizer, sample_batch, forward_kwargs=None, mlm_capacity=None, warmup=3,
                 allreduce=False):
        self.model, self.opt = model, optimizer
        self.kw = dict(forward_kwargs or {})
        self.allreduce = allreduce
        dev = next(model.parameters()).device
        self.static = {k: v.to(dev).clone() for k, v in sample_batch.items()}
        if mlm_capacity is None and "masked_lm_labels_a" in sample_batch:
            # 1.2x the sample's count + 64 (BERT masking is binomial: ~8 sigma at batch 256), 128-row granules
            n_txt = int((sample_batch["masked_lm_labels_a"] > -1).sum())
            n_vis = int((sample_batch["masked_lm_labels_b"] > -1).sum())
            mlm_capacity = (_round_up(n_vis * 1.2 + 64, 128), _round_up(n_txt * 1.2 + 64, 128))
        if mlm_capacity is not None:
            model.mlm_capacity = tuple(int(c) for c in mlm_capacity)
            model.mlm_overflow = torch.zeros((), dtype=torch.bool, device=dev)
        self.epoch_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        optimizer.enable_graph_mode()
        self.losses = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.epoch_host[0] += 1
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()
        # the capture pass advanced the optimizer's host step without executing: undo
        optimizer._step -= 1

    def _step(self):
        _lib.call("mvptr_set_dropout_epoch", self.epoch_host)
        self.model.zero_grad()
        out = self.model(**self.static, **self.kw)
        out[0].backward()
        if self.allreduce:
            from .parallel import allreduce_gradients
            allreduce_gradients(self.model)
        self.opt.step()
        self.losses = torch.stack([o.detach().float() for o in out])

    def load(self, batch, non_blocking=True):
        """Refill the static input buffers (device-to-device or host-to-device copies)."""
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=non_blocking)

    def __call__(self, batch=None):
        if batch is not None:
            self.load(batch)
        self.epoch_host[0] += 1
        self.opt.before_replay()
        self.graph.replay()
        return self.losses

    def release(self):
        """Drop the captured graph (and the NCCL work recorded in it).  Call before destroying the process
        group; the model and optimizer stay usable for eager steps."""
        self.graph.reset()
        self.graph = None

    def check_overflow(self):
        """Raises if any replay saw more masked-LM labels than the captured capacity (host sync)."""
        ovf = getattr(self.model, "mlm_overflow", None)
        if ovf is not None and bool(ovf):
            raise _lib.MvptrError(f"masked-LM capacity {self.model.mlm_capacity} overflowed: labels were dropped; "
                                  "rebuild GraphedTrainStep with a larger mlm_capacity")
