"""Whole training step as ONE CUDA graph.

`GraphedTrainStep` captures zero_grad + forward + backward + gradient all-reduce + fused AdamW of a
model built from this package into a single `torch.cuda.CUDAGraph`.  A replay costs one launch, so
the ~500 kernel launches of a step no longer depend on how fast (or how often interrupted) the
host thread is, and the launch gaps between them disappear.

What makes the step capturable:
* dropout: the per-site seeds are baked into the graph; the graph's first node (`mvptr_step_params`, one
  1-thread kernel) reads this replay's record {lr, step, dropout epoch} from a ring in pinned host memory,
  indexed by a DEVICE-side replay counter, and publishes the epoch to every dropout site, so each replay
  draws fresh masks;
* AdamW: learning rate and step count come from the same record, so LR schedulers keep working
  (`optimizer.param_groups[0]['lr']` is read before every replay).  Because the record is selected by the
  device counter at execution time, a host that enqueues several replays ahead still gives replay n the values
  of step n (one host word copied by a memcpy node per replay would hand late replays the newest value);
  the host blocks only when it is `RING` replays ahead;
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
        # per-replay parameter ring (pinned host) + device replay counter, see mvptr_step_params
        self.ring = torch.zeros(self.RING, 4, dtype=torch.int32).pin_memory()
        self.ring_f = self.ring.view(torch.float32)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.slot_events = [None] * self.RING
        self.n = 0  # step_params kernels launched for execution so far == value the device counter will reach
        optimizer.enable_graph_mode()
        self.losses = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # >= 1: the first call also resolves the library's symbol addresses
                self._publish(optimizer._step + 1)  # opt.step() inside _step() advances to this value
                self._step()
                self._mark_launched()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()
        # the capture pass advanced the optimizer's host step without executing: undo
        optimizer._step -= 1

    RING = 16  # replays the host may run ahead of the device

    def _publish(self, step_value):
        """Fill the record of the next step_params execution (host side, before it is launched)."""
        slot = self.n % self.RING
        ev = self.slot_events[slot]
        if ev is not None:  # the replay that last read this slot must be done before it is overwritten
            ev.synchronize()
        self.ring_f[slot, 0] = float(self.opt.param_groups[0]["lr"])
        self.ring_f[slot, 1] = float(step_value)
        self.ring[slot, 2] = (self.n + 1) & 0x7FFFFFFF  # dropout epoch: a fresh value per executed step

    def _mark_launched(self):
        ev = torch.cuda.Event()
        ev.record()
        self.slot_events[self.n % self.RING] = ev
        self.n += 1

    def _step(self):
        _lib.call("mvptr_step_params", self.ring, self.RING, self.counter, self.opt._dyn)
        self.opt._dyn_external = True  # {lr, step} were just published on the device by step_params
        try:
            self.model.zero_grad()
            out = self.model(**self.static, **self.kw)
            out[0].backward()
            if self.allreduce:
                from .parallel import allreduce_gradients
                allreduce_gradients(self.model)
            self.opt.step()
        finally:
            self.opt._dyn_external = False
        self.losses = torch.stack([o.detach().float() for o in out])

    def load(self, batch, non_blocking=True):
        """Refill the static input buffers (device-to-device or host-to-device copies)."""
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=non_blocking)

    def __call__(self, batch=None):
        if batch is not None:
            self.load(batch)
        self.opt._step += 1
        self._publish(self.opt._step)
        self.graph.replay()
        self._mark_launched()
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
