"""Flat HBM layout of a model's parameters.

All parameters of a model live in ONE fp32 buffer (decay-first order), with a bf16
compute copy and an fp32 gradient buffer of identical layout:

* ``nn.Parameter.data`` / ``.grad`` are views into the flat buffers, so state_dict keys,
  shapes and optimizers keep working exactly as in the reference;
* query/key/value weights (and biases) of a layer are adjacent, so the fused QKV
  projection reads ``[3H, H]`` with zero copies (reference: three nn.Linear,
  modeling_bert.py:293-295);
* weight-gradient GEMMs TMA-reduce-add straight into the gradient buffer (no
  flatten / copy kernels), NCCL all-reduces it in place, and the fused AdamW
  (csrc/loss_optim.cu) walks it in one launch.
"""
import torch

from . import _lib

NO_DECAY = ("bias", "LayerNorm.weight")  # run_pretrain_ml.py:379-387 optimizer grouping
ALIGN = 8  # elements; keeps every bf16 tensor 16-byte aligned for TMA


def _is_no_decay(name):
    return any(nd in name for nd in NO_DECAY)


class ParamArena:
    def __init__(self, model):
        named = [(n, p) for n, p in model.named_parameters()]
        if not named:
            raise ValueError("model has no parameters")
        dev, dt = named[0][1].device, named[0][1].dtype
        if dev.type != "cuda":
            raise _lib.MvptrError("mvp_pytorch_b200 runs on CUDA only (move the model with .cuda()/.to('cuda')); "
                                  "there is no CPU path")
        if dt not in (torch.float32, torch.bfloat16):
            raise _lib.MvptrError(f"parameter dtype {dt} unsupported: use float32 (bf16 compute copy is kept "
                                  "internally) or bfloat16")
        for n, p in named:
            if p.device != dev or p.dtype != dt:
                raise _lib.MvptrError(f"parameter {n} is {p.dtype} on {p.device}; all parameters must share "
                                      f"{dt} on {dev}")
        self.device, self.dtype = dev, dt
        order = [x for x in named if not _is_no_decay(x[0])] + [x for x in named if _is_no_decay(x[0])]
        self.offsets, off = {}, 0
        self.decay_end = None
        for n, p in order:
            if self.decay_end is None and _is_no_decay(n):
                self.decay_end = off
            self.offsets[n] = (off, p.numel(), tuple(p.shape))
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        if self.decay_end is None:
            self.decay_end = off
        self.numel = off
        self.master = torch.zeros(off, device=dev, dtype=dt)
        for n, p in order:
            o, k, shp = self.offsets[n]
            view = self.master[o:o + k].view(shp)
            view.copy_(p.data)
            p.data = view
        self.shadow = self.master if dt == torch.bfloat16 else torch.empty(off, device=dev, dtype=torch.bfloat16)
        self.grad = None
        # names whose gradient some backward has written at least once (sticky, like a torch 1.7 .grad that
        # stays a tensor after zero_grad()): the fused AdamW skips everything else, as the reference skips
        # parameters whose grad is None (optimization.py:140-142)
        self.touched = set()
        self.dirty = False  # set by PreTrainedModel.mark_weights_changed(): p.data writes do not bump versions
        self.params = dict(named)
        self._sentinels = [(p, p.data_ptr()) for _, p in (order[0], order[-1])]
        self._shadow_version = -1
        self.refresh_shadow(force=True)

    # -- validity -------------------------------------------------------------------
    def valid(self):
        return all(p.data_ptr() == ptr and p.dtype == self.dtype for p, ptr in self._sentinels)

    # -- bf16 compute copy ------------------------------------------------------------
    def refresh_shadow(self, force=False):
        if self.shadow is self.master:
            return
        v = self.master._version
        if force or self.dirty or v != self._shadow_version:
            _lib.call("mvptr_cast_f32_bf16", self.master, self.shadow, self.numel)
            self._shadow_version = v
            self.dirty = False

    def mark_shadow_fresh(self):
        self._shadow_version = self.master._version

    def w(self, name):
        """bf16 compute copy of a parameter."""
        o, k, shp = self.offsets[name]
        return self.shadow[o:o + k].view(shp)

    def w_span(self, first, n_tensors_rows, cols=None):
        """bf16 view starting at parameter `first` spanning `n_tensors_rows` rows (fused QKV)."""
        o, k, shp = self.offsets[first]
        if cols is None:
            return self.shadow[o:o + n_tensors_rows]
        return self.shadow[o:o + n_tensors_rows * cols].view(n_tensors_rows, cols)

    def master_of(self, name):
        o, k, shp = self.offsets[name]
        return self.master[o:o + k].view(shp)

    # -- gradients ---------------------------------------------------------------------
    def ensure_grad(self):
        if self.grad is None:
            self.grad = torch.zeros(self.numel, device=self.device, dtype=torch.float32)
            self.bind_grads()
        return self.grad

    def bind_grads(self):
        if self.dtype != torch.float32:
            return  # bf16 parameters: fp32 accumulators stay internal
        for n, p in self.params.items():
            if p.requires_grad:
                o, k, shp = self.offsets[n]
                p.grad = self.grad[o:o + k].view(shp)

    def g(self, name):
        o, k, shp = self.offsets[name]
        self.touched.add(name)
        return self.ensure_grad()[o:o + k].view(shp)

    def g_span(self, first, rows, cols=None):
        o, k, shp = self.offsets[first]
        self.touched.add(first)
        if cols is None:
            return self.ensure_grad()[o:o + rows]
        return self.ensure_grad()[o:o + rows * cols].view(rows, cols)

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()
            first = next(iter(self.params.values()))
            if first.grad is None or first.grad.data_ptr() != self.grad.data_ptr() + 4 * self.offsets[next(iter(self.params))][0]:
                self.bind_grads()  # only when something (e.g. zero_grad(set_to_none)) dropped the views
