"""AdamW + LR schedules with the reference's interface
(transformers/pytorch_transformers/optimization.py:26-103 schedules, :107-189 AdamW).

``AdamW.step`` is ONE fused CUDA launch over the model's flat parameter arena
(csrc/loss_optim.cu: adamw_kernel) instead of the reference's per-tensor Python loop
(~1800 launches/step).  It also refreshes the bf16 compute copy of the
weights, so no separate cast pass is needed in training.  Gradient clipping
(``clip_grad_norm_`` of run_retrieval.py:636 / DeepSpeed ``gradient_clipping``,
tmp_config.json:26) is folded in through ``max_grad_norm``.
"""
import math

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from . import _lib


class ConstantLRSchedule(LambdaLR):
    def __init__(self, optimizer, last_epoch=-1):
        super().__init__(optimizer, lambda _: 1.0, last_epoch=last_epoch)


class WarmupConstantSchedule(LambdaLR):
    def __init__(self, optimizer, warmup_steps, last_epoch=-1):
        self.warmup_steps = warmup_steps
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1.0, self.warmup_steps))
        return 1.


class WarmupLinearSchedule(LambdaLR):
    """Linear warmup then linear decay to 0 at t_total (optimization.py:55-68)."""

    def __init__(self, optimizer, warmup_steps, t_total, last_epoch=-1):
        self.warmup_steps = warmup_steps
        self.t_total = t_total
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1, self.warmup_steps))
        return max(0.0, float(self.t_total - step) / float(max(1.0, self.t_total - self.warmup_steps)))


class WarmupCosineSchedule(LambdaLR):
    def __init__(self, optimizer, warmup_steps, t_total, cycles=.5, last_epoch=-1):
        self.warmup_steps = warmup_steps
        self.t_total = t_total
        self.cycles = cycles
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1.0, self.warmup_steps))
        progress = float(step - self.warmup_steps) / float(max(1, self.t_total - self.warmup_steps))
        return max(0.0, 0.5 * (1. + math.cos(math.pi * float(self.cycles) * 2.0 * progress)))


class AdamW(Optimizer):
    """Same constructor as the reference AdamW.  Construct it either from
    ``model.parameters()`` / parameter groups (reference style; every parameter must belong
    to ONE mvp_pytorch_b200 model whose arena is then used) or with ``AdamW.for_model``.

    Weight decay follows the arena layout: tensors whose name contains ``bias`` or
    ``LayerNorm.weight`` get none (the grouping of run_pretrain_ml.py:379-387); the decay
    value is taken from the first group that has one."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 max_grad_norm=0.0, model=None):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self.model = model
        self.max_grad_norm = float(max_grad_norm)
        self._step = 0
        self._m = self._v = None
        self._norm = None
        self._dyn = None       # device float[2] {lr, step} for CUDA-graph replays
        self._dyn_host = None  # its pinned host source
        self._dyn_external = False  # True while graphs.GraphedTrainStep publishes _dyn itself (mvptr_step_params)

    @classmethod
    def for_model(cls, model, **kw):
        return cls(model.parameters(), model=model, **kw)

    def _arena(self):
        if self.model is None:
            raise _lib.MvptrError("fused AdamW needs the owning model: AdamW(params, ..., model=model)")
        rt = self.model.runtime()
        a = rt.arena
        if a.dtype != torch.float32:
            raise _lib.MvptrError("fused AdamW keeps fp32 master weights: keep the model in float32 "
                                  "(the bf16 compute copy is internal)")
        rt.shadow_managed = True
        return a

    def enable_graph_mode(self):
        """lr and step count are read from device memory at execution time (see graphs.py)."""
        a = self._arena()
        if self._m is None:
            self._m = torch.zeros_like(a.master)
            self._v = torch.zeros_like(a.master)
            self._norm = torch.zeros(1, device=a.device, dtype=torch.float32)
        self._dyn_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._dyn = torch.zeros(2, device=a.device, dtype=torch.float32)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        a = self._arena()
        g = a.ensure_grad()
        if self._m is None:
            self._m = torch.zeros_like(a.master)
            self._v = torch.zeros_like(a.master)
            self._norm = torch.zeros(1, device=a.device, dtype=torch.float32)
        self._step += 1
        group = self.param_groups[0]
        dyn = None
        if self._dyn is not None:  # graph mode: the kernel reads {lr, step} from device memory at execution time
            if not self._dyn_external:  # eager step of a graph-mode optimizer: publish them from the host
                self._dyn_host[0], self._dyn_host[1] = float(group["lr"]), float(self._step)
                self._dyn.copy_(self._dyn_host, non_blocking=True)
            dyn = self._dyn
        wd = max(gr["weight_decay"] for gr in self.param_groups)
        norm = None
        if self.max_grad_norm > 0:
            self._norm.zero_()
            _lib.call("mvptr_sumsq", g, a.numel, self._norm)
            norm = self._norm
        b1, b2 = group["betas"]
        _lib.call("mvptr_adamw", a.master, g, self._m, self._v, None if a.shadow is a.master else a.shadow,
                  a.numel, a.decay_end, float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(wd),
                  self._step, int(bool(group["correct_bias"])), norm, self.max_grad_norm, dyn)
        a.mark_shadow_fresh()
        return loss
