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
    """Same constructor and semantics as the reference AdamW (optimization.py:107-189), run as fused launches
    over the owning model's flat parameter arena.

    ``AdamW(model.parameters() | grouped_parameters, lr=..., eps=...)`` -- the reference call
    (run_retrieval.py:567, run_vqa.py:558, run_pretrain_ml.py:389) -- works unchanged: the optimizer finds the
    mvp_pytorch_b200 model that owns its parameters.  ``model=`` / ``AdamW.for_model`` name it explicitly.

    What is updated follows the reference: only parameters that were handed to the optimizer, have
    ``requires_grad`` and have received a gradient (the reference skips ``p.grad is None``, :140-142) -- a frozen
    backbone (``freeze_backbone()``, run_ve.py:479) or a head that a run never uses (``qa_head`` without
    ``qa_ans``) is neither stepped nor decayed.  Each parameter group's ``weight_decay`` applies to its own
    tensors; the groups must share one ``lr`` / ``betas`` / ``eps`` (they do in every reference script)."""

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
        self._plan_key = self._plan = None

    @classmethod
    def for_model(cls, model, **kw):
        return cls(model.parameters(), model=model, **kw)

    # ---- owning model / arena -------------------------------------------------------------------------
    def _find_model(self):
        """The mvp_pytorch_b200 model whose parameters these are (the outermost one that holds them all)."""
        from .modeling_utils import live_models
        mine = {id(p) for g in self.param_groups for p in g["params"]}
        best, best_n = None, -1
        for m in live_models():
            ids = {id(p) for p in m.parameters()}
            if mine <= ids and len(ids) > best_n:
                best, best_n = m, len(ids)
        if best is None:
            raise _lib.MvptrError("fused AdamW: the parameters do not all belong to one mvp_pytorch_b200 model "
                                  "(pass model=... or use the reference's per-tensor AdamW)")
        return best

    def _arena(self):
        if self.model is None:
            self.model = self._find_model()
        rt = self.model.runtime()
        a = rt.arena
        if a.dtype != torch.float32:
            raise _lib.MvptrError("fused AdamW keeps fp32 master weights: keep the model in float32 "
                                  "(the bf16 compute copy is internal)")
        rt.shadow_managed = True
        return a

    def _moments(self, a):
        if self._m is None or self._m.numel() != a.numel or self._m.device != a.device:
            self._m = torch.zeros_like(a.master)
            self._v = torch.zeros_like(a.master)
            self._norm = torch.zeros(1, device=a.device, dtype=torch.float32)

    def _segments(self, a):
        """[(lo, hi, decay_end, weight_decay)]: maximal runs of arena elements that are stepped, each run
        decaying its first decay_end - lo elements.  With every parameter trainable and the reference's
        decay / no-decay grouping this is ONE run over the whole arena."""
        by_param = {}
        for g in self.param_groups:
            for p in g["params"]:
                by_param[id(p)] = g
        key = (len(a.touched), tuple(p.requires_grad for p in a.params.values()),
               tuple(float(g["weight_decay"]) for g in self.param_groups))
        if key == self._plan_key:
            return self._plan
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            for k in ("lr", "betas", "eps", "correct_bias"):
                if g[k] != g0[k]:
                    raise _lib.MvptrError(f"fused AdamW: parameter groups carry different {k} ({g0[k]} vs {g[k]}); "
                                          "only weight_decay may differ between groups")
        runs = []  # (lo, hi, wd)
        for name, (off, numel, _) in a.offsets.items():
            p = a.params[name]
            g = by_param.get(id(p))
            if g is None or not p.requires_grad or name not in a.touched:
                continue
            hi = off + (numel + 7) // 8 * 8
            wd = float(g["weight_decay"])
            if runs and runs[-1][1] == off and runs[-1][2] == wd:
                runs[-1] = (runs[-1][0], hi, wd)
            else:
                runs.append((off, hi, wd))
        segs = []
        for lo, hi, wd in runs:
            if segs and segs[-1][1] == lo and segs[-1][2] == segs[-1][1] and segs[-1][3] > 0 and wd == 0:
                segs[-1] = (segs[-1][0], hi, segs[-1][2], segs[-1][3])  # decayed run followed by an undecayed one
            else:
                segs.append((lo, hi, hi if wd > 0 else lo, wd))
        self._plan_key, self._plan = key, segs
        return segs

    def enable_graph_mode(self):
        """lr and step count are read from device memory at execution time (see graphs.py)."""
        a = self._arena()
        self._moments(a)
        self._dyn_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._dyn = torch.zeros(2, device=a.device, dtype=torch.float32)

    def zero_grad(self, set_to_none=False):
        """run_pretrain_ml.py:644 calls optimizer.zero_grad(): gradients live in the flat arena (the kernels
        reduce-add into it), so it is zeroed in place and the parameters' .grad views stay bound."""
        try:
            a = self._arena()
        except _lib.MvptrError:
            return super().zero_grad(set_to_none=set_to_none)
        a.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        a = self._arena()
        g = a.ensure_grad()
        self._moments(a)
        self._step += 1
        group = self.param_groups[0]
        segs = self._segments(a)
        dyn = None
        if self._dyn is not None:  # graph mode: the kernel reads {lr, step} from device memory at execution time
            if not self._dyn_external:  # eager step of a graph-mode optimizer: publish them from the host
                self._dyn_host[0], self._dyn_host[1] = float(group["lr"]), float(self._step)
                self._dyn.copy_(self._dyn_host, non_blocking=True)
            dyn = self._dyn
        norm = None
        if self.max_grad_norm > 0:
            self._norm.zero_()
            for lo, hi, _, _ in segs:
                _lib.call("mvptr_sumsq", g[lo:hi], hi - lo, self._norm)
            norm = self._norm
        b1, b2 = group["betas"]
        sh = None if a.shadow is a.master else a.shadow
        for lo, hi, dec, wd in segs:
            _lib.call("mvptr_adamw", a.master[lo:hi], g[lo:hi], self._m[lo:hi], self._v[lo:hi],
                      None if sh is None else sh[lo:hi], hi - lo, dec - lo, float(group["lr"]), float(b1), float(b2),
                      float(group["eps"]), float(wd), self._step, int(bool(group["correct_bias"])), norm,
                      self.max_grad_norm, dyn)
        a.mark_shadow_fresh()
        return loss

    # ---- checkpointing (run_pretrain_ml.py:725 saves optimizer.state_dict()) ---------------------------------
    def state_dict(self):
        """The reference layout: state[i] = {step, exp_avg, exp_avg_sq} per parameter (optimization.py:147-152),
        cut out of the flat moment buffers."""
        sd = super().state_dict()
        if self._m is None:
            return sd
        a = self._arena()
        name_of = {id(p): n for n, p in a.params.items()}
        state, i = {}, 0
        for g in self.param_groups:
            for p in g["params"]:
                n = name_of.get(id(p))
                if n is not None:
                    off, numel, shp = a.offsets[n]
                    state[i] = {"step": self._step, "exp_avg": self._m[off:off + numel].view(shp).clone(),
                                "exp_avg_sq": self._v[off:off + numel].view(shp).clone()}
                i += 1
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        state = state_dict.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": state_dict["param_groups"]})
        if not state:
            return
        a = self._arena()
        self._moments(a)
        name_of = {id(p): n for n, p in a.params.items()}
        i = 0
        with torch.no_grad():
            for g in self.param_groups:
                for p in g["params"]:
                    st = state.get(i, state.get(str(i)))
                    n = name_of.get(id(p))
                    if st is not None and n is not None:
                        off, numel, shp = a.offsets[n]
                        self._m[off:off + numel].view(shp).copy_(st["exp_avg"])
                        self._v[off:off + numel].view(shp).copy_(st["exp_avg_sq"])
                        self._step = max(self._step, int(st["step"]))
                        a.touched.add(n)  # it had optimizer state, i.e. it had been stepped before
                    i += 1
