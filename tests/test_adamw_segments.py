"""CPU: which arena ranges the fused AdamW steps and decays (optimization.AdamW._segments) -- the reference rule
(optimization.py:130-189): only parameters handed to the optimizer, with requires_grad, that received a gradient;
each group's weight_decay on its own tensors; one lr / betas / eps for all groups.  Pure host logic over a stand-in
arena (offsets in the real arena's decay-first order, every tensor padded to 8 elements)."""
import collections
import types

import pytest
import torch

from mvp_pytorch_b200 import _lib
from mvp_pytorch_b200.optimization import AdamW


def _arena(sizes):
    params = collections.OrderedDict((n, torch.nn.Parameter(torch.zeros(k))) for n, k in sizes)
    offsets, off = collections.OrderedDict(), 0
    for n, k in sizes:
        offsets[n] = (off, k, (k,))
        off += (k + 7) // 8 * 8
    return types.SimpleNamespace(params=params, offsets=offsets, touched=set(params), numel=off)


SIZES = [("enc.w", 64), ("head.w", 24), ("enc.b", 8), ("enc.ln", 5), ("head.b", 3)]  # weights first, then the no-decay tensors


def _opt(a, **kw):
    decay = [a.params[n] for n in ("enc.w", "head.w")]
    nodecay = [a.params[n] for n in ("enc.b", "enc.ln", "head.b")]
    return AdamW([{"params": decay, "weight_decay": 0.01}, {"params": nodecay, "weight_decay": 0.0}], lr=1e-4, eps=1e-8, **kw)


def test_everything_trainable_is_one_run_decaying_its_weight_prefix():
    a = _arena(SIZES)
    assert _opt(a)._segments(a) == [(0, a.numel, 88, 0.01)]  # 64 + 24 decayed, then 8 + 8 + 8 padded no-decay elements


def test_frozen_backbone_is_neither_stepped_nor_decayed():
    a = _arena(SIZES)
    for n in ("enc.w", "enc.b", "enc.ln"):
        a.params[n].requires_grad_(False)
    assert _opt(a)._segments(a) == [(64, 88, 88, 0.01), (104, 112, 104, 0.0)]  # head.w decayed; head.b stepped, not decayed


def test_a_head_that_never_received_a_gradient_is_skipped_like_grad_none():
    a = _arena(SIZES)
    a.touched -= {"head.w", "head.b"}
    assert _opt(a)._segments(a) == [(0, 64, 64, 0.01), (88, 104, 88, 0.0)]
    a.touched |= {"head.w", "head.b"}  # the head is used later in the run: the plan is rebuilt
    opt = _opt(a)
    a.touched -= {"head.w"}
    first = opt._segments(a)
    a.touched |= {"head.w"}
    assert opt._segments(a) != first and opt._segments(a) == [(0, a.numel, 88, 0.01)]


def test_parameters_not_given_to_the_optimizer_are_left_alone():
    a = _arena(SIZES)
    opt = AdamW([a.params["head.w"], a.params["head.b"]], lr=1e-3, weight_decay=0.05)
    assert opt._segments(a) == [(64, 88, 88, 0.05), (104, 112, 112, 0.05)]


def test_groups_with_different_learning_rates_are_refused():
    a = _arena(SIZES)
    opt = AdamW([{"params": [a.params["enc.w"]], "lr": 1e-3}, {"params": [a.params["head.w"]], "lr": 1e-4}], lr=1e-3)
    with pytest.raises(_lib.MvptrError, match="different lr"):
        opt._segments(a)
