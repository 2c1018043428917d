"""CPU: the loss switch of the classification heads (reference modeling_vlbert.py:1778-1796 / 1850-1868, helpers
:27-39 `soft_cross_entropy`, :878-883 `instance_bce_with_logits`) against the formulas written out by hand.  These
branches run on tiny fp32 [batch, classes] tensors as plain torch (no kernel), so they are checked here without a GPU:
regression (num_labels == 1), soft labels, KL against a 3129-way answer distribution, instance BCE, cross entropy."""
import math
import types

import torch

from mvp_pytorch_b200 import modeling_vlbert as MV


def _self(loss_type):
    return types.SimpleNamespace(loss_type=loss_type)


def test_regression_branch_is_mean_squared_error():
    g = torch.Generator().manual_seed(0)
    logits, labels = torch.randn(7, 1, generator=g), torch.randint(0, 5, (7,), generator=g)
    got = MV._cls_loss(_self("xe"), logits, labels, False, 1)
    want = sum((float(logits[i, 0]) - float(labels[i])) ** 2 for i in range(7)) / 7
    assert abs(float(got) - want) < 1e-6


def test_soft_label_branch_is_two_class_soft_cross_entropy():
    g = torch.Generator().manual_seed(1)
    logits, t = torch.randn(9, 2, generator=g), torch.rand(9, generator=g)
    got = MV._cls_loss(_self("xe"), logits, t, True, 2)
    want = 0.0
    for i in range(9):
        z = [float(logits[i, 0]), float(logits[i, 1])]
        lse = math.log(math.exp(z[0]) + math.exp(z[1]))
        want -= (1 - float(t[i])) * (z[0] - lse) + float(t[i]) * (z[1] - lse)
    assert abs(float(got) - want / 9) < 1e-5


def test_kl_branch_is_batchmean_kl_over_3129_answers():
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(4, 3129, generator=g)
    tgt = torch.zeros(4, 3129)
    for i in range(4):  # sparse soft answer scores, as the VQA targets are
        idx = torch.randint(0, 3129, (3,), generator=g)
        tgt[i, idx] = torch.tensor([0.6, 0.3, 0.1])
    got = MV._cls_loss(_self("kl"), logits, tgt, False, 3129)
    lq = torch.log_softmax(logits.double(), -1)
    p = tgt.double()
    want = torch.where(p > 0, p * (p.clamp_min(1e-300).log() - lq), torch.zeros_like(p)).sum() / 4
    assert abs(float(got) - float(want)) < 1e-5


def test_bce_branch_is_mean_bce_times_the_number_of_answers():
    g = torch.Generator().manual_seed(3)
    logits, tgt = torch.randn(5, 11, generator=g), torch.rand(5, 11, generator=g)
    got = MV._cls_loss(_self("bce"), logits, tgt, False, 11)
    x, y = logits.double(), tgt.double()
    elem = torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-x.abs()))
    assert abs(float(got) - float(elem.mean() * 11)) < 1e-5


def test_default_branch_is_cross_entropy():
    g = torch.Generator().manual_seed(4)
    logits, labels = torch.randn(6, 3, generator=g), torch.randint(0, 3, (6,), generator=g)
    got = MV._cls_loss(_self("xe"), logits, labels, False, 3)
    lp = torch.log_softmax(logits.double(), -1)
    want = -sum(float(lp[i, int(labels[i])]) for i in range(6)) / 6
    assert abs(float(got) - want) < 1e-6
