"""GPU: the fused AdamW at MODEL level -- the reference call `AdamW(grouped_params, lr=, eps=)` without a
`model=` argument (run_retrieval.py:567, run_vqa.py:558, run_pretrain_ml.py:389), `optimizer.zero_grad()`
(run_pretrain_ml.py:644), frozen backbones (run_ve.py:479), unused heads, and `state_dict()` save / resume
(run_pretrain_ml.py:725).  The checker is the oracle's per-tensor restatement of optimization.py:130-189."""
import copy

import pytest
import torch

from oracle import mvptr_oracle as O
import mvptr_parity_utils as P

pytestmark = pytest.mark.gpu

TINY = dict(vocab_size=1500, only_word_size=1000, hidden_size=128, num_hidden_layers=4, num_attention_heads=2,
            intermediate_size=256, max_position_embeddings=64, img_feature_dim=70, qa_answer_size=37, num_labels=2)
DIMS = (6, 12, 5, 9)


def _grouped(model, wd=0.01):
    """The grouping every reference script builds (run_pretrain_ml.py:379-387)."""
    no_decay = ["bias", "LayerNorm.weight"]
    return [
        {"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": wd},
        {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
    ]


def _retrieval_model(seed=2, train=True):
    cfg = O.Cfg(**TINY)
    sd = O.random_state_dict(cfg, "retrieval", seed=seed)
    model = P.build("BiImageBertForRetrieval", cfg, sd, train=train)
    B, La, Lt, R = DIMS
    batch = P.to_cuda(O.synthetic_batch(cfg, B, La, Lt, R, seed=12, ragged=True))
    return cfg, sd, model, batch


def _step(model, batch, opt, zero):
    zero()
    torch.manual_seed(5)  # same randperm draw for every run compared
    out = model(max_tag_length=DIMS[2], **batch)
    out[0].backward()
    opt.step()
    return float(out[0])


def test_reference_style_constructor_finds_its_model_and_matches_per_tensor_adamw():
    from mvp_pytorch_b200.optimization import AdamW
    cfg, sd, model, batch = _retrieval_model()
    opt = AdamW(_grouped(model), lr=1e-3, eps=1e-8)  # exactly the reference's call: no model=
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    _step(model, batch, opt, model.zero_grad)
    assert opt.model is model
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    worst = 0.0
    for k, p in model.named_parameters():
        ref = before[k].clone().cpu()
        wd = 0.0 if ("bias" in k or "LayerNorm.weight" in k) else 0.01
        O.adamw_step(ref, grads[k].cpu(), torch.zeros_like(ref), torch.zeros_like(ref), 1, 1e-3, eps=1e-8, weight_decay=wd)
        err = float((p.detach().cpu() - ref).abs().max())
        worst = max(worst, err)
        assert err <= 1e-6 + 1e-5 * float(ref.abs().max()), f"{k}: {err}"
    print(f"[parity] fused AdamW vs per-tensor reference update: max |d| = {worst:.2e}")


def test_optimizer_zero_grad_equals_model_zero_grad():
    """optimizer.zero_grad() must zero the flat gradient arena, not just drop the views."""
    from mvp_pytorch_b200.optimization import AdamW
    runs = []
    for use_opt_zero in (False, True):
        cfg, sd, model, batch = _retrieval_model()
        opt = AdamW(_grouped(model), lr=1e-3, eps=1e-8)
        zero = opt.zero_grad if use_opt_zero else model.zero_grad
        losses = [_step(model, batch, opt, zero) for _ in range(3)]
        runs.append((losses, {k: v.detach().clone() for k, v in model.named_parameters()}))
        assert all(p.grad is not None for p in model.parameters())  # views stay bound for clip_grad_norm_
    # (fp32 atomics in the column-sum reductions may reorder between runs: compare to rounding noise, not bits)
    assert max(abs(a - b) for a, b in zip(*[r[0] for r in runs])) < 1e-5
    for k in runs[0][1]:
        assert P.rel_l2(runs[0][1][k], runs[1][1][k].cpu()) < 1e-3, k


def test_frozen_and_unused_parameters_are_neither_stepped_nor_decayed():
    """freeze_backbone() (run_ve.py:479) and heads that never receive a gradient
    (qa_head without qa_ans) keep their values bit for bit, as under the reference's `if p.grad is None`."""
    from mvp_pytorch_b200.optimization import AdamW
    cfg = O.Cfg(**dict(TINY, num_labels=3, loss_type="xe"))
    sd = O.random_state_dict(cfg, "cls_mlp", seed=7)
    model = P.build("BiImageBertForSequenceClassification", cfg, sd, train=True, classifier="mlp", cls_hidden_scale=2)
    B, La, Lt, R = DIMS
    batch = P.to_cuda(O.synthetic_batch(cfg, B, La, Lt, R, seed=16, ragged=True))
    labels = torch.randint(0, 3, (B,), generator=torch.Generator().manual_seed(3)).cuda()
    model.freeze_backbone()
    opt = AdamW(_grouped(model, wd=0.1), lr=1e-2)
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    for _ in range(2):
        opt.zero_grad()
        model(labels=labels, max_tag_length=Lt, **batch)[0].backward()
        opt.step()
    for k, p in model.named_parameters():
        if k.startswith("bert."):
            assert torch.equal(p.detach(), before[k]), f"frozen {k} changed"
        else:
            assert not torch.equal(p.detach(), before[k]), f"trainable {k} did not move"
    # the bf16 compute copy of the frozen part is still the cast of the (unchanged) master
    a = model.runtime().arena
    assert torch.equal(a.shadow.float(), a.master.to(torch.bfloat16).float())

    # pre-training model without qa_ans: qa_head never gets a gradient -> untouched (no decay)
    cfg2 = O.Cfg(**TINY)
    sd2 = O.random_state_dict(cfg2, "pretrain", seed=1)
    b2 = O.synthetic_batch(cfg2, B, La, Lt, R, seed=11, ragged=True, with_labels=True)
    m2 = P.build("BiBertImgForPreTraining", cfg2, sd2, train=True, max_text_seq_length=La)
    cb = P.to_cuda(b2)
    opt2 = AdamW(_grouped(m2, wd=0.1), lr=1e-2)
    qa0 = m2.qa_head.weight.detach().clone()
    opt2.zero_grad()
    m2(input_ids_a=cb["input_ids_a"], token_type_ids_a=cb["token_type_ids_a"], attention_mask_a=cb["attention_mask_a"],
       masked_lm_labels_a=cb["masked_lm_labels_a"], input_ids_b=cb["input_ids_b"], token_type_ids_b=cb["token_type_ids_b"],
       attention_mask_b=cb["attention_mask_b"], masked_lm_labels_b=cb["masked_lm_labels_b"], img_feats=cb["img_feats"],
       max_tag_length=Lt)[0].backward()
    opt2.step()
    assert torch.equal(m2.qa_head.weight.detach(), qa0)
    assert not torch.equal(m2.cls.seq_relationship.weight.detach().cpu(), sd2["cls.seq_relationship.weight"])


def test_groups_with_different_lr_raise():
    from mvp_pytorch_b200 import _lib
    from mvp_pytorch_b200.optimization import AdamW
    cfg, sd, model, batch = _retrieval_model()
    groups = _grouped(model)
    groups[1]["lr"] = 5e-4
    opt = AdamW(groups, lr=1e-3)
    model.zero_grad()
    model(max_tag_length=DIMS[2], **batch)[0].backward()
    with pytest.raises(_lib.MvptrError):
        opt.step()


def test_state_dict_save_resume_continues_the_trajectory():
    """optimizer.state_dict() (run_pretrain_ml.py:725) carries exp_avg / exp_avg_sq / step in
    the reference's per-parameter layout; a resumed run reproduces the uninterrupted one (to fp32 atomic-order noise)."""
    from mvp_pytorch_b200.optimization import AdamW
    cfg, sd, model, batch = _retrieval_model()
    opt = AdamW(_grouped(model), lr=1e-3, eps=1e-8)
    for _ in range(2):
        _step(model, batch, opt, opt.zero_grad)
    ck_model = copy.deepcopy({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    ck_opt = opt.state_dict()
    st = ck_opt["state"]
    assert len(st) == len(list(model.parameters()))
    assert set(st[0]) == {"step", "exp_avg", "exp_avg_sq"} and st[0]["step"] == 2
    assert st[0]["exp_avg"].shape == opt.param_groups[0]["params"][0].shape
    cont = [_step(model, batch, opt, opt.zero_grad) for _ in range(2)]
    final = {k: v.detach().clone() for k, v in model.named_parameters()}

    cfg, _, model2, batch2 = _retrieval_model()
    model2.load_state_dict(ck_model, strict=True)
    opt2 = AdamW(_grouped(model2), lr=1e-3, eps=1e-8)
    opt2.load_state_dict(ck_opt)
    resumed = [_step(model2, batch2, opt2, opt2.zero_grad) for _ in range(2)]
    assert max(abs(a - b) for a, b in zip(resumed, cont)) < 1e-5
    for k, p in model2.named_parameters():
        assert P.rel_l2(p.detach(), final[k].cpu()) < 1e-3, k


def test_qa_loss_ignores_minus_one_labels():
    """CrossEntropyLoss(ignore_index=-1) of the QA head (modeling_vlbert.py:1262-1264) for a
    small answer set -- the path that used to index out of bounds -- against torch on the same logits."""
    import torch.nn.functional as F
    cfg = O.Cfg(**TINY)
    sd = O.random_state_dict(cfg, "pretrain", seed=1)
    B, La, Lt, R = DIMS
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=11, ragged=True, with_labels=True)
    model = P.build("BiBertImgForPreTraining", cfg, sd, train=True, max_text_seq_length=La)
    cb = P.to_cuda(b)
    qa = torch.tensor([3, -1, 0, 36, -1, 7]).cuda()
    kw = dict(input_ids_a=cb["input_ids_a"], token_type_ids_a=cb["token_type_ids_a"], attention_mask_a=cb["attention_mask_a"],
              masked_lm_labels_a=cb["masked_lm_labels_a"], input_ids_b=cb["input_ids_b"],
              token_type_ids_b=cb["token_type_ids_b"], attention_mask_b=cb["attention_mask_b"],
              masked_lm_labels_b=cb["masked_lm_labels_b"], img_feats=cb["img_feats"], max_tag_length=Lt)
    torch.manual_seed(3)
    out = model(qa_ans=qa, **kw)
    assert len(out) == 6  # (total, vis_mlm, vsc, mlm, itm, qa)
    model.zero_grad()
    out[0].backward()
    g_qa = model.qa_head.weight.grad.clone()
    # the same pooled vectors through torch: rerun without qa_ans to fetch them
    torch.manual_seed(3)
    outs, _, _ = model.bert(input_ids_a=kw["input_ids_a"], token_type_ids_a=kw["token_type_ids_a"],
                            attention_mask_a=kw["attention_mask_a"], input_ids_b=kw["input_ids_b"],
                            token_type_ids_b=kw["token_type_ids_b"], attention_mask_b=kw["attention_mask_b"],
                            img_feats=kw["img_feats"], max_tag_length=Lt, encode_hn=True)
    pooled = outs[1].detach().float()
    a = model.runtime().arena
    w = a.w("qa_head.weight").float().requires_grad_(True)
    ref = F.cross_entropy(pooled @ w.t() + a.w("qa_head.bias").float(), qa, ignore_index=-1)
    ref.backward()
    assert abs(float(out[5]) - float(ref)) < 2e-3 * max(1.0, abs(float(ref)))
    assert P.rel_l2(g_qa, w.grad.cpu()) < 1e-2
    # the two ignored rows contribute no gradient: recompute with them dropped
    assert torch.isfinite(out[0])
