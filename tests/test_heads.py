"""SURVEY 8(f-3) heads against golden outputs of the REAL reference (tests/golden/heads_tiny.pt, made by
oracle/make_golden_heads.py): the visual-entailment head (BiImageBertForSequenceClassificationPlus), the 'mlp'
classifier, BiBertImgModel.forward_joint (two images) and hn_mod='sample' (multinomial hard negatives).
CPU tests pin the oracle restatement; GPU tests compare the CUDA path (bf16 tolerances of the other suites)."""
import os

import pytest
import torch

from oracle import mvptr_oracle as O
import mvptr_parity_utils as P

ENC = ("input_ids_a", "token_type_ids_a", "attention_mask_a", "input_ids_b", "token_type_ids_b", "attention_mask_b",
       "img_feats")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "heads_tiny.pt"), weights_only=False)


def _batch(cfg, g, seed):
    B, La, Lt, R = g["dims"]
    return O.synthetic_batch(cfg, B, La, Lt, R, seed=seed, ragged=True), Lt


# ------------------------------------------------------------------ CPU: oracle == reference
def test_oracle_heads_match_reference_golden(gold):
    g = gold
    cfg = O.Cfg(**dict(g["cfg"], num_labels=3, loss_type="xe"))
    sd = O.random_state_dict(cfg, "ve", seed=g["ve"]["wseed"])
    b, Lt = _batch(cfg, g, g["ve"]["bseed"])
    with torch.no_grad():
        loss, logits = O.ve_plus_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                         g["ve"]["labels"], b["input_ids_b"], b["token_type_ids_b"],
                                         b["attention_mask_b"], b["img_feats"], max_tag_length=Lt)
    assert torch.allclose(logits, g["ve"]["logits"], atol=2e-5) and torch.allclose(loss, g["ve"]["loss"], atol=2e-5)
    sd2 = O.random_state_dict(cfg, "cls_mlp", seed=g["cls_mlp"]["wseed"])
    with torch.no_grad():
        _, logits2 = O.seqcls_mlp_forward(sd2, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                          g["cls_mlp"]["labels"], b["input_ids_b"], b["token_type_ids_b"],
                                          b["attention_mask_b"], b["img_feats"], max_tag_length=Lt)
    assert torch.allclose(logits2, g["cls_mlp"]["logits"], atol=2e-5)
    cfg3 = O.Cfg(**g["cfg"])
    sd3 = O.random_state_dict(cfg3, "rep", seed=g["joint"]["wseed"])
    b1, _ = _batch(cfg3, g, g["joint"]["bseed1"])
    b2, _ = _batch(cfg3, g, g["joint"]["bseed2"])
    with torch.no_grad():
        seq, pooled = O.forward_joint(sd3, cfg3, b1["input_ids_a"], b1["token_type_ids_a"], b1["attention_mask_a"], Lt,
                                      b1["input_ids_b"], b1["token_type_ids_b"], b1["attention_mask_b"], b1["img_feats"],
                                      b2["input_ids_b"], b2["token_type_ids_b"], b2["attention_mask_b"], b2["img_feats"])
    assert torch.allclose(seq, g["joint"]["seq"], atol=2e-5) and torch.allclose(pooled, g["joint"]["pooled"], atol=2e-5)
    p1, p2 = O.negative_sampling_probs(g["sample"]["sim"], torch.tensor(g["sample"]["logit"]))
    assert torch.allclose(p1, g["sample"]["p_t2i"], atol=1e-6) and torch.allclose(p2, g["sample"]["p_i2t"], atol=1e-6)
    assert float(p1.diagonal().max()) == 0.0  # the matched pair is never drawn


# ------------------------------------------------------------------ GPU: CUDA path == reference
@pytest.mark.gpu
def test_visual_entailment_head_matches_reference_golden(gold):
    g = gold
    cfg = O.Cfg(**dict(g["cfg"], num_labels=3, loss_type="xe"))
    sd = O.random_state_dict(cfg, "ve", seed=g["ve"]["wseed"])
    b, Lt = _batch(cfg, g, g["ve"]["bseed"])
    model = P.build("BiImageBertForSequenceClassificationPlus", cfg, sd, train=True, classifier="linear")
    assert set(model.state_dict()) == set(sd)
    loss, logits = model(labels=g["ve"]["labels"].cuda(), max_tag_length=Lt, **P.to_cuda({k: b[k] for k in ENC}))[:2]
    model.zero_grad()
    loss.backward()
    P.close(logits.detach(), g["ve"]["logits"], 2e-2, 2e-2, "ve logits")
    P.close(loss.detach(), g["ve"]["loss"], 1e-2, 1e-2, "ve loss")
    params = dict(model.named_parameters())
    for k, gr in g["ve"]["grads"].items():
        rel = P.rel_l2(params[k].grad, gr)
        assert rel < 5e-2, f"grad {k}: relative L2 error {rel:.4f}"


@pytest.mark.gpu
def test_mlp_classifier_matches_reference_golden(gold):
    g = gold
    cfg = O.Cfg(**dict(g["cfg"], num_labels=3, loss_type="xe"))
    sd = O.random_state_dict(cfg, "cls_mlp", seed=g["cls_mlp"]["wseed"])
    b, Lt = _batch(cfg, g, g["cls_mlp"]["bseed"])
    model = P.build("BiImageBertForSequenceClassification", cfg, sd, classifier="mlp", cls_hidden_scale=2)
    with torch.no_grad():
        loss, logits = model(labels=g["cls_mlp"]["labels"].cuda(), max_tag_length=Lt, **P.to_cuda({k: b[k] for k in ENC}))[:2]
    P.close(logits, g["cls_mlp"]["logits"], 2e-2, 2e-2, "mlp logits")
    P.close(loss, g["cls_mlp"]["loss"], 1e-2, 1e-2, "mlp loss")


@pytest.mark.gpu
def test_forward_joint_two_images_matches_reference_golden(gold):
    g = gold
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, "rep", seed=g["joint"]["wseed"])
    b1, Lt = _batch(cfg, g, g["joint"]["bseed1"])
    b2, _ = _batch(cfg, g, g["joint"]["bseed2"])
    model = P.build("BiImageBertRep", cfg, sd)
    model.runtime(); model._adopt(model.bert, "bert.")
    c1, c2 = P.to_cuda(b1), P.to_cuda(b2)
    with torch.no_grad():
        seq, pooled = model.bert.forward_joint(
            input_ids_a=c1["input_ids_a"], token_type_ids_a=c1["token_type_ids_a"], attention_mask_a=c1["attention_mask_a"],
            max_tag_length=Lt, input_ids_b=c1["input_ids_b"], token_type_ids_b=c1["token_type_ids_b"],
            attention_mask_b=c1["attention_mask_b"], img_feats=c1["img_feats"], input_ids_b2=c2["input_ids_b"],
            token_type_ids_b2=c2["token_type_ids_b"], attention_mask_b2=c2["attention_mask_b"], img_feats2=c2["img_feats"])
    jm = torch.cat([b1["attention_mask_a"], b1["attention_mask_b"][:, Lt:], b2["attention_mask_b"][:, Lt:]], 1)
    assert seq.shape == g["joint"]["seq"].shape
    P.valid_rows_close(seq, g["joint"]["seq"], jm, 2e-2, 3e-2, "joint seq")
    P.close(pooled, g["joint"]["pooled"], 2e-2, 2e-2, "joint pooled")


@pytest.mark.gpu
def test_sampled_hard_negatives_match_reference_golden(gold):
    """hn_mod='sample': the sampling distributions are compared with the reference's; the draws themselves
    (torch.multinomial / torch.randperm) are replaced by the recorded ones, as when the golden was made."""
    import mvp_pytorch_b200.modeling_vlbert as mv
    g = gold
    s = g["sample"]
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, "rep", seed=g["joint"]["wseed"])
    b1, Lt = _batch(cfg, g, g["joint"]["bseed1"])
    model = P.build("BiImageBertRep", cfg, sd)
    model.runtime(); model._adopt(model.bert, "bert.")
    draws = [s["draw_img"].view(-1, 1).cuda(), s["draw_txt"].view(-1, 1).cuda()]
    seen = []
    orig_mn, orig_rp = torch.multinomial, torch.randperm
    try:
        torch.multinomial = lambda p, num_samples=1, **kw: (seen.append(p.clone()), draws[len(seen) - 1])[1]
        torch.randperm = lambda n, **kw: s["dice"].to(kw.get("device", "cpu"))
        with torch.no_grad():
            outs, single, hard = model.bert(encode_hn=True, hn_mod="sample", logit=torch.tensor(s["logit"], device="cuda"),
                                            max_tag_length=Lt, **P.to_cuda({k: b1[k] for k in ENC}))
    finally:
        torch.multinomial, torch.randperm = orig_mn, orig_rp
    # probabilities: softmax(14 * sim) amplifies the bf16 similarity error 14x -> compare with matching slack
    P.close(seen[0], s["p_t2i"], 0.15, 2e-2, "p(text -> image negative)")
    P.close(seen[1], s["p_i2t"], 0.15, 2e-2, "p(image -> text negative)")
    assert float(seen[0].diagonal().max()) < 1e-30
    assert torch.equal(hard[0].cpu(), s["hard_txt_index"]) and torch.equal(hard[1].cpu(), s["hard_img_index"])  # integer work
    P.close(outs[3], s["hard_pooled"], 2e-2, 2e-2, "hard pooled")
    with pytest.raises(ValueError):
        model.bert(encode_hn=True, hn_mod="sample", max_tag_length=Lt, **P.to_cuda({k: b1[k] for k in ENC}))


def test_oracle_re_head_matches_reference_golden(gold):
    g = gold["re"]
    cfg = O.Cfg(**dict(gold["cfg"], num_labels=1))
    sd = O.random_state_dict(cfg, "re", seed=g["wseed"])
    b, Lt = _batch(cfg, gold, g["bseed"])
    for name, case in g["cases"].items():
        with torch.no_grad():
            loss, logits = O.re_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"], g["labels"],
                                        b["input_ids_b"], b["token_type_ids_b"], b["attention_mask_b"], b["img_feats"],
                                        max_tag_length=Lt, **case["kw"])
        assert torch.allclose(logits, case["logits"], atol=2e-5), name
        assert torch.allclose(loss, case["loss"], atol=2e-5), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mod1", "mod2", "mod3", "mod1_mid"])
def test_referring_expression_head_matches_reference_golden(gold, name):
    g = gold["re"]
    case = g["cases"][name]
    cfg = O.Cfg(**dict(gold["cfg"], num_labels=1))
    sd = O.random_state_dict(cfg, "re", seed=g["wseed"])
    b, Lt = _batch(cfg, gold, g["bseed"])
    model = P.build("BiImageBertForRE", cfg, sd, train=True)
    assert set(model.state_dict()) == set(sd)
    loss, logits = model(labels=g["labels"].cuda(), max_tag_length=Lt, **case["kw"], **P.to_cuda({k: b[k] for k in ENC}))
    model.zero_grad()
    loss.backward()
    valid = (g["labels"] >= 0)
    P.close(logits.detach().cpu()[valid], case["logits"][valid], 3e-2, 3e-2, f"re {name} logits")  # padded regions unspecified
    P.close(loss.detach(), case["loss"], 2e-2, 1e-2, f"re {name} loss")
    params = dict(model.named_parameters())
    for k, gr in case["grads"].items():
        rel = P.rel_l2(params[k].grad, gr)
        assert rel < 6e-2, f"re {name} grad {k}: relative L2 error {rel:.4f}"
