"""CPU: the oracle restatement (oracle/mvptr_oracle.py) reproduces the outputs the
REAL reference produced in the authoring container (tests/golden/*.pt, written
by oracle/make_golden.py)."""
import os

import torch

from oracle import mvptr_oracle as O

TOL = 2e-5


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _close(a, b, tol=TOL):
    err = (a.double() - b.double()).abs().max().item()
    assert err <= tol * max(1.0, b.double().abs().max().item()), err


def _sum(ts):
    return sum(float(t.double().abs().sum()) for t in ts)


def test_rep_matches_reference(golden_dir):
    g = _load(golden_dir, "rep_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, g["head"], seed=g["wseed"])
    assert abs(_sum(sd.values()) - g["wsum"]) < 1e-6 * g["wsum"], "weight generator drifted"
    B, La, Lt, R = g["dims"]
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    assert abs(_sum([batch["img_feats"], batch["input_ids_a"]]) - g["bsum"]) < 1e-6 * g["bsum"]
    with torch.no_grad():
        seq, pooled, (txt, vis) = O.rep_forward(sd, cfg, max_tag_length=Lt, **batch)
    _close(seq, g["seq"]); _close(pooled, g["pooled"]); _close(txt, g["txt"]); _close(vis, g["vis"])


def test_retrieval_matches_reference(golden_dir):
    g = _load(golden_dir, "retrieval_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, g["head"], seed=g["wseed"])
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    with torch.no_grad():
        gt, gi = O.forward_single(sd, cfg, **b)
        fine = O.retrieval_fine_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                        max_tag_length=Lt, input_ids_b=b["input_ids_b"],
                                        token_type_ids_b=b["token_type_ids_b"],
                                        attention_mask_b=b["attention_mask_b"], img_feats=b["img_feats"])
        total, logits, vsc, itm, labels = O.retrieval_train_forward(
            sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"], b["input_ids_b"],
            b["token_type_ids_b"], b["attention_mask_b"], b["img_feats"], max_tag_length=Lt, dice_index=g["dice"])
    _close(gt, g["global_txt"]); _close(gi, g["global_img"]); _close(fine, g["fine_logits"])
    _close(total, g["train_total"]); _close(logits, g["train_logits"])
    assert torch.equal(labels, g["train_labels"])
    # caching stage 1 and re-running only stage 2 == forward_fine
    with torch.no_grad():
        txt, vis, ma, mb = O.stage1(sd, cfg, **b)
        joint = torch.cat([txt, vis[:, Lt:]], 1)
        jm = torch.cat([ma, mb[..., Lt:]], -1)
        seq, _ = O.encoder(sd, "bert.mul_encoder", joint, jm, cfg.num_hidden_layers // 2,
                           cfg.num_attention_heads, cfg.layer_norm_eps)
        again = O.linear(O.pooler(sd, "bert.pooler", seq), sd, "classifier")
    assert torch.equal(again, fine)


def test_pretrain_losses_and_grads_match_reference(golden_dir):
    g = _load(golden_dir, "pretrain_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = {k: v.requires_grad_(True) for k, v in O.random_state_dict(cfg, g["head"], seed=g["wseed"]).items()}
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True, with_labels=True)
    losses = O.pretrain_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                b["masked_lm_labels_a"], b["input_ids_b"], b["token_type_ids_b"],
                                b["attention_mask_b"], b["masked_lm_labels_b"], b["img_feats"], max_tag_length=Lt,
                                img_index=b["img_index"], phrase_index=b["phrase_index"],
                                dice_index=b["dice_index"], neg_img=b["neg_img"], rand_pos=b["rand_pos"],
                                rand_neg=b["rand_neg"])
    assert len(losses) == 6  # run_pretrain_ml.py:536 unpacks exactly six
    for a, r in zip(losses, g["losses"]):
        _close(a.detach(), r)
    losses[0].backward()
    for k, gr in g["grads"].items():
        _close(sd[k].grad, gr, tol=5e-5)
    for k, n in g["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - n) <= 1e-4 * max(1.0, n), k
    for k in g["no_grad"]:
        assert sd[k].grad is None


def test_vqa_matches_reference(golden_dir):
    g = _load(golden_dir, "vqa_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = {k: v.requires_grad_(True) for k, v in O.random_state_dict(cfg, g["head"], seed=g["wseed"]).items()}
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    loss, logits = O.vqa_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                 g["labels"], b["input_ids_b"], b["token_type_ids_b"], b["attention_mask_b"],
                                 b["img_feats"], max_tag_length=Lt)
    _close(loss.detach(), g["loss"]); _close(logits.detach(), g["logits"])
    loss.backward()
    for k, gr in g["grads"].items():
        _close(sd[k].grad, gr, tol=5e-5)


def test_rank_order_matches_numpy_argsort_reversed(golden_dir):
    g = _load(golden_dir, "rank_order.pt")
    assert torch.equal(O.topk_desc(g["sims"], g["sims"].shape[1]), g["order"])
    # documented tie rule: equal scores -> larger index first
    t = torch.tensor([[1.0, 3.0, 3.0, 0.5, 3.0]])
    assert O.topk_desc(t, 5).tolist() == [[4, 2, 1, 0, 3]]


def test_adamw_known_answers(golden_dir):
    g = _load(golden_dir, "adamw_traj.pt")
    w = torch.tensor([0.1, -0.2, -0.1, 0.7])
    m, v = torch.zeros(4), torch.zeros(4)
    for it, lr in enumerate([0.0] + g["lrs"][:-1]):
        pass
    # WarmupLinear(warmup=2,t_total=10) on base lr 0.02: lr used at step i is schedule(i)
    used = [0.02 * x for x in (0.0, 0.5, 1.0, 0.875, 0.75)]
    for it in range(5):
        grad = (w - torch.tensor([0.4, 0.2, -0.5, 0.1])) * 2
        O.adamw_step(w, grad, m, v, it + 1, used[it], weight_decay=0.01)
        _close(w, g["traj"][it], tol=1e-6)


def test_rep_base_shape_matches_reference(golden_dir):
    """BASELINE.json configs[0] at the REAL base shape (hidden 768, 12 heads, 18 layers, 2054-d regions, batch 8 x
    35 tokens + 50 regions): the oracle against what the unmodified reference produced
    (oracle/make_golden_base.py: pooled vectors in full, sampled valid rows, whole-tensor checksums)."""
    g = _load(golden_dir, "rep_base.pt")
    cfg = O.Cfg()
    sd = O.random_state_dict(cfg, g["head"], seed=g["wseed"])
    assert abs(_sum(sd.values()) - g["wsum"]) < 1e-6 * g["wsum"], "weight generator drifted"
    B, La, Lt, R = g["dims"]
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    assert abs(_sum([batch["img_feats"], batch["input_ids_a"]]) - g["bsum"]) < 1e-6 * g["bsum"]
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        seq, pooled, (txt, vis) = O.rep_forward(sd, cfg, max_tag_length=Lt, **batch)
    take = lambda t, r: torch.gather(t, 1, r[:, :, None].expand(-1, -1, t.shape[2]))
    _close(pooled, g["pooled"])
    _close(take(seq, g["rows"]["seq"]), g["seq_rows"])
    _close(take(txt, g["rows"]["txt"]), g["txt_rows"])
    _close(take(vis, g["rows"]["vis"]), g["vis_rows"])
    jm = torch.cat([batch["attention_mask_a"], batch["attention_mask_b"][:, Lt:]], 1).bool()
    for name, t, m in (("seq", seq, jm), ("txt", txt, batch["attention_mask_a"].bool()),
                       ("vis", vis, batch["attention_mask_b"].bool())):
        s = float(t[m].double().abs().sum())
        assert abs(s - g["valid_abs_sum"][name]) < 1e-5 * g["valid_abs_sum"][name], name


def test_pretrain_base_shape_losses_match_reference(golden_dir):
    """BASELINE.json configs[1] per-pair shape at the base model size (batch 6): the six pre-training losses of the
    oracle against the unmodified reference's (oracle/make_golden_base.py pretrain)."""
    g = _load(golden_dir, "pretrain_base.pt")
    cfg = O.Cfg()
    sd = O.random_state_dict(cfg, g["head"], seed=g["wseed"])
    assert abs(_sum(sd.values()) - g["wsum"]) < 1e-6 * g["wsum"], "weight generator drifted"
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True, with_labels=True)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        losses = O.pretrain_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                    b["masked_lm_labels_a"], b["input_ids_b"], b["token_type_ids_b"],
                                    b["attention_mask_b"], b["masked_lm_labels_b"], b["img_feats"], max_tag_length=Lt,
                                    img_index=b["img_index"], phrase_index=b["phrase_index"],
                                    dice_index=b["dice_index"], neg_img=b["neg_img"], rand_pos=b["rand_pos"],
                                    rand_neg=b["rand_neg"])
    assert len(losses) == len(g["losses"]) == 6
    for a, r in zip(losses, g["losses"]):
        _close(a, r)


def test_vqa_base_shape_matches_reference(golden_dir):
    """BASELINE.json configs[3] shape (133 text tokens, 20 tags, 50 regions, 3129 answers, BCE) at the base model
    size, batch 4: oracle loss / logits against the unmodified reference's (oracle/make_golden_base.py vqa)."""
    g = _load(golden_dir, "vqa_base.pt")
    cfg = O.Cfg(num_labels=3129, loss_type="bce", qa_answer_size=3129)
    sd = O.random_state_dict(cfg, g["head"], seed=g["wseed"])
    assert abs(_sum(sd.values()) - g["wsum"]) < 1e-6 * g["wsum"], "weight generator drifted"
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        loss, logits = O.vqa_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"], g["labels"],
                                     b["input_ids_b"], b["token_type_ids_b"], b["attention_mask_b"], b["img_feats"],
                                     max_tag_length=Lt)
    _close(loss, g["loss"]); _close(logits, g["logits"])


def test_long_sequence_and_retrieval_base_shapes_match_reference(golden_dir):
    """configs[4] per-sequence shape (70 + 20 + 100 -> 170 joint tokens, batch 4) and configs[2] per-pair shape
    (55 caption tokens, 6 pairs) at the base model size: oracle against the unmodified reference
    (oracle/make_golden_base.py long / retrieval)."""
    torch.set_num_threads(os.cpu_count() or 1)
    take = lambda t, r: torch.gather(t, 1, r[:, :, None].expand(-1, -1, t.shape[2]))
    g = _load(golden_dir, "rep_long_base.pt")
    cfg = O.Cfg()
    sd = O.random_state_dict(cfg, "rep", seed=g["wseed"])
    assert abs(_sum(sd.values()) - g["wsum"]) < 1e-6 * g["wsum"]
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    with torch.no_grad():
        seq, pooled, (txt, vis) = O.rep_forward(sd, cfg, max_tag_length=Lt, **b)
    _close(pooled, g["pooled"])
    _close(take(seq, g["rows"]["seq"]), g["seq_rows"])
    _close(take(txt, g["rows"]["txt"]), g["txt_rows"])
    _close(take(vis, g["rows"]["vis"]), g["vis_rows"])
    g = _load(golden_dir, "retrieval_base.pt")
    sd = O.random_state_dict(cfg, "retrieval", seed=g["wseed"])
    assert abs(_sum(sd.values()) - g["wsum"]) < 1e-6 * g["wsum"]
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    with torch.no_grad():
        gt, gi = O.forward_single(sd, cfg, **b)
        fine = O.retrieval_fine_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                        max_tag_length=Lt, input_ids_b=b["input_ids_b"],
                                        token_type_ids_b=b["token_type_ids_b"], attention_mask_b=b["attention_mask_b"],
                                        img_feats=b["img_feats"])
    _close(gt, g["global_txt"]); _close(gi, g["global_img"]); _close(fine, g["fine_logits"])


ENC = ("input_ids_a", "token_type_ids_a", "attention_mask_a", "input_ids_b", "token_type_ids_b", "attention_mask_b",
       "img_feats")


def test_round2_cases_match_reference(golden_dir):
    """tests/golden/r2_tiny.pt (oracle/make_golden_r2.py): BiBertImgForMLM, the 'mlp' retrieval classifier, use_b=True
    and phrase_mod='hard' + qa_ans with ignored labels -- the oracle reproduces the REAL reference's outputs."""
    g = _load(golden_dir, "r2_tiny.pt")
    B, La, Lt, R = g["dims"]
    cfg = O.Cfg(**g["cfg"])
    c = g["mlm"]
    sd = O.random_state_dict(cfg, "mlm", seed=c["wseed"])
    assert abs(_sum(sd.values()) - c["wsum"]) < 1e-6 * c["wsum"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=c["bseed"], ragged=True)
    b["input_ids_a"][c["mask_positions"]] = 103
    with torch.no_grad():
        scores, rel = O.mlm_forward(sd, cfg, *[b[k] for k in ENC], max_tag_length=Lt)
    _close(scores, c["scores"]); _close(rel, c["rel"])

    c = g["retrieval_mlp"]
    sd = O.random_state_dict(cfg, "retrieval_mlp", seed=c["wseed"])
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=c["bseed"], ragged=True)
    with torch.no_grad():
        fine = O.retrieval_fine_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                        max_tag_length=Lt, **{k: b[k] for k in ENC[3:]})
        total, logits, _, _, labels = O.retrieval_train_forward(sd, cfg, *[b[k] for k in ENC], max_tag_length=Lt,
                                                                dice_index=c["dice"])
    _close(fine, c["fine_logits"]); _close(total, c["train_total"]); _close(logits, c["train_logits"])
    assert torch.equal(labels, c["train_labels"])

    c = g["use_b"]
    cfg3 = O.Cfg(**dict(g["cfg"], num_labels=3, loss_type="xe"))
    sd = {k: v.requires_grad_(True) for k, v in O.random_state_dict(cfg3, "cls_linear", seed=c["wseed"]).items()}
    b = O.synthetic_batch(cfg3, B, La, Lt, R, seed=c["bseed"], ragged=True)
    loss, logits = O.seqcls_forward(sd, cfg3, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"], c["labels"],
                                    b["input_ids_b"], b["token_type_ids_b"], b["attention_mask_b"], b["img_feats"],
                                    max_tag_length=Lt, use_b=True)
    loss.backward()
    _close(loss.detach(), c["loss"]); _close(logits.detach(), c["logits"])
    for k, gr in c["grads"].items():
        _close(sd[k].grad, gr, 5e-5)

    c = g["pretrain_hard"]
    sd = {k: v.requires_grad_(True) for k, v in O.random_state_dict(cfg, "pretrain", seed=c["wseed"]).items()}
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=c["bseed"], ragged=True, with_labels=True)
    losses = O.pretrain_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                b["masked_lm_labels_a"], b["input_ids_b"], b["token_type_ids_b"], b["attention_mask_b"],
                                b["masked_lm_labels_b"], b["img_feats"], max_tag_length=Lt, img_index=b["img_index"],
                                phrase_index=b["phrase_index"], dice_index=b["dice_index"], rand_pos=b["rand_pos"],
                                rand_neg=b["rand_neg"], phrase_mod="hard", qa_ans=c["qa_ans"])
    assert len(losses) == 7
    losses[0].backward()
    for a, r in zip(losses, c["losses"]):
        _close(a.detach(), r)
    for k, gr in c["grads"].items():
        _close(sd[k].grad, gr, 5e-5)
    for k, n in c["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - n) <= 1e-4 * n + 1e-7, k


def test_bf16_storage_distance_to_fp32_reference(golden_dir, capsys):
    """What bf16 STORAGE alone costs against the fp32 reference -- no kernel involved: the oracle with its
    activations rounded to bf16 at the CUDA path's store points (O.bf16_stores()) against the reference goldens of
    the pre-training step.  This is the part of the CUDA-vs-fp32 distance that no kernel can remove (the fp32
    verification tier does); the GPU suite asserts the kernels against the bf16-store oracle at rtol 1e-2 and only
    bounds this distance."""
    g = _load(golden_dir, "pretrain_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True, with_labels=True)
    sd = {k: v.requires_grad_(True) for k, v in O.random_state_dict(cfg, "pretrain", seed=g["wseed"]).items()}
    with O.bf16_stores():
        losses = O.pretrain_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"], b["attention_mask_a"],
                                    b["masked_lm_labels_a"], b["input_ids_b"], b["token_type_ids_b"],
                                    b["attention_mask_b"], b["masked_lm_labels_b"], b["img_feats"], max_tag_length=Lt,
                                    img_index=b["img_index"], phrase_index=b["phrase_index"], dice_index=b["dice_index"],
                                    neg_img=b["neg_img"], rand_pos=b["rand_pos"], rand_neg=b["rand_neg"])
        losses[0].backward()
    loss_rel = max(abs(float(a.detach()) - float(r)) / abs(float(r)) for a, r in zip(losses, g["losses"]))
    rels = {k: float((sd[k].grad - gr).norm() / gr.norm()) for k, gr in g["grads"].items() if float(gr.norm()) > 1e-6}
    norm_rel = {k: abs(float(sd[k].grad.norm()) - n) / n for k, n in g["grad_norms"].items() if n > 1e-6}
    with capsys.disabled():
        print(f"\n[bf16 storage vs fp32 reference, tiny pre-training step] losses: max rel {loss_rel:.2e}; gradient "
              f"tensors: max relative L2 {max(rels.values()):.2e}; gradient norms: max rel {max(norm_rel.values()):.2e}")
    assert loss_rel < 1e-2                    # losses survive bf16 storage at rtol 1e-2 ...
    assert 3e-3 < max(rels.values()) < 6e-2   # ... individual gradient tensors do not (1e-2 is NOT reachable in bf16)
