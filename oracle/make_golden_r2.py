"""Golden vectors for the round-2 parity cases, produced by the UNMODIFIED reference classes.
TEST INFRASTRUCTURE ONLY (runs in the authoring container, where /root/reference exists).

    python -B -m oracle.make_golden_r2      ->  tests/golden/r2_tiny.pt

Cases (each first asserts that the oracle restatement reproduces the reference):
* ``BiBertImgForMLM`` (modeling_vlbert.py:2559-2645): prediction scores at the [MASK] (id 103) positions + ITM logits;
* ``BiImageBertForRetrieval`` with ``config.classifier = 'mlp'`` (:1616-1629): train (recorded randperm) and fine;
* ``BiImageBertForSequenceClassification`` with ``use_b=True`` (:514-519, :1762-1798), forward + backward;
* ``BiBertImgForPreTraining`` with ``phrase_mod='hard'`` (:1270-1283) and ``qa_ans`` carrying an ignored (-1) label
  (:1260-1264): the 7 losses and gradients.
"""
import os

import torch

from oracle import mvptr_oracle as O
from oracle import ref_shim
from oracle.make_golden import TINY, OUT, Inject, close, checksum


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    mv = ref_shim.load()
    B, La, Lt, R = 6, 12, 5, 9
    out = dict(cfg=TINY, dims=(B, La, Lt, R))

    # ---- BiBertImgForMLM ----------------------------------------------------------------------------
    cfg = O.Cfg(**TINY)
    sd = O.random_state_dict(cfg, "mlm", seed=21)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=31, ragged=True)
    g = torch.Generator().manual_seed(41)
    pick = (torch.rand(B, La, generator=g) < 0.25) & (batch["attention_mask_a"] > 0) & (batch["input_ids_a"] < cfg.only_word_size)
    pick[:, 2] = True      # at least one [MASK] per caption (position 2 is valid for every row: n_txt >= 6)
    pick[3] = False        # ... except one caption without any
    batch["input_ids_a"][pick] = 103
    model = mv.BiBertImgForMLM(ref_shim.make_config(mv, cfg, max_text_seq_length=La)).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        scores, rel = model(max_tag_length=Lt, **batch)
        o_scores, o_rel = O.mlm_forward(sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"], batch["attention_mask_a"],
                                        batch["input_ids_b"], batch["token_type_ids_b"], batch["attention_mask_b"],
                                        batch["img_feats"], max_tag_length=Lt)
    print("mlm scores", tuple(scores.shape), close(o_scores, scores), "rel", close(o_rel, rel))
    out["mlm"] = dict(wseed=21, bseed=31, mask_positions=pick, scores=scores, rel=rel, wsum=checksum(sd.values()))

    # ---- retrieval, classifier='mlp' ------------------------------------------------------------------
    sd = O.random_state_dict(cfg, "retrieval_mlp", seed=22)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=32, ragged=True)
    model = mv.BiImageBertForRetrieval(ref_shim.make_config(mv, cfg, classifier="mlp", cls_hidden_scale=2)).eval()
    model.load_state_dict(sd, strict=True)
    dice = torch.randperm(B, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        model.forward_mod = "fine"
        fine = model(max_tag_length=Lt, **batch)
        model.forward_mod = "train"
        with Inject(mv, dice=dice):
            total, logits, vsc, itm, labels = model(max_tag_length=Lt, **batch)
        o_fine = O.retrieval_fine_forward(sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"],
                                          batch["attention_mask_a"], max_tag_length=Lt, input_ids_b=batch["input_ids_b"],
                                          token_type_ids_b=batch["token_type_ids_b"],
                                          attention_mask_b=batch["attention_mask_b"], img_feats=batch["img_feats"])
        o_total, o_logits, _, _, o_labels = O.retrieval_train_forward(
            sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"], batch["attention_mask_a"], batch["input_ids_b"],
            batch["token_type_ids_b"], batch["attention_mask_b"], batch["img_feats"], max_tag_length=Lt, dice_index=dice)
    print("retrieval mlp fine", close(o_fine, fine), "total", close(o_total, total), "logits", close(o_logits, logits))
    assert torch.equal(o_labels, labels)
    out["retrieval_mlp"] = dict(wseed=22, bseed=32, dice=dice, fine_logits=fine, train_total=total, train_logits=logits,
                                train_vsc=vsc, train_itm=itm, train_labels=labels)

    # ---- sequence classification, use_b=True ----------------------------------------------------------
    cfg3 = O.Cfg(**dict(TINY, num_labels=3, loss_type="xe"))
    sd = O.random_state_dict(cfg3, "cls_linear", seed=23)
    batch = O.synthetic_batch(cfg3, B, La, Lt, R, seed=33, ragged=True)
    labels3 = torch.randint(0, 3, (B,), generator=torch.Generator().manual_seed(3))
    model = mv.BiImageBertForSequenceClassification(ref_shim.make_config(mv, cfg3)).train()
    model.load_state_dict(sd, strict=True)
    loss, logits = model(labels=labels3, max_tag_length=Lt, use_b=True, **batch)[:2]
    model.zero_grad()
    loss.backward()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_loss, o_logits = O.seqcls_forward(sdg, cfg3, batch["input_ids_a"], batch["token_type_ids_a"], batch["attention_mask_a"],
                                        labels3, batch["input_ids_b"], batch["token_type_ids_b"], batch["attention_mask_b"],
                                        batch["img_feats"], max_tag_length=Lt, use_b=True)
    o_loss.backward()
    print("use_b loss", close(o_loss.detach(), loss.detach()), "logits", close(o_logits.detach(), logits.detach()))
    params = dict(model.named_parameters())
    keep = ["classifier.weight", "bert.pooler.dense.weight", "bert.mul_encoder.layer.0.attention.self.value.weight",
            "bert.vis_encoder.layer.1.output.dense.weight", "bert.embeddings.token_type_embeddings.weight"]
    for k in keep:
        close(sdg[k].grad, params[k].grad, tol=5e-5, what="use_b grad " + k)
    out["use_b"] = dict(wseed=23, bseed=33, labels=labels3, loss=loss.detach(), logits=logits.detach(),
                        grads={k: params[k].grad.clone() for k in keep})

    # ---- pre-training, phrase_mod='hard' + qa_ans with an ignored label ---------------------------------
    sd = O.random_state_dict(cfg, "pretrain", seed=24)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=34, ragged=True, with_labels=True)
    qa_ans = torch.tensor([3, -1, 0, 36, -1, 7])
    model = mv.BiBertImgForPreTraining(ref_shim.make_config(mv, cfg, max_text_seq_length=La)).train()
    model.load_state_dict(sd, strict=True)
    kw = dict(input_ids_a=batch["input_ids_a"], token_type_ids_a=batch["token_type_ids_a"],
              attention_mask_a=batch["attention_mask_a"], masked_lm_labels_a=batch["masked_lm_labels_a"],
              input_ids_b=batch["input_ids_b"], token_type_ids_b=batch["token_type_ids_b"],
              attention_mask_b=batch["attention_mask_b"], masked_lm_labels_b=batch["masked_lm_labels_b"],
              img_feats=batch["img_feats"], max_tag_length=Lt, img_index=batch["img_index"],
              phrase_index=batch["phrase_index"], qa_ans=qa_ans)
    # hard-negative rows first (they fix which phrases the negative pass reads), from the oracle (== reference, asserted below)
    with torch.no_grad():
        o = O.bibert_forward(sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"], batch["attention_mask_a"],
                             max_tag_length=Lt, input_ids_b=batch["input_ids_b"], token_type_ids_b=batch["token_type_ids_b"],
                             attention_mask_b=batch["attention_mask_b"], img_feats=batch["img_feats"], encode_hn=True,
                             dice_index=batch["dice_index"])
    hti = o[2][0]
    n_ph = (batch["phrase_index"][:, 1] - batch["phrase_index"][:, 0]).tolist()
    n_ph_hard = [n_ph[int(j)] for j in hti]
    # call order inside phrase_mod='hard' (:1274-1275): get_pos_sims over the batch, then over the hard batch
    randint_seq = [batch["rand_pos"][b, : n_ph[b]] for b in range(B) if n_ph[b] > 0] + \
                  [batch["rand_neg"][b, : n_ph_hard[b]] for b in range(B) if n_ph_hard[b] > 0]
    with Inject(mv, dice=batch["dice_index"], randint_seq=randint_seq):
        losses = model(phrase_mod="hard", **kw)
    assert len(losses) == 7
    model.zero_grad()
    losses[0].backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_losses = O.pretrain_forward(sdg, cfg, dice_index=batch["dice_index"], rand_pos=batch["rand_pos"],
                                  rand_neg=batch["rand_neg"], phrase_mod="hard", **kw)
    o_losses[0].backward()
    for i, (a, b) in enumerate(zip(o_losses, losses)):
        print("pretrain(hard, qa) loss", i, float(b), close(a.detach(), b.detach(), what=f"loss{i}"))
    worst = 0.0
    for k, gr in grads.items():
        worst = max(worst, close(sdg[k].grad, gr, tol=5e-5, what="grad " + k))
    print("pretrain(hard, qa) grads ok, worst", worst)
    keep = ["qa_head.weight", "qa_head.bias", "bert.mul_encoder.layer.1.output.dense.weight", "bert.txt_proj",
            "bert.pooler.dense.weight", "bert.embeddings.word_embeddings.weight"]
    out["pretrain_hard"] = dict(wseed=24, bseed=34, qa_ans=qa_ans, losses=[l.detach() for l in losses],
                                grad_norms={k: float(v.norm()) for k, v in grads.items()},
                                grads={k: grads[k] for k in keep})
    path = os.path.join(OUT, "r2_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
