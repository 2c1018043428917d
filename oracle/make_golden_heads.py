"""Golden vectors for the SURVEY 8(f-3) heads, produced by the UNMODIFIED reference classes.
TEST INFRASTRUCTURE ONLY (runs in the authoring container, where /root/reference exists).

    python -m oracle.make_golden_heads      ->  tests/golden/heads_tiny.pt

Cases: BiImageBertForSequenceClassificationPlus (visual entailment head) forward + backward,
BiImageBertForSequenceClassification with classifier='mlp', BiBertImgModel.forward_joint (two images),
BiBertImgModel.forward with hn_mod='sample' (torch.multinomial / torch.randperm replaced by recorded
draws).  Each case first asserts that the oracle restatement reproduces the reference.
"""
import os

import torch

from oracle import mvptr_oracle as O
from oracle import ref_shim
from oracle.make_golden import TINY, OUT, close, checksum


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    mv = ref_shim.load()
    B, La, Lt, R = 6, 12, 5, 9
    out = dict(cfg=TINY, dims=(B, La, Lt, R))

    # ---- visual entailment head (classifier='linear', 3 classes, CE) ----
    cfg = O.Cfg(**dict(TINY, num_labels=3, loss_type="xe"))
    sd = O.random_state_dict(cfg, "ve", seed=6)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=16, ragged=True)
    labels = torch.randint(0, 3, (B,), generator=torch.Generator().manual_seed(3))
    model = mv.BiImageBertForSequenceClassificationPlus(ref_shim.make_config(mv, cfg, classifier="linear")).train()
    model.load_state_dict(sd, strict=True)
    loss, logits = model(labels=labels, max_tag_length=Lt, **batch)[:2]
    model.zero_grad()
    loss.backward()
    with torch.no_grad():
        o_loss, o_logits = O.ve_plus_forward(sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"], batch["attention_mask_a"],
                                             labels, batch["input_ids_b"], batch["token_type_ids_b"],
                                             batch["attention_mask_b"], batch["img_feats"], max_tag_length=Lt)
    print("ve loss", close(o_loss, loss), "logits", close(o_logits, logits))
    params = dict(model.named_parameters())
    keep = ["single_mapping.0.weight", "single_mapping.2.bias", "classifier.weight", "bert.txt_proj", "bert.vis_proj",
            "bert.mul_encoder.layer.1.output.dense.weight", "bert.txt_encoder.layer.0.attention.self.query.weight"]
    out["ve"] = dict(wseed=6, bseed=16, labels=labels, loss=loss.detach(), logits=logits.detach(),
                     wsum=checksum(sd.values()), grads={k: params[k].grad.clone() for k in keep})

    # ---- BiImageBertForSequenceClassification, classifier='mlp' ----
    cfg2 = O.Cfg(**dict(TINY, num_labels=3, loss_type="xe"))
    sd2 = O.random_state_dict(cfg2, "cls_mlp", seed=7)
    model = mv.BiImageBertForSequenceClassification(ref_shim.make_config(mv, cfg2, classifier="mlp", cls_hidden_scale=2)).eval()
    model.load_state_dict(sd2, strict=True)
    with torch.no_grad():
        loss2, logits2 = model(labels=labels, max_tag_length=Lt, **batch)[:2]
        o_loss2, o_logits2 = O.seqcls_mlp_forward(sd2, cfg2, batch["input_ids_a"], batch["token_type_ids_a"],
                                                  batch["attention_mask_a"], labels, batch["input_ids_b"],
                                                  batch["token_type_ids_b"], batch["attention_mask_b"], batch["img_feats"],
                                                  max_tag_length=Lt)
    print("cls_mlp loss", close(o_loss2, loss2), "logits", close(o_logits2, logits2))
    out["cls_mlp"] = dict(wseed=7, bseed=16, labels=labels, loss=loss2, logits=logits2)

    # ---- forward_joint (two images) and hn_mod='sample' on the backbone ----
    cfg3 = O.Cfg(**TINY)
    sd3 = O.random_state_dict(cfg3, "rep", seed=8)
    b1 = O.synthetic_batch(cfg3, B, La, Lt, R, seed=17, ragged=True)
    b2 = O.synthetic_batch(cfg3, B, La, Lt, R, seed=18, ragged=True)
    rep = mv.BiImageBertRep(ref_shim.make_config(mv, cfg3)).eval()
    rep.load_state_dict(sd3, strict=True)
    with torch.no_grad():
        seq, pooled = rep.bert.forward_joint(
            input_ids_a=b1["input_ids_a"], token_type_ids_a=b1["token_type_ids_a"], attention_mask_a=b1["attention_mask_a"],
            max_tag_length=Lt, input_ids_b=b1["input_ids_b"], token_type_ids_b=b1["token_type_ids_b"],
            attention_mask_b=b1["attention_mask_b"], img_feats=b1["img_feats"], input_ids_b2=b2["input_ids_b"],
            token_type_ids_b2=b2["token_type_ids_b"], attention_mask_b2=b2["attention_mask_b"], img_feats2=b2["img_feats"])
        o_seq, o_pooled = O.forward_joint(sd3, cfg3, b1["input_ids_a"], b1["token_type_ids_a"], b1["attention_mask_a"], Lt,
                                          b1["input_ids_b"], b1["token_type_ids_b"], b1["attention_mask_b"], b1["img_feats"],
                                          b2["input_ids_b"], b2["token_type_ids_b"], b2["attention_mask_b"], b2["img_feats"])
    print("joint seq", close(o_seq, seq), "pooled", close(o_pooled, pooled))
    out["joint"] = dict(wseed=8, bseed1=17, bseed2=18, seq=seq, pooled=pooled)

    g = torch.Generator().manual_seed(9)
    dice = torch.randperm(B, generator=g)
    draw_img = torch.tensor([(i + 1 + int(torch.randint(0, B - 1, (1,), generator=g))) % B for i in range(B)])
    draw_txt = torch.tensor([(i + 1 + int(torch.randint(0, B - 1, (1,), generator=g))) % B for i in range(B)])
    draws = [draw_img.view(B, 1), draw_txt.view(B, 1)]
    seen = []
    orig_mn, orig_rp = torch.multinomial, torch.randperm
    try:
        torch.multinomial = lambda p, num_samples=1, **kw: (seen.append(p.clone()), draws[len(seen) - 1])[1]
        torch.randperm = lambda n, **kw: dice.clone()
        with torch.no_grad():
            outs, single, hard = rep.bert(encode_hn=True, hn_mod="sample", logit=torch.tensor(14.0), max_tag_length=Lt, **b1)
    finally:
        torch.multinomial, torch.randperm = orig_mn, orig_rp
    o_p1, o_p2 = O.negative_sampling_probs(single[2], torch.tensor(14.0))
    print("sample probs", close(o_p1, seen[0]), close(o_p2, seen[1]))
    out["sample"] = dict(logit=14.0, dice=dice, draw_img=draw_img, draw_txt=draw_txt, p_t2i=seen[0], p_i2t=seen[1],
                         hard_seq=outs[2], hard_pooled=outs[3], hard_txt_index=hard[0], hard_img_index=hard[1], sim=single[2])
    # ---- referring-expression head: mods 1-3, plus the mid-encoder (phrase_layer) variant, fwd + bwd ----
    cfg4 = O.Cfg(**dict(TINY, num_labels=1))
    sd4 = O.random_state_dict(cfg4, "re", seed=10)
    b4 = O.synthetic_batch(cfg4, B, La, Lt, R, seed=19, ragged=True)
    gl = torch.Generator().manual_seed(4)
    re_labels = torch.rand(B, R, generator=gl)
    re_labels[b4["attention_mask_b"][:, Lt:] == 0] = -1.0   # padded regions carry no label
    model = mv.BiImageBertForRE(ref_shim.make_config(mv, cfg4)).train()
    model.load_state_dict(sd4, strict=True)
    re_out = dict(wseed=10, bseed=19, labels=re_labels, cases={})
    keep = ["bert.mul_encoder.layer.0.output.dense.weight", "bert.txt_encoder.layer.1.attention.self.key.weight",
            "bert.img_embedding.weight"]
    for name, kw in (("mod1", dict(mod=1)), ("mod2", dict(mod=2)), ("mod3", dict(mod=3)),
                     ("mod1_mid", dict(mod=1, phrase_layer=0))):
        loss, logits = model(labels=re_labels, max_tag_length=Lt, **kw, **b4)
        model.zero_grad()
        loss.backward()
        with torch.no_grad():
            o_loss, o_logits = O.re_forward(sd4, cfg4, b4["input_ids_a"], b4["token_type_ids_a"], b4["attention_mask_a"],
                                            re_labels, b4["input_ids_b"], b4["token_type_ids_b"], b4["attention_mask_b"],
                                            b4["img_feats"], max_tag_length=Lt, **kw)
        print("re", name, "loss", close(o_loss, loss), "logits", close(o_logits, logits))
        params = dict(model.named_parameters())
        grads = {k: params[k].grad.clone() for k in keep}
        if name == "mod3":
            grads["classifier.weight"] = params["classifier.weight"].grad.clone()
        re_out["cases"][name] = dict(kw=kw, loss=loss.detach(), logits=logits.detach(), grads=grads)
    out["re"] = re_out
    torch.save(out, os.path.join(OUT, "heads_tiny.pt"))
    print("wrote", os.path.join(OUT, "heads_tiny.pt"), os.path.getsize(os.path.join(OUT, "heads_tiny.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
