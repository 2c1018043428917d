"""Generate tests/golden/*.pt from the REAL reference.  TEST INFRASTRUCTURE ONLY.

Run in the authoring container (needs /root/reference):

    python -B oracle/make_golden.py

For every case it (1) instantiates the unmodified reference class from
``/root/reference/oscar/modeling/modeling_vlbert.py``, (2) loads the seeded
weights, (3) runs it on the seeded synthetic batch with the reference's RNG
draws (randperm / randint / random.choice) replaced by recorded values,
(4) asserts ``oracle/mvptr_oracle.py`` reproduces the reference to <=2e-5 and
(5) writes inputs-recipe + reference outputs to ``tests/golden/``.

Weights and inputs are NOT stored: they are regenerated from (cfg, seed) by
``mvptr_oracle.random_state_dict`` / ``synthetic_batch`` (torch CPU generator,
same torch build here and on the GPU box); a checksum of both is stored so a
drift in the generator is detected rather than silently re-pinned.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import mvptr_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

TINY = dict(vocab_size=1500, only_word_size=1000, hidden_size=128, num_hidden_layers=4,
            num_attention_heads=2, intermediate_size=256, max_position_embeddings=64,
            img_feature_dim=70, qa_answer_size=37, num_labels=2)


def checksum(tensors):
    acc = 0.0
    for t in tensors:
        acc += float(t.double().abs().sum())
    return acc


class Inject:
    """Replace the reference's RNG draws by recorded values, in call order."""

    def __init__(self, mv, dice=None, randint_seq=None, choice_seq=None):
        self.mv, self.dice = mv, dice
        self.randint_seq = list(randint_seq or [])
        self.choice_seq = list(choice_seq or [])

    def __enter__(self):
        self._randperm, self._randint = torch.randperm, torch.randint
        self._choice = self.mv.random.choice
        if self.dice is not None:
            torch.randperm = lambda n, **kw: self.dice.clone()
        if self.randint_seq:
            def fake_randint(lo, hi, shape, **kw):
                v = self.randint_seq.pop(0)
                assert v.shape == tuple(shape), (v.shape, shape)
                return v
            torch.randint = fake_randint
        if self.choice_seq:
            self.mv.random.choice = lambda seq: self.choice_seq.pop(0)
        return self

    def __exit__(self, *a):
        torch.randperm, torch.randint = self._randperm, self._randint
        self.mv.random.choice = self._choice


def close(a, b, tol=2e-5, what=""):
    err = (a.double() - b.double()).abs().max().item()
    ref = b.double().abs().max().item() + 1e-12
    assert err <= tol * max(1.0, ref), f"{what}: oracle vs reference max|d|={err:.3e} (ref max {ref:.3e})"
    return err


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    mv = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    B, La, Lt, R = 6, 12, 5, 9

    # ---------------- case 1: BiImageBertRep (config-1 shape family) -------------
    cfg = O.Cfg(**TINY)
    sd = O.random_state_dict(cfg, "rep", seed=1)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=11, ragged=True)
    model = mv.BiImageBertRep(ref_shim.make_config(mv, cfg)).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **batch)
        o_seq, o_pooled, (o_txt, o_vis) = O.rep_forward(sd, cfg, max_tag_length=Lt, **batch)
    for a, b, n in ((o_seq, seq, "seq"), (o_pooled, pooled, "pooled"), (o_txt, txt, "txt"), (o_vis, vis, "vis")):
        print("rep", n, close(a, b, what=n))
    torch.save(dict(cfg=TINY, head="rep", wseed=1, bseed=11, dims=(B, La, Lt, R),
                    wsum=checksum(sd.values()), bsum=checksum([batch["img_feats"], batch["input_ids_a"]]),
                    seq=seq, pooled=pooled, txt=txt, vis=vis), os.path.join(OUT, "rep_tiny.pt"))

    # ---------------- case 2: BiImageBertForRetrieval coarse / fine / train ------
    sd = O.random_state_dict(cfg, "retrieval", seed=2)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=12, ragged=True)
    rcfg = ref_shim.make_config(mv, cfg)
    model = mv.BiImageBertForRetrieval(rcfg).eval()
    model.load_state_dict(sd, strict=True)
    dice = torch.randperm(B, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        model.forward_mod = "coarse"
        gt, gi = model(max_tag_length=Lt, **batch)
        model.forward_mod = "fine"
        fine = model(max_tag_length=Lt, **batch)
        model.forward_mod = "train"
        with Inject(mv, dice=dice):
            total, logits, vsc, itm, labels = model(max_tag_length=Lt, **batch)
        o_gt, o_gi = O.forward_single(sd, cfg, **batch)
        o_fine = O.retrieval_fine_forward(sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"],
                                          batch["attention_mask_a"], max_tag_length=Lt,
                                          input_ids_b=batch["input_ids_b"], token_type_ids_b=batch["token_type_ids_b"],
                                          attention_mask_b=batch["attention_mask_b"], img_feats=batch["img_feats"])
        o_total, o_logits, o_vsc, o_itm, o_labels = O.retrieval_train_forward(
            sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"], batch["attention_mask_a"],
            batch["input_ids_b"], batch["token_type_ids_b"], batch["attention_mask_b"], batch["img_feats"],
            max_tag_length=Lt, dice_index=dice)
    print("retr gt", close(o_gt, gt), "gi", close(o_gi, gi), "fine", close(o_fine, fine),
          "total", close(o_total, total), "logits", close(o_logits, logits))
    assert torch.equal(o_labels, labels)
    # fact used by the scorer: caching stage 1 and re-pairing == forward_fine
    torch.save(dict(cfg=TINY, head="retrieval", wseed=2, bseed=12, dims=(B, La, Lt, R), dice=dice,
                    wsum=checksum(sd.values()), global_txt=gt, global_img=gi, fine_logits=fine,
                    train_total=total, train_logits=logits, train_vsc=vsc, train_itm=itm, train_labels=labels),
               os.path.join(OUT, "retrieval_tiny.pt"))

    # ---------------- case 3: BiBertImgForPreTraining fwd + bwd (config-2 family) -
    sd = O.random_state_dict(cfg, "pretrain", seed=3)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=13, ragged=True, with_labels=True)
    model = mv.BiBertImgForPreTraining(ref_shim.make_config(mv, cfg, max_text_seq_length=La)).train()
    missing = model.load_state_dict(sd, strict=True)
    n_ph = (batch["phrase_index"][:, 1] - batch["phrase_index"][:, 0]).tolist()
    randint_seq, choice_seq = [], []
    for b in range(B):  # call order inside get_pos_neg_sims: pos randint, random.choice, neg randint
        if n_ph[b] > 0:
            randint_seq.append(batch["rand_pos"][b, : n_ph[b]])
        choice_seq.append(int(batch["neg_img"][b]))
        if n_ph[b] > 0:
            randint_seq.append(batch["rand_neg"][b, : n_ph[b]])
    kw = dict(input_ids_a=batch["input_ids_a"], token_type_ids_a=batch["token_type_ids_a"],
              attention_mask_a=batch["attention_mask_a"], masked_lm_labels_a=batch["masked_lm_labels_a"],
              input_ids_b=batch["input_ids_b"], token_type_ids_b=batch["token_type_ids_b"],
              attention_mask_b=batch["attention_mask_b"], masked_lm_labels_b=batch["masked_lm_labels_b"],
              img_feats=batch["img_feats"], max_tag_length=Lt, img_index=batch["img_index"],
              phrase_index=batch["phrase_index"])
    with Inject(mv, dice=batch["dice_index"], randint_seq=randint_seq, choice_seq=choice_seq):
        losses = model(**kw)
    losses[0].backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_losses = O.pretrain_forward(sdg, cfg, dice_index=batch["dice_index"], neg_img=batch["neg_img"],
                                  rand_pos=batch["rand_pos"], rand_neg=batch["rand_neg"], **kw)
    o_losses[0].backward()
    for i, (a, b) in enumerate(zip(o_losses, losses)):
        print("pretrain loss", i, float(b), close(a.detach(), b.detach(), what=f"loss{i}"))
    worst = 0.0
    for k, g in grads.items():
        og = sdg[k].grad
        assert og is not None, k
        worst = max(worst, close(og, g, tol=5e-5, what="grad " + k))
    no_grad = sorted(k for k in sd if k not in grads)
    print("pretrain grads ok, worst", worst, "params without grad in reference:", no_grad)
    keep = ["bert.embeddings.word_embeddings.weight", "bert.embeddings.position_embeddings.weight",
            "bert.img_embedding.weight", "bert.txt_proj", "logit_scale",
            "bert.mul_encoder.layer.1.attention.self.query.weight", "bert.txt_encoder.layer.0.output.dense.weight",
            "bert.vis_encoder.layer.0.attention.output.LayerNorm.weight", "cls.predictions.transform.dense.weight",
            "half_mlm.bias", "cls.seq_relationship.weight", "bert.pooler.dense.bias"]
    torch.save(dict(cfg=TINY, head="pretrain", wseed=3, bseed=13, dims=(B, La, Lt, R),
                    wsum=checksum(sd.values()), losses=[l.detach() for l in losses],
                    grad_norms={k: float(g.norm()) for k, g in grads.items()},
                    grads={k: grads[k] for k in keep}, no_grad=no_grad),
               os.path.join(OUT, "pretrain_tiny.pt"))

    # ---------------- case 4: BiImageBertForVQA fwd + bwd (config-4 family) -------
    vcfg = O.Cfg(**dict(TINY, num_labels=37, loss_type="bce"))
    sd = O.random_state_dict(vcfg, "vqa", seed=4)
    batch = O.synthetic_batch(vcfg, B, La, Lt, R, seed=14, ragged=True)
    g = torch.Generator().manual_seed(44)
    labels = torch.zeros(B, 37)
    for b in range(B):
        idx = torch.randperm(37, generator=g)[:3]
        labels[b, idx] = torch.tensor([0.3, 0.6, 1.0])
    model = mv.BiImageBertForVQA(ref_shim.make_config(mv, vcfg)).train()
    model.load_state_dict(sd, strict=True)
    loss, logits = model(labels=labels, max_tag_length=Lt, **batch)[:2]
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_loss, o_logits = O.vqa_forward(sdg, vcfg, batch["input_ids_a"], batch["token_type_ids_a"],
                                     batch["attention_mask_a"], labels, batch["input_ids_b"],
                                     batch["token_type_ids_b"], batch["attention_mask_b"], batch["img_feats"],
                                     max_tag_length=Lt)
    o_loss.backward()
    print("vqa loss", float(loss), close(o_loss.detach(), loss.detach()), "logits", close(o_logits.detach(), logits.detach()))
    for k, gr in grads.items():
        close(sdg[k].grad, gr, tol=5e-5, what="vqa grad " + k)
    torch.save(dict(cfg=dict(TINY, num_labels=37, loss_type="bce"), head="vqa", wseed=4, bseed=14,
                    dims=(B, La, Lt, R), labels=labels, wsum=checksum(sd.values()), loss=loss.detach(),
                    logits=logits.detach(), grad_norms={k: float(v.norm()) for k, v in grads.items()},
                    grads={k: grads[k] for k in ("cls.predictions.decoder.weight", "bert.img_embedding.weight",
                                                  "bert.mul_encoder.layer.0.intermediate.dense.weight")}),
               os.path.join(OUT, "vqa_tiny.pt"))

    # ---------------- case 5: ranking order of run_retrieval.py --------------------
    g = torch.Generator().manual_seed(7)
    sims = torch.randn(7, 40, generator=g)
    np_order = np.stack([np.argsort(sims[i].numpy())[::-1] for i in range(7)])  # run_retrieval.py:487
    mine = O.topk_desc(sims, 40).numpy()
    assert (np_order == mine).all()
    torch.save(dict(sims=sims, order=torch.from_numpy(np_order.copy())), os.path.join(OUT, "rank_order.pt"))

    # ---------------- case 6: AdamW known answers (optimization.py:130-189) -------
    from transformers.pytorch_transformers.optimization import AdamW, WarmupLinearSchedule
    w = torch.nn.Parameter(torch.tensor([0.1, -0.2, -0.1, 0.7]))
    opt = AdamW([w], lr=0.02, weight_decay=0.01)
    sched = WarmupLinearSchedule(opt, warmup_steps=2, t_total=10)
    traj, lrs = [], []
    # the reference calls the deprecated add_(scalar, tensor) overloads; emulate them on modern torch
    for it in range(5):
        w.grad = (w.detach() - torch.tensor([0.4, 0.2, -0.5, 0.1])) * 2
        try:
            opt.step()
        except TypeError:
            st = opt.state[w]
            if len(st) == 0:
                st["step"], st["exp_avg"], st["exp_avg_sq"] = 0, torch.zeros_like(w.data), torch.zeros_like(w.data)
            st["step"] += 1
            O.adamw_step(w.data, w.grad, st["exp_avg"], st["exp_avg_sq"], st["step"],
                         opt.param_groups[0]["lr"], weight_decay=0.01)
        sched.step()
        traj.append(w.detach().clone())
        lrs.append(opt.param_groups[0]["lr"])
    torch.save(dict(traj=torch.stack(traj), lrs=lrs), os.path.join(OUT, "adamw_traj.pt"))
    print("adamw lrs", lrs)
    print("golden written to", OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})


if __name__ == "__main__":
    main()
