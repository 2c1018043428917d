"""Golden vectors of BASELINE.json configs[0] from the REAL reference.  TEST INFRASTRUCTURE ONLY.

    python -B oracle/make_golden_base.py        (authoring container: needs /root/reference)

"MVP base cross-modal encoder forward, random-init, synthetic batch 8 x (35 text+phrase tokens, 50 regions x
2054-d) on CPU fp32": the unmodified reference ``BiImageBertRep`` (oscar/modeling/modeling_vlbert.py:2509-2557)
at the base shape (hidden 768, 12 heads, 6+6+6 layers, vocabulary 86 051) on the seeded synthetic batch with
ragged masks.  Asserts that ``oracle/mvptr_oracle.py`` reproduces it to <= 2e-5 and stores a compact sample of
the outputs (the pooled vectors in full, ROWS valid positions per sequence of the three token outputs, and
whole-tensor checksums) in ``tests/golden/rep_base.pt``.  Weights / inputs are regenerated from seeds.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import mvptr_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import checksum, close  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "rep_base.pt")
ROWS = 6
DIMS = (8, 35, 20, 50)  # B, La, Lt, R of configs[0]
WSEED, BSEED = 0, 1


def sample_rows(mask):
    """ROWS evenly spread VALID positions per sequence -> [B, ROWS] int64 (deterministic)."""
    out = []
    for m in mask:
        idx = torch.nonzero(m > 0).reshape(-1)
        pick = torch.linspace(0, idx.numel() - 1, ROWS).round().long()
        out.append(idx[pick])
    return torch.stack(out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    mv = ref_shim.load()
    cfg = O.Cfg()
    B, La, Lt, R = DIMS
    sd = O.random_state_dict(cfg, "rep", seed=WSEED)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=BSEED, ragged=True)
    model = mv.BiImageBertRep(ref_shim.make_config(mv, cfg)).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **batch)
        o_seq, o_pooled, (o_txt, o_vis) = O.rep_forward(sd, cfg, max_tag_length=Lt, **batch)
    jm = torch.cat([batch["attention_mask_a"], batch["attention_mask_b"][:, Lt:]], 1)
    for a, b, m, n in ((o_seq, seq, jm, "seq"), (o_txt, txt, batch["attention_mask_a"], "txt"),
                       (o_vis, vis, batch["attention_mask_b"], "vis")):
        print("base", n, close(a[m.bool()], b[m.bool()], what=n))
    print("base pooled", close(o_pooled, pooled, what="pooled"))
    rows = {"seq": sample_rows(jm), "txt": sample_rows(batch["attention_mask_a"]), "vis": sample_rows(batch["attention_mask_b"])}
    take = lambda t, r: torch.gather(t, 1, r[:, :, None].expand(-1, -1, t.shape[2])).clone()
    torch.save(dict(head="rep", wseed=WSEED, bseed=BSEED, dims=DIMS, wsum=checksum(sd.values()),
                    bsum=checksum([batch["img_feats"], batch["input_ids_a"]]), pooled=pooled.clone(),
                    rows=rows, seq_rows=take(seq, rows["seq"]), txt_rows=take(txt, rows["txt"]),
                    vis_rows=take(vis, rows["vis"]),
                    valid_abs_sum=dict(seq=float(seq[jm.bool()].double().abs().sum()),
                                       txt=float(txt[batch["attention_mask_a"].bool()].double().abs().sum()),
                                       vis=float(vis[batch["attention_mask_b"].bool()].double().abs().sum()))), OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def pretrain_case():
    """BASELINE.json configs[1] per-pair shape (40 text+phrase tokens, 20 tags, 50 regions) at the base model size,
    batch 6, forward + backward of the unmodified reference BiBertImgForPreTraining (modeling_vlbert.py:1133-1311)
    with its RNG draws replaced by recorded values -> tests/golden/pretrain_base.pt (six losses, every gradient
    norm, a few small gradient tensors in full)."""
    from oracle.make_golden import Inject
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    mv = ref_shim.load()
    cfg = O.Cfg()
    B, La, Lt, R = 6, 40, 20, 50
    sd = O.random_state_dict(cfg, "pretrain", seed=7)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=17, ragged=True, with_labels=True)
    model = mv.BiBertImgForPreTraining(ref_shim.make_config(mv, cfg, max_text_seq_length=La)).train()
    model.load_state_dict(sd, strict=True)
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
    with torch.no_grad():
        o_losses = O.pretrain_forward(sd, cfg, dice_index=batch["dice_index"], neg_img=batch["neg_img"],
                                      rand_pos=batch["rand_pos"], rand_neg=batch["rand_neg"], **kw)
    for i, (a, b) in enumerate(zip(o_losses, losses)):
        print("pretrain base loss", i, float(b), close(a.detach(), b.detach(), what=f"loss{i}"))
    keep = [k for k in grads if grads[k].numel() <= 3072 and ("layer.0." in k or "layer.5." in k or "layer" not in k)]
    out = os.path.join(os.path.dirname(OUT), "pretrain_base.pt")
    torch.save(dict(head="pretrain", wseed=7, bseed=17, dims=(B, La, Lt, R), wsum=checksum(sd.values()),
                    losses=[l.detach().clone() for l in losses],
                    grad_norms={k: float(g.norm()) for k, g in grads.items()},
                    grads={k: grads[k] for k in keep},
                    no_grad=sorted(k for k in sd if k not in grads)), out)
    print("wrote", out, os.path.getsize(out), "bytes;", len(keep), "gradient tensors in full,", len(grads), "norms")


def vqa_case():
    """BASELINE.json configs[3] shape (max_seq 128 + 5 phrase slots, 20 tags, 50 regions, 3129 answers, BCE loss) at
    the base model size, batch 4: forward + backward of the unmodified reference BiImageBertForVQA
    (modeling_vlbert.py:1801-1870) -> tests/golden/vqa_base.pt (loss, logits, all gradient norms)."""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    mv = ref_shim.load()
    cfg = O.Cfg(num_labels=3129, loss_type="bce", qa_answer_size=3129)
    B, La, Lt, R = 4, 133, 20, 50
    sd = O.random_state_dict(cfg, "vqa", seed=8)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=18, ragged=True)
    g = torch.Generator().manual_seed(48)
    labels = torch.zeros(B, 3129)
    for b in range(B):  # sparse soft scores as in VQA v2
        idx = torch.randperm(3129, generator=g)[:4]
        labels[b, idx] = torch.tensor([0.3, 0.6, 0.9, 1.0])
    model = mv.BiImageBertForVQA(ref_shim.make_config(mv, cfg)).train()
    model.load_state_dict(sd, strict=True)
    loss, logits = model(labels=labels, max_tag_length=Lt, **batch)[:2]
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        o_loss, o_logits = O.vqa_forward(sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"],
                                         batch["attention_mask_a"], labels, batch["input_ids_b"],
                                         batch["token_type_ids_b"], batch["attention_mask_b"], batch["img_feats"],
                                         max_tag_length=Lt)
    print("vqa base loss", float(loss), close(o_loss, loss.detach(), what="loss"), "logits",
          close(o_logits, logits.detach(), what="logits"))
    out = os.path.join(os.path.dirname(OUT), "vqa_base.pt")
    torch.save(dict(head="vqa", wseed=8, bseed=18, dims=(B, La, Lt, R), labels=labels, wsum=checksum(sd.values()),
                    loss=loss.detach().clone(), logits=logits.detach().clone(),
                    grad_norms={k: float(v.norm()) for k, v in grads.items()}), out)
    print("wrote", out, os.path.getsize(out), "bytes")


def rep_long_case():
    """BASELINE.json configs[4] per-sequence shape (70 text+phrase tokens, 20 tags, 100 regions -> 170 joint tokens) at the
    base model size, batch 4, inference: the unmodified reference BiImageBertRep -> tests/golden/rep_long_base.pt."""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    mv = ref_shim.load()
    cfg = O.Cfg()
    B, La, Lt, R = 4, 70, 20, 100
    sd = O.random_state_dict(cfg, "rep", seed=9)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=19, ragged=True)
    model = mv.BiImageBertRep(ref_shim.make_config(mv, cfg)).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **batch)
        o_seq, o_pooled, (o_txt, o_vis) = O.rep_forward(sd, cfg, max_tag_length=Lt, **batch)
    jm = torch.cat([batch["attention_mask_a"], batch["attention_mask_b"][:, Lt:]], 1)
    for a, b, m, n in ((o_seq, seq, jm, "seq"), (o_txt, txt, batch["attention_mask_a"], "txt"),
                       (o_vis, vis, batch["attention_mask_b"], "vis")):
        print("long", n, close(a[m.bool()], b[m.bool()], what=n))
    print("long pooled", close(o_pooled, pooled, what="pooled"))
    rows = {"seq": sample_rows(jm), "txt": sample_rows(batch["attention_mask_a"]), "vis": sample_rows(batch["attention_mask_b"])}
    take = lambda t, r: torch.gather(t, 1, r[:, :, None].expand(-1, -1, t.shape[2])).clone()
    out = os.path.join(os.path.dirname(OUT), "rep_long_base.pt")
    torch.save(dict(head="rep", wseed=9, bseed=19, dims=(B, La, Lt, R), wsum=checksum(sd.values()),
                    pooled=pooled.clone(), rows=rows, seq_rows=take(seq, rows["seq"]), txt_rows=take(txt, rows["txt"]),
                    vis_rows=take(vis, rows["vis"])), out)
    print("wrote", out, os.path.getsize(out), "bytes")


def retrieval_case():
    """BASELINE.json configs[2] per-pair shape (55 caption tokens, 20 tags, 50 regions) at the base model size, 6 pairs:
    the unmodified reference BiImageBertForRetrieval in its 'coarse' (uni-modal embeddings) and 'fine' (ITM logits)
    modes (modeling_vlbert.py:1643-1712) -> tests/golden/retrieval_base.pt."""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    mv = ref_shim.load()
    cfg = O.Cfg()
    B, La, Lt, R = 6, 55, 20, 50
    sd = O.random_state_dict(cfg, "retrieval", seed=10)
    batch = O.synthetic_batch(cfg, B, La, Lt, R, seed=20, ragged=True)
    model = mv.BiImageBertForRetrieval(ref_shim.make_config(mv, cfg)).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        model.forward_mod = "coarse"
        gt, gi = model(max_tag_length=Lt, **batch)
        model.forward_mod = "fine"
        fine = model(max_tag_length=Lt, **batch)
        o_gt, o_gi = O.forward_single(sd, cfg, **batch)
        o_fine = O.retrieval_fine_forward(sd, cfg, batch["input_ids_a"], batch["token_type_ids_a"],
                                          batch["attention_mask_a"], max_tag_length=Lt,
                                          input_ids_b=batch["input_ids_b"], token_type_ids_b=batch["token_type_ids_b"],
                                          attention_mask_b=batch["attention_mask_b"], img_feats=batch["img_feats"])
    print("retrieval base gt", close(o_gt, gt), "gi", close(o_gi, gi), "fine", close(o_fine, fine))
    out = os.path.join(os.path.dirname(OUT), "retrieval_base.pt")
    torch.save(dict(head="retrieval", wseed=10, bseed=20, dims=(B, La, Lt, R), wsum=checksum(sd.values()),
                    global_txt=gt.clone(), global_img=gi.clone(), fine_logits=fine.clone()), out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "rep":
        main()
    if len(sys.argv) < 2 or sys.argv[1] == "long":
        rep_long_case()
    if len(sys.argv) < 2 or sys.argv[1] == "retrieval":
        retrieval_case()
    if len(sys.argv) < 2 or sys.argv[1] == "vqa":
        vqa_case()
    if len(sys.argv) < 2 or sys.argv[1] == "pretrain":
        pretrain_case()
