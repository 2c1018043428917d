"""GPU: BASELINE.json's FULL-size configurations (C2 pre-training batch 256, C4 VQA batch 512 x 183
tokens, C5 batch 1024 x 170 tokens) checked through size-independent properties, chained to the
oracle: the CPU oracle is only affordable on a handful of samples, so

    oracle(sub-batch)  ~=  cuda(sub-batch)  ==  cuda(full batch)[sub-batch]      (bit exact: a row of
    a GEMM / LayerNorm / attention head never depends on which other rows share the launch)

plus exact scaling of the backward pass (loss x 2 -> every gradient x 2), the masking property of
SURVEY 8(c) and sim_mat == product of the forward_single embeddings.
"""
import os

import pytest
import torch

from oracle import mvptr_oracle as O
import mvptr_parity_utils as P

pytestmark = pytest.mark.gpu

ENC = ("input_ids_a", "token_type_ids_a", "attention_mask_a", "input_ids_b", "token_type_ids_b", "attention_mask_b",
       "img_feats")


def _sub(batch, n):
    return {k: v[:n].contiguous() for k, v in batch.items() if k in ENC}


def test_c5_long_sequence_batch_1024_rows_match_small_batch_and_oracle():
    """config 5: 100 regions + 70 text+phrase tokens, batch 1024, all-token outputs (BiImageBertRep)."""
    cfg = O.Cfg()
    sd = O.random_state_dict(cfg, "rep", seed=0)
    B, La, Lt, R, n = 1024, 70, 20, 100, 2
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=5, ragged=True)
    enc = {k: b[k] for k in ENC}
    model = P.build("BiImageBertRep", cfg, sd)
    with torch.no_grad():
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **P.to_cuda(enc))
        seq_s, pooled_s, (txt_s, vis_s) = model(max_tag_length=Lt, **P.to_cuda(_sub(b, n)))
    assert seq.shape == (B, La + R, cfg.hidden_size) and txt.shape == (B, La, cfg.hidden_size)
    assert vis.shape == (B, Lt + R, cfg.hidden_size) and pooled.shape == (B, cfg.hidden_size)
    assert torch.isfinite(seq.float()).all()
    jm = torch.cat([b["attention_mask_a"], b["attention_mask_b"][:, Lt:]], 1)[:n].bool().cuda()
    assert torch.equal(pooled[:n], pooled_s)
    assert torch.equal(seq[:n][jm], seq_s[jm])
    assert torch.equal(txt[:n][b["attention_mask_a"][:n].bool().cuda()], txt_s[b["attention_mask_a"][:n].bool().cuda()])
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():
        o_seq, o_pooled, (o_txt, o_vis) = O.rep_forward(sd, cfg, max_tag_length=Lt, **_sub(b, n))
    P.valid_rows_close(seq_s, o_seq, jm.cpu(), 2e-2, 4e-2, "seq")
    P.valid_rows_close(vis_s, o_vis, b["attention_mask_b"][:n], 2e-2, 3e-2, "vis")
    P.close(pooled_s, o_pooled, 2e-2, 3e-2, "pooled")


def test_c4_vqa_batch_512_logits_match_small_batch_and_oracle_and_backward_scales_exactly():
    """config 4: VQA-shaped fine-tune step, max_seq 128(+5) + 50 regions, 3129-way head, batch 512."""
    cfg = O.Cfg(loss_type="bce", num_labels=3129)
    sd = O.random_state_dict(cfg, "vqa", seed=0)
    B, La, Lt, R, n = 512, 133, 20, 50, 2
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=4, ragged=True)
    g = torch.Generator().manual_seed(44)
    labels = torch.zeros(B, cfg.qa_answer_size)
    pick = torch.randint(0, cfg.qa_answer_size, (B, 10), generator=g)
    labels.scatter_(1, pick, torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (B, 10), generator=g)])
    model = P.build("BiImageBertForVQA", cfg, sd, train=True)  # dropout 0: train mode only enables backward
    cb = P.to_cuda({k: b[k] for k in ENC})
    loss, logits = model(labels=labels.cuda(), max_tag_length=Lt, **cb)[:2]
    model.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    assert logits.shape == (B, cfg.qa_answer_size) and torch.isfinite(loss)
    arena = model.runtime().arena
    g1 = arena.grad.clone()
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
    # loss x 2 -> every gradient x 2 (power-of-two scaling is exact in bf16 and fp32; the only slack is
    # the order of the fp32 atomic / split-K accumulations)
    loss2 = model(labels=labels.cuda(), max_tag_length=Lt, **cb)[0] * 2.0
    model.zero_grad()
    loss2.backward()
    torch.cuda.synchronize()
    rel = float((arena.grad - 2.0 * g1).norm() / (2.0 * g1).norm())
    # (the answer-decoder dgrad is a split-K fp32 reduce-add followed by ONE cast to bf16: an unordered sum that
    # lands on the other side of a bf16 rounding boundary moves that element by one bf16 ulp -> ~5e-5 overall;
    # a real scaling bug would show up at >= 1e-2)
    assert rel < 2e-4, f"backward is not linear in the incoming gradient: rel {rel:.2e}"
    # rows of the full batch == the same rows run alone == the oracle
    # (same mode as the big step: with activations saved for backward, LayerNorm normalises the bf16
    # rounding of its input -- the value backward will see -- so no_grad outputs differ by an ulp)
    small = model(max_tag_length=Lt, **P.to_cuda(_sub(b, n)))[0].detach()
    assert torch.equal(logits[:n].detach(), small)
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():
        o_loss, o_logits = O.vqa_forward(sd, cfg, b["input_ids_a"][:n], b["token_type_ids_a"][:n], b["attention_mask_a"][:n],
                                         labels[:n], b["input_ids_b"][:n], b["token_type_ids_b"][:n],
                                         b["attention_mask_b"][:n], b["img_feats"][:n], max_tag_length=Lt)
    P.close(small, o_logits, 2e-2, 3e-2, "vqa logits (rows of the batch-512 step)")


def test_c2_pretrain_batch_256_properties():
    """config 2 at full size: sim_mat == product of the forward_single embeddings, masked positions do
    not leak, integer outputs (ITM labels / hard-negative indexes) are valid, six finite losses."""
    cfg = O.Cfg()
    sd = O.random_state_dict(cfg, "pretrain", seed=0)
    B, La, Lt, R = 256, 40, 20, 50
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=2, ragged=True, with_labels=True)
    model = P.build("BiBertImgForPreTraining", cfg, sd, train=True, max_text_seq_length=La)
    cb = P.to_cuda(b)
    enc = {k: cb[k] for k in ENC}
    with torch.no_grad():
        outs, (txt, vis, sim), (hard_txt, hard_img) = model.bert(max_tag_length=Lt, encode_hn=True, **enc)
        gt, gi = model.bert.forward_single(max_tag_length=Lt, **enc)
    assert sim.shape == (B, B)
    P.close(sim, (gt.float() @ gi.float().t()).cpu(), 1e-3, 1e-3, "sim_mat vs forward_single product")
    # hard negatives are integer work: valid indexes, never the positive itself on the replaced side
    assert hard_txt.dtype == torch.int64 and int(hard_txt.min()) >= 0 and int(hard_txt.max()) < B
    assert int(hard_img.min()) >= 0 and int(hard_img.max()) < B
    assert bool((hard_txt != hard_img).all())  # a hard-negative pair never is the matched pair
    # masking property at full size
    b2 = dict(enc)
    ids = enc["input_ids_a"].clone()
    ids[enc["attention_mask_a"] == 0] = 1234
    b2["input_ids_a"] = ids
    with torch.no_grad():
        gt2, gi2 = model.bert.forward_single(max_tag_length=Lt, **b2)
    assert torch.equal(gt2, gt) and torch.equal(gi2, gi)
    # the full step: six finite losses, gradients reach the tied word embeddings
    losses = model(max_tag_length=Lt, **{k: cb[k] for k in ENC + ("masked_lm_labels_a", "masked_lm_labels_b",
                                                                  "phrase_index")}, img_index=cb["img_index"])
    assert len(losses) == 6 and all(torch.isfinite(x) for x in losses)
    P.close(losses[0].detach(), sum(x.detach().float().cpu() for x in losses[1:]), 1e-4, 1e-4, "total = sum of parts")
    model.zero_grad()
    losses[0].backward()
    gw = model.bert.embeddings.word_embeddings.weight.grad
    assert torch.isfinite(gw).all() and float(gw[: cfg.only_word_size].abs().sum()) > 0
