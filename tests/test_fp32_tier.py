"""GPU: the fp32 VERIFICATION tier (model.set_precision("fp32"); csrc/fp32_tier.cu + engine_fp32.py).

BASELINE.json north_star: "bit-exact for token/region indexing, masking and top-k ranking order under fp32, and
logits/losses/gradients within ... 1e-4 in fp32".  Here the CUDA path is compared with the fp32 goldens of the REAL
reference (tests/golden) and with the fp32 oracle at rtol 1e-4, and -- unlike the bf16 suite -- NOTHING is injected
except the reference's own RNG draws: the hard-negative arg-max picks and the coarse top-k candidate lists are the
CUDA path's own and must equal the reference's.  Metric: mvptr_parity_utils.rel_err (|d| / max(|ref|, rms(ref))).
"""
import os

import pytest
import torch

from oracle import mvptr_oracle as O
import mvptr_parity_utils as P
from test_model_parity import ENC, LOSS_NAMES, _golden, _pad_choices, _pre_kw, _threads, oracle_pretrain

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _lines():
    import conftest
    return conftest.PARITY_LINES


def check(name, got, ref, tol=TOL):
    e = P.rel_err(got, ref)
    msg = f"[fp32 tier] {name}: vs fp32 reference {e:.2e} (tol {tol:.0e})"
    _lines().append(msg)
    print(msg)
    assert e <= tol, msg
    return e


def build32(cls_name, cfg, sd, train=False, **extra):
    model = P.build(cls_name, cfg, sd, dropout=0.0, train=train, **extra)
    return model.set_precision("fp32")


def test_split_gemm_is_fp32_accurate():
    """The contraction primitive: six bf16 products of 3-way splits against float64, on the K=2054 region
    projection shape (ragged M, N; K not a multiple of 8) and on MN-major operands (the wgrad layout)."""
    from mvp_pytorch_b200 import engine_fp32 as F
    g = torch.Generator(device="cuda").manual_seed(0)
    M, N, K = 130, 72, 2054
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) * 0.02
    bias = torch.randn(N, device="cuda", generator=g)
    D = torch.empty(M, N, device="cuda")
    F.gemm(F.split3(A, M, K, K), F.split3(B, N, K, K), D, M, N, K, ldd=N, bias=bias)
    ref = (A.double() @ B.double().t() + bias.double())
    scale = float((A.double().abs() @ B.double().abs().t()).max())
    err = float((D.double() - ref).abs().max()) / scale
    # wgrad layout: D2[N2, K2] = X^T Y with X [T, N2], Y [T, K2]
    T, N2, K2 = 300, 64, 136
    X = torch.randn(T, N2, device="cuda", generator=g)
    Y = torch.randn(T, K2, device="cuda", generator=g)
    D2 = torch.zeros(N2, K2, device="cuda")
    F.gemm(F.split3(X, T, N2, N2), F.split3(Y, T, K2, K2), D2, N2, K2, T, ldd=K2, a_mn=True, b_mn=True, accumulate=True)
    ref2 = X.double().t() @ Y.double()
    err2 = float((D2.double() - ref2).abs().max()) / float((X.double().abs().t() @ Y.double().abs()).max())
    _lines().append(f"[fp32 tier] split-3 tensor-core GEMM vs float64: {err:.2e} (K-major, K=2054), {err2:.2e} (MN-major) "
                    "of sum|a||b|")
    assert err < 5e-7 and err2 < 5e-7


def test_rep_tiny_fp32(golden_dir):
    g = _golden(golden_dir, "rep_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, "rep", seed=g["wseed"])
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    model = build32("BiImageBertRep", cfg, sd)
    with torch.no_grad():
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **P.to_cuda(b))
    assert seq.dtype == torch.float32
    jm = torch.cat([b["attention_mask_a"], b["attention_mask_b"][:, Lt:]], 1)
    for name, got, m in (("txt", txt, b["attention_mask_a"]), ("vis", vis, b["attention_mask_b"]), ("seq", seq, jm)):
        check(f"rep_tiny {name} (valid rows)", got.cpu()[m.bool()], g[name][m.bool()])
    check("rep_tiny pooled", pooled, g["pooled"])


def test_retrieval_tiny_fp32_with_its_own_hard_negatives(golden_dir):
    g = _golden(golden_dir, "retrieval_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, "retrieval", seed=g["wseed"])
    B, La, Lt, R = g["dims"]
    cpu_b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    b = P.to_cuda(cpu_b)
    model = build32("BiImageBertForRetrieval", cfg, sd)
    with torch.no_grad():
        model.forward_mod = "coarse"
        gt, gi = model(max_tag_length=Lt, **b)
        model.forward_mod = "fine"
        fine = model(max_tag_length=Lt, **b)
    check("retrieval_tiny global_txt", gt, g["global_txt"])
    check("retrieval_tiny global_img", gi, g["global_img"])
    check("retrieval_tiny fine ITM logits", fine, g["fine_logits"])
    model.forward_mod = "train"
    model.train()
    orig = torch.randperm
    try:  # only the reference's RNG draw is replayed; the arg-max picks are the CUDA path's own
        torch.randperm = lambda n, **kw: g["dice"].to(kw.get("device", "cpu"))
        total, logits, vsc, itm, labels = model(max_tag_length=Lt, **b)
    finally:
        torch.randperm = orig
    assert torch.equal(labels.cpu(), g["train_labels"])
    check("retrieval_tiny train vsc", vsc, g["train_vsc"])
    check("retrieval_tiny train ITM logits (own hard negatives)", logits, g["train_logits"])
    check("retrieval_tiny train total", total, g["train_total"])


def _run_pretrain32(cfg, sd, b, Lt, **fwd_kw):
    model = build32("BiBertImgForPreTraining", cfg, sd, train=True, max_text_seq_length=b["input_ids_a"].shape[1])
    cb = P.to_cuda(b)
    orig = torch.randperm
    try:
        torch.randperm = lambda n, **kw: b["dice_index"].to(kw.get("device", "cpu"))
        kw = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in _pre_kw(b, Lt).items()}
        losses = model(wra_choices=(cb["neg_img"], _pad_choices(cb["rand_pos"]), _pad_choices(cb["rand_neg"])), **kw, **fwd_kw)
    finally:
        torch.randperm = orig
    model.zero_grad()
    losses[0].backward()
    torch.cuda.synchronize()
    return model, losses


def _check_grads(name, params, ref, tol=TOL):
    top = max(float(v.float().norm()) for v in ref.values())
    worst, wk = 0.0, ""
    for k, r in ref.items():
        if float(r.float().norm()) <= 1e-5 * top:  # (numerically) zero-gradient tensors: key biases
            assert float(params[k].grad.float().norm()) <= 2e-5 * top, k
            continue
        e = P.rel_l2(params[k].grad, r)
        if e > worst:
            worst, wk = e, k
    msg = f"[fp32 tier] {name}: {len(ref)} gradient tensors, worst relative L2 vs fp32 reference {worst:.2e} ({wk}) (tol {tol:.0e})"
    _lines().append(msg)
    print(msg)
    assert worst <= tol, msg


def test_pretrain_tiny_fp32_losses_and_grads(golden_dir):
    g = _golden(golden_dir, "pretrain_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, "pretrain", seed=g["wseed"])
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True, with_labels=True)
    model, losses = _run_pretrain32(cfg, sd, b, Lt)
    for n, a, r in zip(LOSS_NAMES, losses, g["losses"]):
        check(f"pretrain_tiny loss {n}", a.detach(), r)
    _, g32 = oracle_pretrain(cfg, sd, b, Lt, bf16=False)  # == the reference's gradients (<= 1.5e-7), all tensors
    _check_grads("pretrain_tiny gradients", dict(model.named_parameters()), g32)
    for k, gr in g["grads"].items():
        assert P.rel_l2(dict(model.named_parameters())[k].grad, gr) < TOL, k


def test_pretrain_hard_phrase_mode_and_qa_fp32(golden_dir):
    g = _golden(golden_dir, "r2_tiny.pt")
    c = g["pretrain_hard"]
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, "pretrain", seed=c["wseed"])
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=c["bseed"], ragged=True, with_labels=True)
    model, losses = _run_pretrain32(cfg, sd, b, Lt, phrase_mod="hard", qa_ans=c["qa_ans"].cuda())
    for n, a, r in zip(["total", "vis_mlm", "vsc", "mlm", "itm", "qa", "wra"], losses, c["losses"]):
        check(f"pretrain(hard, qa) loss {n}", a.detach(), r)
    params = dict(model.named_parameters())
    for k, gr in c["grads"].items():
        assert P.rel_l2(params[k].grad, gr) < TOL, k


def test_vqa_and_mlm_tiny_fp32(golden_dir):
    g = _golden(golden_dir, "vqa_tiny.pt")
    cfg = O.Cfg(**g["cfg"])
    sd = O.random_state_dict(cfg, "vqa", seed=g["wseed"])
    B, La, Lt, R = g["dims"]
    cpu_b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    model = build32("BiImageBertForVQA", cfg, sd, train=True)
    out = model(labels=g["labels"].cuda(), max_tag_length=Lt, **P.to_cuda(cpu_b))
    model.zero_grad()
    out[0].backward()
    check("vqa_tiny loss", out[0].detach(), g["loss"])
    check("vqa_tiny logits", out[1].detach(), g["logits"])
    params = dict(model.named_parameters())
    for k, gr in g["grads"].items():
        assert P.rel_l2(params[k].grad, gr) < TOL, k
    g2 = _golden(golden_dir, "r2_tiny.pt")
    c = g2["mlm"]
    cfg = O.Cfg(**g2["cfg"])
    sd = O.random_state_dict(cfg, "mlm", seed=c["wseed"])
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=c["bseed"], ragged=True)
    b["input_ids_a"][c["mask_positions"]] = 103
    model = build32("BiBertImgForMLM", cfg, sd, max_text_seq_length=La)
    with torch.no_grad():
        scores, rel = model(max_tag_length=Lt, **P.to_cuda(b))
    check("mlm_tiny prediction scores", scores, c["scores"])
    check("mlm_tiny ITM logits", rel, c["rel"])
    assert torch.equal(scores.cpu().argmax(1), c["scores"].argmax(1))  # predicted tokens: bit-exact integer output


def test_base_shape_fp32_forward_and_pretrain_step(golden_dir):
    """Base model size: configs[0] forward against the reference's stored rows, and the configs[1]-shaped
    pre-training step (batch 6) -- six losses and all 314 gradient norms -- with the CUDA path's own hard negatives."""
    g = _golden(golden_dir, "rep_base.pt")
    cfg = O.Cfg()
    sd = O.random_state_dict(cfg, g["head"], seed=g["wseed"])
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True)
    model = build32("BiImageBertRep", cfg, sd)
    with torch.no_grad():
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **P.to_cuda(b))
    take = lambda t, r: torch.gather(t.float().cpu(), 1, r[:, :, None].expand(-1, -1, t.shape[2]))
    check("rep_base pooled", pooled, g["pooled"])
    for name, t in (("txt", txt), ("vis", vis), ("seq", seq)):
        check(f"rep_base {name} (reference's sampled valid rows)", take(t, g["rows"][name]), g[name + "_rows"])
    del model
    g = _golden(golden_dir, "pretrain_base.pt")
    sd = O.random_state_dict(cfg, "pretrain", seed=g["wseed"])
    B, La, Lt, R = g["dims"]
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True, with_labels=True)
    model, losses = _run_pretrain32(cfg, sd, b, Lt)
    for n, a, r in zip(LOSS_NAMES, losses, g["losses"]):
        check(f"pretrain_base loss {n}", a.detach(), r)
    params = dict(model.named_parameters())
    # every one of the 314 gradient tensors against the fp32 oracle (== the reference: its losses / gradient norms /
    # small tensors are the golden's, checked below)
    _, g32 = oracle_pretrain(cfg, sd, b, Lt, bf16=False)
    _check_grads("pretrain_base gradients", params, g32)
    top = max(g["grad_norms"].values())
    worst, wk = 0.0, ""
    for k, n in g["grad_norms"].items():
        if n <= 1e-5 * top or params[k].numel() > (1 << 22):
            continue  # (the stored fp32 norm of the 66 M-element embedding gradient carries ~1e-4 of summation noise itself)
        got = float(params[k].grad.double().norm())
        if abs(got - n) / n > worst:
            worst, wk = abs(got - n) / n, k
    _lines().append(f"[fp32 tier] pretrain_base: gradient norms vs the reference's stored norms, worst relative error {worst:.2e} ({wk})")
    assert worst <= TOL
    for k, gr in g["grads"].items():
        if float(gr.norm()) > 1e-5 * top:
            assert P.rel_l2(params[k].grad, gr) < TOL, k


def test_retrieval_subset_ranking_order_fp32():
    """SURVEY 8d C3 'ranking order checked in fp32 mode on a 200 x 1000 subset': 200 images x 1000 captions at the base
    model size through RetrievalScorer in fp32 -- the per-image top-128 caption lists and per-caption top-64 image
    lists (run_retrieval.py:481-522) and the in-batch hard negatives (modeling_vlbert.py:530-534) against the fp32
    oracle, no injection.  Two fp32 implementations with different summation order agree to ~1e-6, not to the bit,
    so 'equal' is evaluated at fp32 resolution: every position of every list must hold the reference's candidate or
    one whose reference score differs from it by <= 4e-6 (a tie at fp32 accuracy); the fraction of bit-identical
    positions is reported.  The ITM re-rank probabilities of a sample of candidate pairs close the chain."""
    from mvp_pytorch_b200.retrieval import RetrievalScorer
    _threads()
    cfg = O.Cfg()
    sd = O.random_state_dict(cfg, "retrieval", seed=5)
    n_img, caps_per_img, La, Lt, R = 200, 5, 55, 20, 50
    n_cap = n_img * caps_per_img
    cb = O.synthetic_batch(cfg, n_cap, La, Lt, R, seed=41, ragged=True)
    ib = O.synthetic_batch(cfg, n_img, La, Lt, R, seed=42, ragged=True)
    caps = {k: cb[k] for k in ENC[:3]}
    imgs = {k: ib[k] for k in ENC[3:]}
    model = build32("BiImageBertForRetrieval", cfg, sd)
    sc = RetrievalScorer(model, max_tag_length=Lt, stage1_batch=250, pair_batch=256)
    with torch.no_grad():
        gt, gi = sc.encode(P.to_cuda(caps), P.to_cuda(imgs))
        i2t, t2i = sc.coarse(128, 64)
    # oracle stage 1 (CPU fp32), in chunks
    with torch.no_grad():
        o_txt, o_gt, o_vis, o_gi = [], [], [], []
        nl, nh, eps = cfg.num_hidden_layers // 2, cfg.num_attention_heads, cfg.layer_norm_eps
        for s in range(0, n_cap, 250):
            ids, seg, m = (caps[k][s:s + 250] for k in ENC[:3])
            t, _ = O.encoder(sd, "bert.txt_encoder", O.embeddings(sd, "bert.embeddings", ids, seg, eps), O.ext_mask(m), nl, nh, eps)
            o_txt.append(t)
        o_txt = torch.cat(o_txt)
        eb = O.embeddings(sd, "bert.embeddings", imgs["input_ids_b"], imgs["token_type_ids_b"], eps)
        ie = O.layer_norm(O.linear(imgs["img_feats"], sd, "bert.img_embedding"), sd["bert.LayerNorm.weight"],
                          sd["bert.LayerNorm.bias"], cfg.img_layer_norm_eps)
        o_vis, _ = O.encoder(sd, "bert.vis_encoder", torch.cat([eb, ie], 1), O.ext_mask(imgs["attention_mask_b"]), nl, nh, eps)
        o_gt, o_gi = O.global_embeddings(sd, o_txt, o_vis)
        sims, o_i2t, o_t2i = O.coarse_candidates(o_gi, o_gt, 128, 64)
    check("retrieval subset: 1000 caption embeddings", gt, o_gt)
    check("retrieval subset: 200 image embeddings", gi, o_gi)

    def tie_aware(got, ref, scores, what):
        got = got.cpu()
        exact = float((got == ref).float().mean())
        s_got, s_ref = torch.gather(scores, 1, got), torch.gather(scores, 1, ref)
        worst = float((s_got - s_ref).abs().max())
        _lines().append(f"[fp32 tier] retrieval subset {what}: {exact:.4%} of positions bit-identical to the reference's "
                        f"list; worst reference-score gap at a differing position {worst:.2e} (fp32 tie <= 4e-6)")
        assert worst <= 4e-6, what
        assert exact > 0.98, what
        # same SET of candidates wherever the k-th / (k+1)-th scores are not tied
        return exact

    tie_aware(i2t, o_i2t, sims, "top-128 captions per image")
    tie_aware(t2i, o_t2i, sims.t().contiguous(), "top-64 images per caption")
    # in-batch hard negatives on the first 200 captions x 200 images (the arg-max of modeling_vlbert.py:530-534)
    import mvp_pytorch_b200.engine as E
    rt = model.runtime()
    block = E.sim_matrix(rt, gt[:n_img].contiguous(), gi)[:, :n_img].contiguous()
    h_img, h_txt = E.hard_negatives(rt, block)
    o_img, o_txt_idx = O.hard_negative_indexes(o_gt[:n_img] @ o_gi.t())
    ob = o_gt[:n_img] @ o_gi.t() - 2 * torch.eye(n_img)
    ar = torch.arange(n_img)
    gap = max(float((ob.max(1)[0] - ob[ar, h_img.cpu()]).max()), float((ob.max(0)[0] - ob[h_txt.cpu(), ar]).max()))
    same = float(((h_img.cpu() == o_img) & (h_txt.cpu() == o_txt_idx)).float().mean())
    _lines().append(f"[fp32 tier] retrieval subset hard negatives (200 x 200): {same:.2%} identical to the reference's "
                    f"arg-max; worst reference-score gap at a differing pick {gap:.2e}")
    assert gap <= 4e-6 and same > 0.97
    # fine stage on a sample of the candidate pairs: ITM match probability through the cached stage-1 tokens
    pick = torch.arange(0, n_img * 128, 401)[:64]
    cap_idx = i2t.reshape(-1)[pick.cuda()]
    img_idx = (pick // 128).cuda()
    with torch.no_grad():
        prob = sc.fine(cap_idx, img_idx).cpu()
        ci, ii = cap_idx.cpu(), img_idx.cpu()
        joint = torch.cat([o_txt[ci], o_vis[ii][:, Lt:]], 1)
        jm = torch.cat([O.ext_mask(caps["attention_mask_a"][ci]), O.ext_mask(imgs["attention_mask_b"][ii])[..., Lt:]], -1)
        s2, _ = O.encoder(sd, "bert.mul_encoder", joint, jm, nl, nh, eps)
        o_prob = O.itm_match_prob(O.linear(O.pooler(sd, "bert.pooler", s2), sd, "classifier"))
    check("retrieval subset: ITM match probability of 64 sampled candidate pairs", prob, o_prob)
