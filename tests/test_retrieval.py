"""GPU: the sharded two-stage retrieval scorer (config 3) against the oracle's restatement of
run_retrieval.py's scoring loop."""
import pytest
import torch

from oracle import mvptr_oracle as O
import mvptr_parity_utils as P

pytestmark = pytest.mark.gpu


def test_topk_rows_reference_order():
    from mvp_pytorch_b200.retrieval import topk_rows
    g = torch.Generator().manual_seed(0)
    for rows, n, k in ((7, 40, 40), (33, 5000, 64), (9, 25000, 128), (3, 100, 1)):
        s = torch.randn(rows, n, generator=g)
        s[:, n // 2] = s[:, 1]          # exact ties
        s[0, :10] = 0.25
        idx, val = topk_rows(s.cuda(), k)
        ref = O.topk_desc(s, k)
        assert torch.equal(idx.cpu(), ref), (rows, n, k)   # integer ranking: bit exact
        assert torch.equal(val.cpu(), torch.gather(s, 1, ref))


def test_scorer_matches_reference_scoring_loop():
    cfg = O.Cfg(vocab_size=1500, only_word_size=1000, hidden_size=128, num_hidden_layers=4, num_attention_heads=2,
                intermediate_size=256, max_position_embeddings=64, img_feature_dim=70)
    sd = O.random_state_dict(cfg, "retrieval", seed=2)
    n_img, per, La, Lt, R = 6, 2, 12, 5, 9
    n_cap = n_img * per
    cb = O.synthetic_batch(cfg, n_cap, La, Lt, R, seed=21, ragged=True)
    ib = O.synthetic_batch(cfg, n_img, La, Lt, R, seed=22, ragged=True)
    caps = {k: cb[k] for k in ("input_ids_a", "token_type_ids_a", "attention_mask_a")}
    imgs = {k: ib[k] for k in ("input_ids_b", "token_type_ids_b", "attention_mask_b", "img_feats")}
    # ---- oracle: the reference loop (every pair re-encoded through all three encoders)
    with torch.no_grad():
        o_gt, _ = O.forward_single(sd, cfg, caps["input_ids_a"], caps["token_type_ids_a"], caps["attention_mask_a"],
                                   ib["input_ids_b"][:1].expand(n_cap, -1), ib["token_type_ids_b"][:1].expand(n_cap, -1),
                                   ib["attention_mask_b"][:1].expand(n_cap, -1), ib["img_feats"][:1].expand(n_cap, -1, -1))
        _, o_gi = O.forward_single(sd, cfg, cb["input_ids_a"][:n_img], cb["token_type_ids_a"][:n_img],
                                   cb["attention_mask_a"][:n_img], imgs["input_ids_b"], imgs["token_type_ids_b"],
                                   imgs["attention_mask_b"], imgs["img_feats"])
        o_sims, o_i2t, o_t2i = O.coarse_candidates(o_gi, o_gt, 4, 3)
        ci = torch.arange(n_cap).repeat_interleave(n_img)
        ii = torch.arange(n_img).repeat(n_cap)
        logits = O.retrieval_fine_forward(sd, cfg, caps["input_ids_a"][ci], caps["token_type_ids_a"][ci],
                                          caps["attention_mask_a"][ci], max_tag_length=Lt,
                                          input_ids_b=imgs["input_ids_b"][ii], token_type_ids_b=imgs["token_type_ids_b"][ii],
                                          attention_mask_b=imgs["attention_mask_b"][ii], img_feats=imgs["img_feats"][ii])
        o_prob = O.itm_match_prob(logits).view(n_cap, n_img)
    # ---- CUDA scorer
    from mvp_pytorch_b200.retrieval import RetrievalScorer, topk_rows, rank_of_first_positive
    model = P.build("BiImageBertForRetrieval", cfg, sd)
    sc = RetrievalScorer(model, max_tag_length=Lt, stage1_batch=5, pair_batch=16)
    gt, gi = sc.encode(P.to_cuda(caps), P.to_cuda(imgs))
    P.close(gt, o_gt, 1e-2, 1e-2, "global_txt"); P.close(gi, o_gi, 1e-2, 1e-2, "global_img")
    i2t, t2i = sc.coarse(4, 3)
    sims = sc.sims_rows.cpu()
    P.close(sims, o_sims, 1e-2, 1e-2, "coarse sims")
    # ranking is integer work: given the CUDA similarities it must equal the reference order exactly
    assert torch.equal(i2t.cpu(), O.topk_desc(sims, 4))
    prob = sc.fine(ci.cuda(), ii.cuda()).view(n_cap, n_img)
    P.close(prob, o_prob, 1e-2, 1e-2, "fine match probability")
    pos = (torch.arange(n_cap)[:, None] // per) == torch.arange(n_img)[None]
    ranks = rank_of_first_positive(prob, pos.cuda()).cpu().tolist()
    assert ranks == O.rank_of_first_positive(prob.cpu(), pos)
    res = sc.evaluate(P.to_cuda(caps), P.to_cuda(imgs), per, k_i2t=4, k_t2i=3)
    assert set(res) >= {"i2t_R@1", "t2i_R@10"} and res["i2t_ranks"].shape[0] == n_img
