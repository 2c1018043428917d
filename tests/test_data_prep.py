"""SURVEY f-2: GPU-side feature decode and BERT / phrase masking against the CPU restatement of the reference's
dataset code (oracle/data_oracle.py: base64.b64decode + np.frombuffer, random_word, random_phrases) -- byte and
integer work, bit-exact."""
import base64

import numpy as np
import pytest
import torch

from oracle import data_oracle as D


def _samples(B, K, seed, max_boxes=60):
    g = np.random.RandomState(seed)
    texts, nbs = [], []
    for b in range(B):
        nb = int(g.randint(1, max_boxes + 1))
        feat = g.randn(nb, K).astype(np.float32)
        feat[g.rand(nb, K) < 0.01] = 0.0
        texts.append(base64.b64encode(feat.tobytes()))
        nbs.append(nb)
    return texts, nbs


def test_oracle_decode_is_the_reference_expression():
    texts, nbs = _samples(3, 2054, 0)
    out = D.decode_features(texts, nbs, 50, 2054)
    for b in range(3):
        ref = np.frombuffer(base64.b64decode(texts[b]), dtype=np.float32).reshape(nbs[b], 2054)
        n = min(nbs[b], 50)
        assert np.array_equal(out[b, :n].numpy(), ref[:n]) and float(out[b, n:].abs().sum()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("K,R,B", [(2054, 50, 9), (70, 9, 5), (7, 3, 4)])
def test_b64_decode_features_bit_exact(K, R, B):
    from mvp_pytorch_b200 import data
    texts, nbs = _samples(B, K, seed=K, max_boxes=R + 10)  # ragged: fewer AND more boxes than R (truncation)
    for dtype in (torch.float32, torch.bfloat16):
        got = data.decode_features(texts, nbs, R, K, "cuda", dtype)
        ref = D.decode_features(texts, nbs, R, K, dtype)
        assert torch.equal(got.cpu(), ref), (K, dtype)


@pytest.mark.gpu
def test_b64_decode_rejects_malformed_text():
    from mvp_pytorch_b200 import _lib, data
    texts, nbs = _samples(2, 70, 1)
    bad = [texts[0], texts[1][:-4]]                       # truncated
    with pytest.raises(_lib.MvptrError):
        data.decode_features(bad, nbs, 20, 70, "cuda")
    bad = [texts[0][:10] + b"*" + texts[0][11:], texts[1]]  # not a base64 character
    with pytest.raises(_lib.MvptrError):
        data.decode_features(bad, nbs, 20, 70, "cuda")


@pytest.mark.gpu
def test_mlm_mask_replays_the_reference_masking_bit_exactly():
    from mvp_pytorch_b200 import data
    g = torch.Generator().manual_seed(3)
    B, L, V, PV = 64, 40, 30522, 55529
    ids = torch.zeros(B, L, dtype=torch.int64)
    tok_first = torch.ones(B, dtype=torch.int32)
    tok_count = torch.randint(5, 30, (B,), generator=g, dtype=torch.int32)
    phr_count = torch.randint(0, 6, (B,), generator=g, dtype=torch.int32)
    phr_first = tok_first + tok_count
    u = torch.rand(B, L, generator=g)
    r = torch.randint(0, 1 << 40, (B, L), generator=g)
    max_links = 3
    links = torch.full((B, L, max_links), -1, dtype=torch.int32)
    maps = []
    for b in range(B):
        n, p = int(tok_count[b]), int(phr_count[b])
        ids[b, 0] = 101
        ids[b, 1:1 + n] = torch.randint(1000, V, (n,), generator=g)
        ids[b, 1 + n:1 + n + p] = torch.randint(V, V + PV, (p,), generator=g)
        ids[b, 1 + n + p] = 102
        m = {}
        for i in range(n):
            if p and float(torch.rand(1, generator=g)) < 0.3:
                k = int(torch.randint(1, max_links + 1, (1,), generator=g))
                m[i] = [int(x) for x in torch.randint(0, p, (k,), generator=g)]
                links[b, i, :k] = torch.tensor(m[i], dtype=torch.int32)
        maps.append(m)
    d_ids = ids.clone().cuda()
    labels = data.mask_tokens(d_ids, tok_first, tok_count, 103, V, phr_first, phr_count, links.cuda(), PV, V,
                              uniforms=u.cuda(), draws=r.cuda())
    for b in range(B):
        n, p = int(tok_count[b]), int(phr_count[b])
        toks = ids[b, 1:1 + n].tolist()
        toks, t1 = D.random_word_ids(toks, u[b, 1:1 + n].numpy(), r[b, 1:1 + n].numpy(), 103, V)
        phr = ids[b, 1 + n:1 + n + p].tolist()
        phr = D.random_phrases_ids(phr, t1, maps[b], u[b, 1 + n:1 + n + p].numpy(), r[b, 1 + n:1 + n + p].numpy(), 103, PV, V)
        exp_ids = ids[b].clone()
        exp_ids[1:1 + n] = torch.tensor(toks, dtype=torch.int64)
        exp_ids[1 + n:1 + n + p] = torch.tensor(phr, dtype=torch.int64)
        exp_lab = torch.full((L,), -1, dtype=torch.int64)
        exp_lab[1:1 + n] = torch.tensor(t1, dtype=torch.int64)
        assert torch.equal(d_ids[b].cpu(), exp_ids), b
        assert torch.equal(labels[b].cpu(), exp_lab), b
    # hashed draws: the masking rate and the 80 / 10 / 10 split of BERT
    big = torch.randint(1000, V, (4096, 40), dtype=torch.int64).cuda()
    orig = big.clone()
    lab = data.mask_tokens(big, torch.zeros(4096, dtype=torch.int32), torch.full((4096,), 40, dtype=torch.int32), 103, V, seed=11)
    sel = lab >= 0
    rate = float(sel.float().mean())
    masked = float((big[sel] == 103).float().mean())
    kept = float((big[sel] == orig[sel]).float().mean())
    assert abs(rate - 0.15) < 0.01 and abs(masked - 0.8) < 0.02 and abs(kept - 0.1) < 0.02
    assert torch.equal(lab[sel], orig[sel]) and torch.equal(big[~sel], orig[~sel])
