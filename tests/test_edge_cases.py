"""GPU: edge cases of the hot path against the oracle / a torch restatement -- a single pair, the longest
supported sequence, fully masked key rows, images whose regions are all padding, the error conventions of the
reference (modeling_vlbert.py:435,451; modeling_bert.py:283-286) and loud failures for unsupported options."""
import pytest
import torch

from oracle import mvptr_oracle as O
import mvptr_parity_utils as P

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32
CFG = dict(vocab_size=1500, only_word_size=1000, hidden_size=128, num_hidden_layers=4, num_attention_heads=2,
           intermediate_size=256, max_position_embeddings=64, img_feature_dim=70)


@pytest.fixture(scope="module")
def lib():
    from mvp_pytorch_b200 import _lib
    _lib.lib()
    return _lib


def _attn_ref(qkv, maskadd, B, L, nh, H):
    q, k, v = qkv.float().view(B, L, 3, nh, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0 + maskadd[:, None, None, :]
    return (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, H)


@pytest.mark.parametrize("path", ["tc", "mma"])
@pytest.mark.parametrize("L", [1, 2, 17, 64, 65, 96, 97, 127, 128, 255, 256])
def test_attention_extreme_lengths_and_fully_masked_keys(lib, L, path):
    """L = 1 and the maximum L = 256; batch row 1 has EVERY key masked: the reference's additive -10000 then
    cancels in the softmax and the row attends uniformly-by-score to all keys (modeling_vlbert.py:430-460)."""
    lib.set_attention_path(path, path)  # tcgen05 kernels (L <= 128; tile-row classes 64 / 96 / 128) and mma.sync kernels
    B, nh = 2, 2
    H = nh * 64
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = (torch.randn(B * L, 3 * H, device="cuda", generator=g) * 0.5).to(BF16)
    mask = torch.ones(B, L, device="cuda")
    mask[1] = 0
    maskadd = ((1 - mask) * -10000.0).contiguous()
    ctx = torch.empty(B * L, H, device="cuda", dtype=BF16)
    lse = torch.empty(B, nh, L, device="cuda", dtype=F32)
    lib.call("mvptr_attn_fwd", qkv, 3 * H, maskadd, ctx, H, lse, B, L, nh, H, 0.0, 0)
    x = qkv.float().requires_grad_(True)
    ref = _attn_ref(x, maskadd, B, L, nh, H)
    assert torch.isfinite(ctx.float()).all()
    assert (ctx.float() - ref).abs().max() < 2e-2
    dctx = (torch.randn(B * L, H, device="cuda", generator=g) * 0.5).to(BF16)
    ref.backward(dctx.float())
    dqkv = torch.empty_like(qkv)
    lib.call("mvptr_attn_bwd", qkv, 3 * H, maskadd, ctx, dctx, H, lse, dqkv, None, B, L, nh, H, 0.0, 0)
    assert torch.isfinite(dqkv.float()).all()
    rel = ((dqkv.float() - x.grad).norm() / (x.grad.norm() + 1e-12)).item()
    assert rel < 3e-2, f"attn bwd rel l2 {rel} at L={L}"
    lib.set_attention_path("auto", "auto")


def test_attention_rejects_unsupported_shapes(lib):
    H = 128
    qkv = torch.zeros(257, 3 * H, device="cuda", dtype=BF16)
    ctx = torch.zeros(257, H, device="cuda", dtype=BF16)
    m = torch.zeros(1, 257, device="cuda")
    with pytest.raises(lib.MvptrError):  # longer than the kernel supports: loud, no fallback
        lib.call("mvptr_attn_fwd", qkv, 3 * H, m, ctx, H, None, 1, 257, 2, H, 0.0, 0)
    with pytest.raises(lib.MvptrError):  # head size != 64
        lib.call("mvptr_attn_fwd", qkv, 3 * H, m, ctx, H, None, 1, 8, 4, H, 0.0, 0)


def test_single_pair_matches_oracle():
    """B = 1: every kernel runs with a single sequence (one M tile, one attention CTA per head)."""
    cfg = O.Cfg(**CFG)
    sd = O.random_state_dict(cfg, "rep", seed=4)
    B, La, Lt, R = 1, 11, 4, 7
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=31, ragged=True)
    with torch.no_grad():
        seq_r, pooled_r, (txt_r, vis_r) = O.rep_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"],
                                                        b["attention_mask_a"], max_tag_length=Lt,
                                                        input_ids_b=b["input_ids_b"], token_type_ids_b=b["token_type_ids_b"],
                                                        attention_mask_b=b["attention_mask_b"], img_feats=b["img_feats"])
        model = P.build("BiImageBertRep", cfg, sd)
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **P.to_cuda(b))
    jm = torch.cat([b["attention_mask_a"], b["attention_mask_b"][:, Lt:]], 1)
    P.valid_rows_close(txt, txt_r, b["attention_mask_a"], 2e-2, 2e-2, "txt (B=1)")
    P.valid_rows_close(vis, vis_r, b["attention_mask_b"], 2e-2, 2e-2, "vis (B=1)")
    P.valid_rows_close(seq, seq_r, jm, 2e-2, 3e-2, "seq (B=1)")
    P.close(pooled, pooled_r, 2e-2, 2e-2, "pooled (B=1)")


def test_image_without_valid_regions_matches_oracle():
    """One image whose regions are ALL padding (zero features, mask 0) and one caption of a single token: the
    valid positions must still agree with the oracle (padding rows are unspecified)."""
    cfg = O.Cfg(**CFG)
    sd = O.random_state_dict(cfg, "rep", seed=5)
    B, La, Lt, R = 3, 9, 4, 6
    b = O.synthetic_batch(cfg, B, La, Lt, R, seed=32, ragged=True)
    b["attention_mask_b"][1, Lt:] = 0
    b["img_feats"][1] = 0
    b["attention_mask_a"][2, 1:] = 0
    b["input_ids_a"][2, 1:] = 0
    with torch.no_grad():
        seq_r, pooled_r, (txt_r, vis_r) = O.rep_forward(sd, cfg, b["input_ids_a"], b["token_type_ids_a"],
                                                        b["attention_mask_a"], max_tag_length=Lt,
                                                        input_ids_b=b["input_ids_b"], token_type_ids_b=b["token_type_ids_b"],
                                                        attention_mask_b=b["attention_mask_b"], img_feats=b["img_feats"])
        model = P.build("BiImageBertRep", cfg, sd)
        seq, pooled, (txt, vis) = model(max_tag_length=Lt, **P.to_cuda(b))
    assert torch.isfinite(seq.float()).all() and torch.isfinite(pooled.float()).all()
    jm = torch.cat([b["attention_mask_a"], b["attention_mask_b"][:, Lt:]], 1)
    P.valid_rows_close(txt, txt_r, b["attention_mask_a"], 2e-2, 2e-2, "txt")
    P.valid_rows_close(vis, vis_r, b["attention_mask_b"], 2e-2, 2e-2, "vis")
    P.valid_rows_close(seq, seq_r, jm, 2e-2, 3e-2, "seq")
    P.close(pooled, pooled_r, 2e-2, 2e-2, "pooled")


def test_error_conventions_of_the_reference():
    from mvp_pytorch_b200 import _lib
    cfg = O.Cfg(**CFG)
    sd = O.random_state_dict(cfg, "rep", seed=6)
    B, La, Lt, R = 2, 8, 4, 5
    b = P.to_cuda(O.synthetic_batch(cfg, B, La, Lt, R, seed=33, ragged=False))
    model = P.build("BiImageBertRep", cfg, sd)
    with torch.no_grad():
        bad = dict(b)
        bad["attention_mask_a"] = torch.ones(B, La, La, La, dtype=torch.long, device="cuda")
        with pytest.raises(NotImplementedError):  # mask rank not in {2, 3}: modeling_vlbert.py:435
            model(max_tag_length=Lt, **bad)
        bad["attention_mask_a"] = torch.ones(B, La, La, dtype=torch.long, device="cuda")
        with pytest.raises(NotImplementedError):  # 3-D masks exist in the reference but not in the kernels: loud
            model(max_tag_length=Lt, **bad)
        with pytest.raises(NotImplementedError):  # head_mask is out of scope: loud, never ignored
            model(max_tag_length=Lt, head_mask=torch.ones(4, device="cuda"), **b)
        wrong = dict(b)
        wrong["attention_mask_b"] = b["attention_mask_b"][:, :-1]
        with pytest.raises(ValueError):
            model(max_tag_length=Lt, **wrong)
    # CPU tensors never reach a CPU fallback
    cpu_model = P.build("BiImageBertRep", cfg, sd).cpu()
    with pytest.raises((_lib.MvptrError, RuntimeError)):
        cpu_model(max_tag_length=Lt, **{k: v.cpu() for k, v in b.items()})
