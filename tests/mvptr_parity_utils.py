"""Shared helpers of the GPU parity tests: build the CUDA model and the oracle inputs from one
(cfg, seed) recipe."""
import torch

from oracle import mvptr_oracle as O


def make_config(cfg: O.Cfg, dropout=0.0, **extra):
    from mvp_pytorch_b200.modeling_utils import BertConfig
    c = BertConfig(vocab_size_or_config_json_file=cfg.vocab_size, hidden_size=cfg.hidden_size,
                   num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                   intermediate_size=cfg.intermediate_size, max_position_embeddings=cfg.max_position_embeddings,
                   type_vocab_size=cfg.type_vocab_size, layer_norm_eps=cfg.layer_norm_eps,
                   hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout)
    c.only_word_size = cfg.only_word_size
    c.qa_answer_size = cfg.qa_answer_size
    c.img_feature_dim = cfg.img_feature_dim
    c.img_feature_type = "faster_r-cnn"
    c.use_img_layernorm = cfg.use_img_layernorm
    c.img_layer_norm_eps = cfg.img_layer_norm_eps
    c.loss_type = cfg.loss_type
    c.num_labels = cfg.num_labels
    c.num_contrast_classes = cfg.num_contrast_classes
    for k, v in extra.items():
        setattr(c, k, v)
    return c


def build(cls_name, cfg, sd, dropout=0.0, train=False, **extra):
    import mvp_pytorch_b200.modeling_vlbert as mv
    model = getattr(mv, cls_name)(make_config(cfg, dropout, **extra))
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    model.train(train)
    return model


def to_cuda(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


def valid_rows_close(got, ref, mask, rtol, atol, what=""):
    """compare [B,L,H] tensors on rows where mask==1 (padded rows are unspecified)."""
    g = got.float().cpu()[mask.bool()]
    r = ref.float()[mask.bool()]
    err = (g - r).abs()
    tol = atol + rtol * r.abs()
    frac = (err > tol).float().mean().item()
    assert frac < 2e-3, f"{what}: {frac:.4%} elements beyond tol; max err {err.max():.4f} (ref absmax {r.abs().max():.3f})"
    return err.max().item()


def close(got, ref, rtol, atol, what=""):
    g, r = got.float().cpu(), ref.float()
    err = (g - r).abs()
    tol = atol + rtol * r.abs()
    assert (err <= tol).all(), f"{what}: max err {err.max():.5f} vs tol (ref absmax {r.abs().max():.4f}); got {g.flatten()[:4]} ref {r.flatten()[:4]}"
    return err.max().item()


def rel_l2(got, ref):
    g, r = got.float().cpu(), ref.float()
    return ((g - r).norm() / (r.norm() + 1e-12)).item()


# ------------------------------------------------------------------------------------------------------------
# north_star tolerances.  bf16 compute: logits / losses / gradients within rtol 1e-2 -- asserted against the
# oracle evaluated with the CUDA path's bf16 STORAGE points (oracle.bf16_stores(): same fp32 arithmetic as the
# reference, activations rounded to bf16 where the kernels write them).  The distance to the plain fp32 reference
# is reported next to it and bounded separately: it is a property of bf16 storage (measured on the CPU alone in
# tests/test_oracle_golden.py::test_bf16_storage_distance_to_fp32_reference), not of the kernels.
#
# Element-wise metric: |got - ref| / max(|ref|, rms(ref)) -- a relative error whose denominator is floored at the
# tensor's own typical magnitude, because the relative error of an element that happens to be ~0 is unbounded for
# ANY finite-precision implementation (including two fp32 runs with different summation order).
# ------------------------------------------------------------------------------------------------------------
RTOL_BF16 = 1e-2


def _lines():
    import conftest
    return conftest.PARITY_LINES


def rel_err(got, ref):
    g, r = got.detach().float().cpu(), ref.detach().float().cpu()
    if r.numel() == 0:
        return 0.0
    scale = r.pow(2).mean().sqrt().clamp_min(1e-30)
    return float(((g - r).abs() / torch.maximum(r.abs(), scale)).max())


def report(name, got, ref_bf16, ref_fp32=None, tol=RTOL_BF16, tol_fp32=None, metric=rel_err):
    """Assert `got` within `tol` of the bf16-store oracle; report (and optionally bound) the fp32-reference distance."""
    e16 = metric(got, ref_bf16)
    msg = f"{name}: vs bf16-store oracle {e16:.2e} (tol {tol:.0e})"
    e32 = None
    if ref_fp32 is not None:
        e32 = metric(got, ref_fp32)
        msg += f"; vs fp32 reference {e32:.2e}" + (f" (bound {tol_fp32:.0e})" if tol_fp32 else "")
    _lines().append(msg)
    print("[parity]", msg)
    assert e16 <= tol, msg
    if tol_fp32 is not None and e32 is not None:
        assert e32 <= tol_fp32, msg
    return e16, e32


def rows_err(got, ref, mask):
    """rel_err over the rows where mask == 1 (padded rows are unspecified)."""
    m = mask.bool().cpu()
    return rel_err(got.detach().float().cpu()[m], ref.detach().float().cpu()[m])


def grads_report(name, params, ref_bf16, ref_fp32=None, tol=RTOL_BF16, tol_fp32=None, floor=None):
    """Per-tensor relative L2 error of every gradient in ref_bf16 (dict name -> tensor).  Tensors whose reference
    gradient is (numerically) zero -- key biases: softmax is invariant to them -- are checked against `floor`,
    an absolute norm relative to the largest gradient norm of the model."""
    norms = {k: float(v.float().norm()) for k, v in ref_bf16.items()}
    top = max(norms.values())
    floor = 1e-4 * top if floor is None else floor
    worst16, worst32, wk16, wk32 = 0.0, 0.0, "", ""
    bad = []
    for k, r in ref_bf16.items():
        g = params[k].grad
        if norms[k] <= floor:
            if float(g.float().norm()) > 2 * floor:
                bad.append((k, "zero-gradient tensor", float(g.float().norm())))
            continue
        e = rel_l2(g, r)
        if e > worst16:
            worst16, wk16 = e, k
        if e > tol:
            bad.append((k, e))
        if ref_fp32 is not None and k in ref_fp32:
            e2 = rel_l2(g, ref_fp32[k])
            if e2 > worst32:
                worst32, wk32 = e2, k
            if tol_fp32 is not None and e2 > tol_fp32:
                bad.append((k, "fp32", e2))
    msg = f"{name}: {len(ref_bf16)} gradient tensors, worst relative L2 vs bf16-store oracle {worst16:.2e} ({wk16}) (tol {tol:.0e})"
    if ref_fp32 is not None:
        msg += f"; vs fp32 reference {worst32:.2e} ({wk32})" + (f" (bound {tol_fp32:.0e})" if tol_fp32 else "")
    _lines().append(msg)
    print("[parity]", msg)
    assert not bad, f"{name}: gradients beyond tolerance: {bad[:8]}"
