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
