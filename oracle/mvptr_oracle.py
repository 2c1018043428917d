"""CPU oracle for the MVPTR two-stage encoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 restatement (plain torch CPU ops over a
``state_dict``) of the reference algorithm in
``/root/reference/oscar/modeling/modeling_vlbert.py`` and
``/root/reference/transformers/pytorch_transformers/modeling_bert.py``.
Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker / reported baseline -- never as the product path.  The product
(``mvp_pytorch_b200``) never imports anything from ``oracle/`` and raises
when its CUDA library is missing.

Parity pin: ``oracle/make_golden.py`` imports the REAL reference classes from
``/root/reference`` (in the authoring container), runs them and this
restatement on the same weights/inputs, asserts they agree to <=1e-5, and
writes the reference's outputs to ``tests/golden/*.pt``.
``tests/test_oracle_golden.py`` re-checks this file against those committed
fixtures everywhere (CPU).  The reference repo has no tests/golden vectors of
its own for this path, so the pin is "outputs of the
reference itself run here".

State-dict keys are the reference's (e.g.
``bert.txt_encoder.layer.0.attention.self.query.weight``).
"""
from __future__ import annotations

import math
import random
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# bf16-store mode.  The reference algorithm is fp32; the CUDA path computes every contraction with fp32
# accumulation but STORES activations (and their gradients) in bf16.  With ``bf16_stores()`` active, ``_r`` rounds a
# tensor to bf16 at exactly the points where the CUDA path writes one to HBM (and rounds its gradient on the way
# back, since the gradient of a stored activation is itself a stored bf16 tensor); all arithmetic stays fp32.
# Outside the context ``_r`` is the identity and this file is the plain fp32 restatement that the goldens pin
# (max |d| = 0.0 against the reference).  tests/ compare the CUDA path with THIS mode at rtol 1e-2 and report
# the bf16-vs-fp32 distance (a property of the storage type, not of the kernels) separately.
# --------------------------------------------------------------------------
_PRECISION = "fp32"


class _Bf16Store(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def _r(x):
    return _Bf16Store.apply(x) if _PRECISION == "bf16" else x


def _r_saved(x):
    """Pre-LayerNorm sums are rounded only when they are saved for backward (csrc/norm_embed.cu ln_fwd_kernel:
    'LN runs on the bf16 rounding of what backward will normalise'); inference normalises the fp32 sum."""
    return _r(x) if torch.is_grad_enabled() else x


class bf16_stores:
    """``with O.bf16_stores(): ...`` -- evaluate the oracle with the CUDA path's bf16 storage points."""

    def __enter__(self):
        global _PRECISION
        self._old, _PRECISION = _PRECISION, "bf16"
        return self

    def __exit__(self, *a):
        global _PRECISION
        _PRECISION = self._old


# --------------------------------------------------------------------------
# L0 blocks  (modeling_bert.py)
# --------------------------------------------------------------------------
def layer_norm(x, w, b, eps):
    """TF-style LayerNorm, eps inside the sqrt.  modeling_bert.py:242-246."""
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return w * ((x - mu) / torch.sqrt(var + eps)) + b


def gelu_erf(x):
    """Exact-erf GELU.  modeling_bert.py:142-148."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x, sd: SD, prefix: str):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def embeddings(sd: SD, pfx: str, ids, type_ids, eps, position_ids=None):
    """word + position + token-type gather, LN.  modeling_bert.py:262-277.
    Dropout is identity here (the oracle models eval / p=0)."""
    L = ids.shape[1]
    if position_ids is None:
        position_ids = torch.arange(L, dtype=torch.long).unsqueeze(0).expand_as(ids)
    if type_ids is None:
        type_ids = torch.zeros_like(ids)
    e = (sd[pfx + ".word_embeddings.weight"][ids]
         + sd[pfx + ".position_embeddings.weight"][position_ids]
         + sd[pfx + ".token_type_embeddings.weight"][type_ids])
    return _r(layer_norm(e, sd[pfx + ".LayerNorm.weight"], sd[pfx + ".LayerNorm.bias"], eps))


def self_attention(sd: SD, pfx: str, h, ext_mask, n_heads):
    """modeling_vlbert.py:63-103 with history_state=None, head_mask=None.
    ext_mask is additive, broadcastable to [B, nh, L, L]."""
    B, L, H = h.shape
    d = H // n_heads

    def split(t):  # transpose_for_scores, modeling_bert.py:299-303
        return t.view(B, L, n_heads, d).permute(0, 2, 1, 3)

    q = split(_r(linear(h, sd, pfx + ".query")))
    k = split(_r(linear(h, sd, pfx + ".key")))
    v = split(_r(linear(h, sd, pfx + ".value")))
    s = q @ k.transpose(-1, -2) / math.sqrt(d) + ext_mask
    if _PRECISION == "bf16":
        # csrc/attention.cu: exp(s - max) is rounded to bf16 for the PV tensor-core product, the fp32 row sum
        # normalises the 64-wide output afterwards
        e = torch.exp(s - s.max(dim=-1, keepdim=True)[0])
        ctx = (_r(e) @ v) / e.sum(dim=-1, keepdim=True)
    else:
        p = torch.softmax(s, dim=-1)
        ctx = p @ v
    ctx = ctx.permute(0, 2, 1, 3).contiguous().view(B, L, H)
    return _r(ctx)


def encoder_layer(sd: SD, pfx: str, h, ext_mask, n_heads, eps):
    """CaptionBertLayer: attention -> SelfOutput -> Intermediate -> Output.
    modeling_vlbert.py:191-199; modeling_bert.py:348-352, 394-397, 407-411."""
    ctx = self_attention(sd, pfx + ".attention.self", h, ext_mask, n_heads)
    a = _r(layer_norm(_r_saved(_r(linear(ctx, sd, pfx + ".attention.output.dense")) + h),
                      sd[pfx + ".attention.output.LayerNorm.weight"],
                      sd[pfx + ".attention.output.LayerNorm.bias"], eps))
    inter = _r(gelu_erf(linear(a, sd, pfx + ".intermediate.dense")))
    out = _r(layer_norm(_r_saved(_r(linear(inter, sd, pfx + ".output.dense")) + a),
                        sd[pfx + ".output.LayerNorm.weight"],
                        sd[pfx + ".output.LayerNorm.bias"], eps))
    return out


def encoder(sd: SD, pfx: str, h, ext_mask, n_layers, n_heads, eps, return_at_layer=None):
    """CaptionBertEncoder.forward, modeling_vlbert.py:134-178 (single mask)."""
    mid = None
    for i in range(n_layers):
        h = encoder_layer(sd, f"{pfx}.layer.{i}", h, ext_mask, n_heads, eps)
        if return_at_layer is not None and i == return_at_layer:
            mid = h
    return h, mid


def pooler(sd: SD, pfx: str, h):
    """tanh(h[:,0] W^T + b).  modeling_bert.py:468-474."""
    return _r(torch.tanh(linear(h[:, 0], sd, pfx + ".dense")))


def lm_head(sd: SD, pfx: str, x, eps, decoder_weight):
    """BertPredictionHeadTransform + decoder + bias.  modeling_bert.py:487-491, 513-516."""
    t = _r(gelu_erf(linear(x, sd, pfx + ".transform.dense")))
    t = _r(layer_norm(t, sd[pfx + ".transform.LayerNorm.weight"], sd[pfx + ".transform.LayerNorm.bias"], eps))
    return t @ decoder_weight.t() + sd[pfx + ".bias"]


# --------------------------------------------------------------------------
# L1 backbone  (modeling_vlbert.py:354-874)
# --------------------------------------------------------------------------
class Cfg:
    """Subset of BertConfig the path reads (modeling_bert.py:189-225 + the ad-hoc
    attributes listed in SURVEY.md Appendix A)."""

    def __init__(self, **kw):
        self.vocab_size = 86051
        self.only_word_size = 30522
        self.hidden_size = 768
        self.num_hidden_layers = 12
        self.num_attention_heads = 12
        self.intermediate_size = 3072
        self.max_position_embeddings = 512
        self.type_vocab_size = 2
        self.layer_norm_eps = 1e-12
        self.img_layer_norm_eps = 1e-12
        self.img_feature_dim = 2054
        self.use_img_layernorm = 1
        self.num_labels = 2
        self.qa_answer_size = 3129
        self.num_contrast_classes = 2
        self.loss_type = "sfmx"
        self.__dict__.update(kw)


def ext_mask(mask):
    """(1-mask)*-10000 broadcast to [B,1,1,L].  modeling_vlbert.py:430-460 (2-D masks only)."""
    if mask.dim() != 2:
        raise NotImplementedError
    return (1.0 - mask[:, None, None, :].to(torch.float32)) * -10000.0


def stage1(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a,
           input_ids_b, token_type_ids_b, attention_mask_b, img_feats, bert="bert"):
    """Uni-modal encoders.  modeling_vlbert.py:479-512."""
    eps = cfg.layer_norm_eps
    nl, nh = cfg.num_hidden_layers // 2, cfg.num_attention_heads
    if attention_mask_a is None:
        attention_mask_a = torch.ones_like(input_ids_a)
    if attention_mask_b is None:
        attention_mask_b = torch.ones_like(input_ids_b)
    ma, mb = ext_mask(attention_mask_a), ext_mask(attention_mask_b)
    ea = embeddings(sd, bert + ".embeddings", input_ids_a, token_type_ids_a, eps)
    eb = embeddings(sd, bert + ".embeddings", input_ids_b, token_type_ids_b, eps)
    if img_feats is not None:
        ie = _r(linear(img_feats.to(torch.float32), sd, bert + ".img_embedding"))  # :498
        if cfg.use_img_layernorm:
            ie = _r(layer_norm(ie, sd[bert + ".LayerNorm.weight"], sd[bert + ".LayerNorm.bias"],
                               cfg.img_layer_norm_eps))  # :499-500
        eb = torch.cat([eb, ie], dim=1)  # :506
    txt, _ = encoder(sd, bert + ".txt_encoder", ea, ma, nl, nh, eps)  # :509
    vis, _ = encoder(sd, bert + ".vis_encoder", eb, mb, nl, nh, eps)  # :512
    return txt, vis, ma, mb


def global_embeddings(sd: SD, txt, vis, bert="bert"):
    """normalize(cls @ proj).  modeling_vlbert.py:525-526 / :717-718."""
    gt = F.normalize(txt[:, 0, :] @ sd[bert + ".txt_proj"], p=2, dim=-1)
    gi = F.normalize(vis[:, 0, :] @ sd[bert + ".vis_proj"], p=2, dim=-1)
    return gt, gi


def hard_negative_indexes(sim_mat):
    """In-batch hardest negatives, hn_mod='hard'.  modeling_vlbert.py:530-534."""
    masked = sim_mat - 2 * torch.eye(sim_mat.shape[0], dtype=sim_mat.dtype)
    return masked.max(dim=1)[1], masked.max(dim=0)[1]  # hard_img_index, hard_txt_index


def bibert_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a=None, attention_mask_a=None,
                   max_tag_length=None, use_b=False, input_ids_b=None, token_type_ids_b=None,
                   attention_mask_b=None, img_feats=None, encode_hn=False, dice_index=None,
                   phrase_layer=None, bert="bert"):
    """BiBertImgModel.forward, modeling_vlbert.py:410-609.

    ``dice_index`` replaces the reference's ``torch.randperm`` draw (:556) so
    the oracle and the CUDA path consume identical random choices."""
    eps = cfg.layer_norm_eps
    nl, nh = cfg.num_hidden_layers // 2, cfg.num_attention_heads
    txt, vis, ma, mb = stage1(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                              input_ids_b, token_type_ids_b, attention_mask_b, img_feats, bert)
    cut = 1 if use_b else max_tag_length  # :514-519
    only_vis, only_vis_mask = vis[:, cut:, :], mb[:, :, :, cut:]
    gt, gi = global_embeddings(sd, txt, vis, bert)
    sim_mat = gt @ gi.t()  # :527

    hard_seq = hard_pooled = hti_full = hii_full = mid_hard = None
    if encode_hn:
        hard_img_index, hard_txt_index = hard_negative_indexes(sim_mat)
        n = txt.shape[0]
        if dice_index is None:
            dice_index = torch.randperm(n)
        first, second = dice_index[: n // 2], dice_index[n // 2:]
        # :542-566 -- text i with its hardest image, image j with its hardest text
        seq_a = torch.cat([txt[first], only_vis[hard_img_index[first]]], dim=1)
        msk_a = torch.cat([ma[first], only_vis_mask[hard_img_index[first]]], dim=-1)
        seq_b = torch.cat([txt[hard_txt_index[second]], only_vis[second]], dim=1)
        msk_b = torch.cat([ma[hard_txt_index[second]], only_vis_mask[second]], dim=-1)
        hard_in = torch.cat([seq_a, seq_b], dim=0)
        hard_mask = torch.cat([msk_a, msk_b], dim=0)
        ar = torch.arange(n)
        hti_full = torch.cat([ar[first], hard_txt_index[second]])
        hii_full = torch.cat([hard_img_index[first], ar[second]])
        hard_seq, mid_hard = encoder(sd, bert + ".mul_encoder", hard_in, hard_mask, nl, nh, eps, phrase_layer)
        hard_pooled = pooler(sd, bert + ".pooler", hard_seq)

    joint = torch.cat([txt, only_vis], dim=1)  # :586
    joint_mask = torch.cat([ma, only_vis_mask], dim=-1)  # :587
    seq, mid_joint = encoder(sd, bert + ".mul_encoder", joint, joint_mask, nl, nh, eps, phrase_layer)
    pooled = pooler(sd, bert + ".pooler", seq)  # :600
    outs = ((seq, pooled, hard_seq, hard_pooled), (txt, vis, sim_mat), (hti_full, hii_full))
    if phrase_layer is not None:
        outs = outs + ((mid_joint, mid_hard),)
    return outs


def forward_single(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                   input_ids_b, token_type_ids_b, attention_mask_b, img_feats, bert="bert"):
    """BiBertImgModel.forward_single, modeling_vlbert.py:611-723."""
    txt, vis, _, _ = stage1(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                            input_ids_b, token_type_ids_b, attention_mask_b, img_feats, bert)
    return global_embeddings(sd, txt, vis, bert)


# --------------------------------------------------------------------------
# L2 task heads / losses
# --------------------------------------------------------------------------
def cross_entropy(logits, labels, ignore_index=-1):
    return F.cross_entropy(logits, labels, ignore_index=ignore_index)


def vsc_loss(sim_mat, logit_scale):
    """Symmetric CE on exp(logit_scale)*sim.  modeling_vlbert.py:1238-1241."""
    m = sim_mat * logit_scale.exp()
    lab = torch.arange(m.shape[0])
    return (cross_entropy(m, lab) + cross_entropy(m.t(), lab)) / 2


def t2i_sim(sim, rand_index):
    """top-3 per phrase row, pick one by rand_index, mean.  modeling_vlbert.py:1543-1550.
    ``rand_index`` replaces the reference's torch.randint(0,3) draw."""
    if sim.shape[0] == 0:
        return torch.zeros((), dtype=sim.dtype)
    top = sim.topk(3, dim=1)[0]
    return top[torch.arange(top.shape[0]), rand_index].mean()


def wra_sample_loss(seq, phrase_index, img_index, neg_img, rand_pos, rand_neg, margin=0.2):
    """Weakly-supervised phrase grounding, phrase_mod='sample'.
    modeling_vlbert.py:1285-1300 + helpers :1502-1508, :1553-1596.

    neg_img[b]   : the image the reference draws with random.choice (:1573)
    rand_pos/neg : [B, max_phrases] ints in [0,3): the torch.randint draws (:1548)
    """
    B = seq.shape[0]
    pos, neg = [], []
    for b in range(B):
        p0, p1 = int(phrase_index[b, 0]), int(phrase_index[b, 1])
        ph = F.normalize(seq[b, p0:p1], p=2, dim=-1)
        i0, i1 = int(img_index[b, 0]), int(img_index[b, 1])
        own = F.normalize(seq[b, i0:i1], p=2, dim=-1)
        nb = int(neg_img[b])
        n0, n1 = int(img_index[nb, 0]), int(img_index[nb, 1])
        other = F.normalize(seq[nb, n0:n1], p=2, dim=-1)
        n_ph = p1 - p0
        pos.append(t2i_sim(ph @ own.t(), rand_pos[b, :n_ph]))
        neg.append(t2i_sim(ph @ other.t(), rand_neg[b, :n_ph]))
    pos, neg = torch.stack(pos), torch.stack(neg)
    loss = torch.clamp(neg + margin - pos, min=0)
    valid = (phrase_index[:, 1] - phrase_index[:, 0]) > 0
    return loss[valid].mean()


def pos_sims_only(seq, text_index, img_index, rand):
    """get_pos_sims, modeling_vlbert.py:1510-1527: per row, phrases against the regions of the SAME row."""
    out = []
    for b in range(text_index.shape[0]):
        p0, p1 = int(text_index[b, 0]), int(text_index[b, 1])
        i0, i1 = int(img_index[b, 0]), int(img_index[b, 1])
        ph = F.normalize(seq[b, p0:p1], p=2, dim=-1)
        im = F.normalize(seq[b, i0:i1], p=2, dim=-1)
        out.append(t2i_sim(ph @ im.t(), rand[b, : p1 - p0]))
    return torch.stack(out)


def wra_hard_loss(seq, hard_seq, phrase_index, img_index, hard_txt_index, hard_img_index, rand_pos, rand_neg,
                  margin=0.2):
    """phrase_mod='hard', modeling_vlbert.py:1270-1283: negatives are the phrases of the hard-negative text
    against the regions of the hard-negative image, read from the hard-negative sequences."""
    hard_phrase_index = phrase_index[hard_txt_index]
    hard_object_index = img_index[hard_img_index]
    pos = pos_sims_only(seq, phrase_index, img_index, rand_pos)
    neg = pos_sims_only(hard_seq, hard_phrase_index, hard_object_index, rand_neg)
    loss = torch.clamp(neg + margin - pos, min=0)
    valid = ((phrase_index[:, 1] - phrase_index[:, 0]) > 0) & ((hard_phrase_index[:, 1] - hard_phrase_index[:, 0]) > 0)
    return loss[valid].mean()


def pretrain_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, masked_lm_labels_a,
                     input_ids_b, token_type_ids_b, attention_mask_b, masked_lm_labels_b, img_feats,
                     max_tag_length=20, img_index=None, phrase_index=None, dice_index=None,
                     neg_img=None, rand_pos=None, rand_neg=None, phrase_mod="sample", qa_ans=None):
    """BiBertImgForPreTraining.forward (phrase_mod='sample'), modeling_vlbert.py:1218-1311.
    Returns (total, vis_mlm, retrieval, mlm, itm[, wra])."""
    eps = cfg.layer_norm_eps
    outs, single, hard = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                                        max_tag_length=max_tag_length, input_ids_b=input_ids_b,
                                        token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                                        img_feats=img_feats, encode_hn=True, dice_index=dice_index)
    seq, pooled, hard_seq, hard_pooled = outs
    txt, vis, sim_mat = single
    dec_w = sd["bert.embeddings.word_embeddings.weight"][: cfg.only_word_size]  # tie, :1212-1216
    H = cfg.hidden_size
    # visual-tag MLM :1231-1235
    vm = masked_lm_labels_b > -1
    vis_rows = vis[vm].reshape(-1, H)
    vis_mlm = cross_entropy(lm_head(sd, "half_mlm", vis_rows, eps, dec_w), masked_lm_labels_b[vm])
    # VSC :1238-1241
    retrieval = vsc_loss(sim_mat, sd["logit_scale"])
    # MLM :1244-1249
    lm = masked_lm_labels_a > -1
    rows = seq[:, : input_ids_a.shape[1]][lm].reshape(-1, H)
    mlm = cross_entropy(lm_head(sd, "cls.predictions", rows, eps, dec_w), masked_lm_labels_a[lm])
    # ITM: label 0 = matched pair, 1 = hard negative  :1247-1251
    rel = linear(torch.cat([pooled, hard_pooled], 0), sd, "cls.seq_relationship")
    itm_lab = torch.cat([torch.zeros(pooled.shape[0], dtype=torch.long),
                         torch.ones(hard_pooled.shape[0], dtype=torch.long)])
    itm = cross_entropy(rel, itm_lab)
    total = vis_mlm + retrieval + mlm + itm
    res = (vis_mlm, retrieval, mlm, itm)
    if qa_ans is not None:  # :1260-1264
        qa = cross_entropy(linear(pooled, sd, "qa_head"), qa_ans)
        total = total + qa
        res = res + (qa,)
    if phrase_index is not None:
        if phrase_mod == "hard":
            wra = wra_hard_loss(seq, hard_seq, phrase_index, img_index, hard[0], hard[1], rand_pos, rand_neg)
        else:
            wra = wra_sample_loss(seq, phrase_index, img_index, neg_img, rand_pos, rand_neg)
        total = total + wra
        return (total,) + res + (wra,)
    return (total,) + res


def classifier(sd: SD, x, prefix="classifier"):
    """nn.Linear, or the Linear-ReLU-Linear variant of config.classifier == 'mlp'
    (modeling_vlbert.py:1615-1629 / :1730-1744): keys classifier.weight | classifier.0 / classifier.2."""
    if prefix + ".0.weight" in sd:
        return linear(_r(F.relu(linear(x, sd, prefix + ".0"))), sd, prefix + ".2")
    return linear(x, sd, prefix)


def retrieval_train_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                            input_ids_b, token_type_ids_b, attention_mask_b, img_feats,
                            max_tag_length=20, dice_index=None):
    """BiImageBertForRetrieval.forward_train (linear classifier), modeling_vlbert.py:1659-1687.
    ITM labels here are 1 = matched (:1681), the opposite of pre-training."""
    outs, single, _ = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                                     max_tag_length=max_tag_length, input_ids_b=input_ids_b,
                                     token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                                     img_feats=img_feats, encode_hn=True, dice_index=dice_index)
    seq, pooled, hard_seq, hard_pooled = outs
    vsc = vsc_loss(single[2], sd["logit_scale"])
    logits = classifier(sd, torch.cat([pooled, hard_pooled], 0))
    labels = torch.cat([torch.ones(pooled.shape[0], dtype=torch.long),
                        torch.zeros(hard_pooled.shape[0], dtype=torch.long)])
    itm = cross_entropy(logits, labels)
    return vsc + itm, logits, vsc, itm, labels


def retrieval_fine_forward(sd: SD, cfg: Cfg, *args, **kw):
    """BiImageBertForRetrieval.forward_fine, modeling_vlbert.py:1699-1712: raw ITM logits [B,2]."""
    outs, _, _ = bibert_forward(sd, cfg, *args, encode_hn=False, **kw)
    return classifier(sd, outs[1])


def vqa_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, labels,
                input_ids_b, token_type_ids_b, attention_mask_b, img_feats, max_tag_length=20):
    """BiImageBertForVQA.forward with loss_type='bce', modeling_vlbert.py:1834-1870.
    Head = BertQAPredictionHead (modeling_bert.py:518-533) on sequence_output[:,0]."""
    outs, _, _ = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                                max_tag_length=max_tag_length, input_ids_b=input_ids_b,
                                token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                                img_feats=img_feats, encode_hn=False)
    cls_tok = outs[0][:, 0]
    logits = lm_head(sd, "cls.predictions", cls_tok, cfg.layer_norm_eps, sd["cls.predictions.decoder.weight"])
    if labels is None:
        return (logits,)
    # instance_bce_with_logits, modeling_vlbert.py:878-883
    loss = F.binary_cross_entropy_with_logits(logits, labels) * labels.size(1)
    return loss, logits


def rep_forward(sd: SD, cfg: Cfg, *args, **kw):
    """BiImageBertRep.forward, modeling_vlbert.py:2536-2557."""
    outs, single, _ = bibert_forward(sd, cfg, *args, encode_hn=False, **kw)
    return outs[0], outs[1], single[:2]


def mlm_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, input_ids_b, token_type_ids_b,
                attention_mask_b, img_feats, max_tag_length=20):
    """BiBertImgForMLM.forward, modeling_vlbert.py:2632-2645: prediction scores [n_mask, only_word_size] at the
    [MASK] (id 103) positions of the text part -- row-major masked_select order -- and the ITM logits [B, 2].
    The decoder is an UNTIED nn.Linear here (no tie_weights call in :2604-2617)."""
    outs, _, _ = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                                max_tag_length=max_tag_length, input_ids_b=input_ids_b,
                                token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                                img_feats=img_feats, encode_hn=False)
    seq, pooled = outs[0], outs[1]
    lm_mask = input_ids_a == 103
    rows = seq[:, : input_ids_a.shape[1]][lm_mask].reshape(-1, cfg.hidden_size)
    scores = lm_head(sd, "cls.predictions", rows, cfg.layer_norm_eps, sd["cls.predictions.decoder.weight"])
    return scores, linear(pooled, sd, "cls.seq_relationship")


def seqcls_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, labels, input_ids_b,
                   token_type_ids_b, attention_mask_b, img_feats, max_tag_length=20, use_b=False):
    """BiImageBertForSequenceClassification.forward, modeling_vlbert.py:1762-1798 (linear or mlp classifier,
    cross-entropy); use_b=True joins the text with vis[:, 1:] instead of vis[:, max_tag_length:] (:514-519)."""
    outs, _, _ = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                                max_tag_length=max_tag_length, use_b=use_b, input_ids_b=input_ids_b,
                                token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                                img_feats=img_feats, encode_hn=False)
    logits = classifier(sd, outs[1])
    if labels is None:
        return (logits,)
    return cross_entropy(logits.view(-1, cfg.num_labels), labels.view(-1), ignore_index=-100), logits


def negative_sampling_probs(sim_mat, logit):
    """hn_mod='sample', modeling_vlbert.py:535-540: the two row-wise multinomial distributions."""
    masked = (logit * sim_mat) - 10000 * torch.eye(sim_mat.shape[0], dtype=sim_mat.dtype)
    return F.softmax(masked, dim=1), F.softmax(masked.t(), dim=1)


def forward_joint(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, max_tag_length,
                  input_ids_b, token_type_ids_b, attention_mask_b, img_feats,
                  input_ids_b2, token_type_ids_b2, attention_mask_b2, img_feats2, bert="bert"):
    """BiBertImgModel.forward_joint, modeling_vlbert.py:725-869: text + two images ->
    mul_encoder over [text | regions 1 | regions 2] -> (sequence_output, pooled_output)."""
    eps = cfg.layer_norm_eps
    nl, nh = cfg.num_hidden_layers // 2, cfg.num_attention_heads
    txt, vis, ma, mb = stage1(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                              input_ids_b, token_type_ids_b, attention_mask_b, img_feats, bert)
    _, vis2, _, mb2 = stage1(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                             input_ids_b2, token_type_ids_b2, attention_mask_b2, img_feats2, bert)
    cut = max_tag_length
    joint = torch.cat([txt, vis[:, cut:], vis2[:, cut:]], dim=1)               # :852
    jmask = torch.cat([ma, mb[:, :, :, cut:], mb2[:, :, :, cut:]], dim=-1)    # :853
    seq, _ = encoder(sd, bert + ".mul_encoder", joint, jmask, nl, nh, eps, None)
    return seq, pooler(sd, bert + ".pooler", seq)


def ve_plus_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, labels,
                    input_ids_b, token_type_ids_b, attention_mask_b, img_feats, max_tag_length=20):
    """BiImageBertForSequenceClassificationPlus.forward (eval / dropout 0), modeling_vlbert.py:2029-2068,
    classifier='linear', cross-entropy loss."""
    outs, single, _ = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                                     max_tag_length=max_tag_length, input_ids_b=input_ids_b,
                                     token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                                     img_feats=img_feats, encode_hn=False)
    txt, vis, _ = single
    gt = txt[:, 0, :] @ sd["bert.txt_proj"]
    gi = vis[:, 0, :] @ sd["bert.vis_proj"]
    single_out = torch.cat([gt, gi, gi - gt, gi * gt], dim=1)
    hid = F.relu(linear(single_out, sd, "single_mapping.0"))
    single_hidden = linear(hid, sd, "single_mapping.2")
    logits = linear(torch.cat([outs[1], single_hidden], dim=1), sd, "classifier")
    if labels is None:
        return (logits,)
    return cross_entropy(logits.view(-1, cfg.num_labels), labels.view(-1), ignore_index=-100), logits


def re_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, labels, input_ids_b,
               token_type_ids_b, attention_mask_b, img_feats, max_tag_length=20, mod=1, phrase_layer=None):
    """BiImageBertForRE.forward (eval / dropout 0), modeling_vlbert.py:1913-1966."""
    res = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a, max_tag_length=max_tag_length,
                         input_ids_b=input_ids_b, token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                         img_feats=img_feats, encode_hn=False, phrase_layer=phrase_layer)
    seq = res[0][0] if phrase_layer is None else res[3][0]
    La = input_ids_a.shape[1]
    vis, cls_tok = seq[:, La:], seq[:, 0]
    label_mask = labels >= 0
    if mod == 1:
        logits = torch.bmm(F.normalize(vis, p=2, dim=-1), F.normalize(cls_tok, p=2, dim=-1).unsqueeze(-1)).squeeze(-1)
        loss = F.mse_loss(torch.masked_select(labels, label_mask), torch.masked_select(logits, label_mask))
    elif mod == 2:
        logits = torch.bmm(vis, cls_tok.unsqueeze(-1)).squeeze(-1)
        loss = F.binary_cross_entropy_with_logits(torch.masked_select(logits, label_mask),
                                                  torch.masked_select((labels >= 0.5).float(), label_mask))
        logits = torch.sigmoid(logits)
    elif mod == 3:
        logits = linear(vis, sd, "classifier").squeeze(-1)
        loss = F.binary_cross_entropy_with_logits(torch.masked_select(logits, label_mask),
                                                  torch.masked_select(labels, label_mask))
    else:
        raise NotImplementedError
    return loss, logits


def seqcls_mlp_forward(sd: SD, cfg: Cfg, input_ids_a, token_type_ids_a, attention_mask_a, labels,
                       input_ids_b, token_type_ids_b, attention_mask_b, img_feats, max_tag_length=20):
    """BiImageBertForSequenceClassification.forward with classifier='mlp' (:1730-1744, :1762-1798)."""
    outs, _, _ = bibert_forward(sd, cfg, input_ids_a, token_type_ids_a, attention_mask_a,
                                max_tag_length=max_tag_length, input_ids_b=input_ids_b,
                                token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                                img_feats=img_feats, encode_hn=False)
    logits = linear(F.relu(linear(outs[1], sd, "classifier.0")), sd, "classifier.2")
    if labels is None:
        return (logits,)
    return cross_entropy(logits.view(-1, cfg.num_labels), labels.view(-1), ignore_index=-100), logits


# --------------------------------------------------------------------------
# Retrieval scoring loop (run_retrieval.py)
# --------------------------------------------------------------------------
def topk_desc(scores: torch.Tensor, k: int):
    """Order used by ``np.argsort(x)[::-1][:k]`` (run_retrieval.py:487,506) made
    deterministic: descending score, ties broken by DESCENDING index (what the
    reversal of a stable ascending sort gives)."""
    n = scores.shape[-1]
    idx = torch.arange(n - 1, -1, -1)
    rev = scores.flip(-1)
    order = torch.sort(rev, dim=-1, descending=True, stable=True)[1]
    return idx[order][..., :k]


def coarse_candidates(img_emb, txt_emb, k_i2t, k_t2i):
    """full_sims = img_emb @ txt_emb^T (run_retrieval.py:739) then per-image top
    captions and per-caption top images (compute_ranks_coarse, :481-522)."""
    sims = img_emb @ txt_emb.t()
    return sims, topk_desc(sims, k_i2t), topk_desc(sims.t().contiguous(), k_t2i)


def rank_of_first_positive(scores, is_pos):
    """compute_ranks inner loop, run_retrieval.py:440-447: rank of the first
    positive in descending-score order (len if none)."""
    order = topk_desc(scores, scores.shape[-1])
    ranks = []
    for r in range(scores.shape[0]):
        hit = is_pos[r][order[r]].nonzero()
        ranks.append(int(hit[0]) if hit.numel() else scores.shape[1])
    return ranks


def itm_match_prob(logits):
    """softmax(logits)[:,1], run_retrieval.py:776-777 / :818-820."""
    return torch.softmax(logits.float(), dim=1)[:, 1]


# --------------------------------------------------------------------------
# AdamW (optimization.py:130-189), for the fused optimizer parity test
# --------------------------------------------------------------------------
def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, correct_bias=True):
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
    return p, m, v


# --------------------------------------------------------------------------
# Deterministic synthetic inputs and random-init weights
# --------------------------------------------------------------------------
def state_dict_keys(cfg: Cfg, head: str):
    """Names/shapes of every tensor of the reference classes (checkpoint format)."""
    H, I = cfg.hidden_size, cfg.intermediate_size
    shapes = {
        "bert.embeddings.word_embeddings.weight": (cfg.vocab_size, H),
        "bert.embeddings.position_embeddings.weight": (cfg.max_position_embeddings, H),
        "bert.embeddings.token_type_embeddings.weight": (cfg.type_vocab_size, H),
        "bert.embeddings.LayerNorm.weight": (H,), "bert.embeddings.LayerNorm.bias": (H,),
        "bert.txt_proj": (H, H), "bert.vis_proj": (H, H),
        "bert.img_embedding.weight": (H, cfg.img_feature_dim), "bert.img_embedding.bias": (H,),
        "bert.LayerNorm.weight": (H,), "bert.LayerNorm.bias": (H,),
        "bert.pooler.dense.weight": (H, H), "bert.pooler.dense.bias": (H,),
    }
    for enc in ("vis_encoder", "txt_encoder", "mul_encoder"):
        for i in range(cfg.num_hidden_layers // 2):
            p = f"bert.{enc}.layer.{i}"
            for nm in ("query", "key", "value"):
                shapes[f"{p}.attention.self.{nm}.weight"] = (H, H)
                shapes[f"{p}.attention.self.{nm}.bias"] = (H,)
            shapes[f"{p}.attention.output.dense.weight"] = (H, H)
            shapes[f"{p}.attention.output.dense.bias"] = (H,)
            shapes[f"{p}.attention.output.LayerNorm.weight"] = (H,)
            shapes[f"{p}.attention.output.LayerNorm.bias"] = (H,)
            shapes[f"{p}.intermediate.dense.weight"] = (I, H)
            shapes[f"{p}.intermediate.dense.bias"] = (I,)
            shapes[f"{p}.output.dense.weight"] = (H, I)
            shapes[f"{p}.output.dense.bias"] = (H,)
            shapes[f"{p}.output.LayerNorm.weight"] = (H,)
            shapes[f"{p}.output.LayerNorm.bias"] = (H,)

    def lm(pfx, n_out, with_decoder):
        shapes[f"{pfx}.bias"] = (n_out,)
        shapes[f"{pfx}.transform.dense.weight"] = (H, H)
        shapes[f"{pfx}.transform.dense.bias"] = (H,)
        shapes[f"{pfx}.transform.LayerNorm.weight"] = (H,)
        shapes[f"{pfx}.transform.LayerNorm.bias"] = (H,)
        if with_decoder:
            shapes[f"{pfx}.decoder.weight"] = (n_out, H)

    if head == "pretrain":
        lm("cls.predictions", cfg.only_word_size, False)
        lm("half_mlm", cfg.only_word_size, False)
        shapes["cls.seq_relationship.weight"] = (cfg.num_contrast_classes, H)
        shapes["cls.seq_relationship.bias"] = (cfg.num_contrast_classes,)
        shapes["qa_head.weight"] = (cfg.qa_answer_size, H)
        shapes["qa_head.bias"] = (cfg.qa_answer_size,)
        shapes["logit_scale"] = ()
    elif head == "retrieval":
        shapes["classifier.weight"] = (2, H)
        shapes["classifier.bias"] = (2,)
        shapes["logit_scale"] = ()
    elif head == "retrieval_mlp":  # config.classifier == 'mlp', modeling_vlbert.py:1622-1627
        shapes["classifier.0.weight"] = (2 * H, H)
        shapes["classifier.0.bias"] = (2 * H,)
        shapes["classifier.2.weight"] = (cfg.num_labels, 2 * H)
        shapes["classifier.2.bias"] = (cfg.num_labels,)
        shapes["logit_scale"] = ()
    elif head == "mlm":  # BiBertImgForMLM, modeling_vlbert.py:2604-2617: untied decoders
        lm("cls.predictions", cfg.only_word_size, True)
        lm("half_mlm", cfg.only_word_size, True)
        shapes["cls.seq_relationship.weight"] = (cfg.num_contrast_classes, H)
        shapes["cls.seq_relationship.bias"] = (cfg.num_contrast_classes,)
        shapes["logit_scale"] = ()
    elif head == "cls_linear":  # BiImageBertForSequenceClassification, default linear classifier
        shapes["classifier.weight"] = (cfg.num_labels, H)
        shapes["classifier.bias"] = (cfg.num_labels,)
    elif head == "vqa":
        lm("cls.predictions", cfg.num_labels, True)
    elif head == "rep":
        pass
    elif head == "ve":  # BiImageBertForSequenceClassificationPlus with classifier='linear', modeling_vlbert.py:1975-2010
        shapes["single_mapping.0.weight"] = (2 * H, 4 * H)
        shapes["single_mapping.0.bias"] = (2 * H,)
        shapes["single_mapping.2.weight"] = (H, 2 * H)
        shapes["single_mapping.2.bias"] = (H,)
        shapes["classifier.weight"] = (cfg.num_labels, 2 * H)
        shapes["classifier.bias"] = (cfg.num_labels,)
    elif head == "re":  # BiImageBertForRE, :1873-1903 (default linear classifier, num_labels from the config)
        shapes["classifier.weight"] = (cfg.num_labels, H)
        shapes["classifier.bias"] = (cfg.num_labels,)
    elif head == "cls_mlp":  # BiImageBertForSequenceClassification with classifier='mlp', :1730-1744
        hid = H * 2
        shapes["classifier.0.weight"] = (hid, H)
        shapes["classifier.0.bias"] = (hid,)
        shapes["classifier.2.weight"] = (cfg.num_labels, hid)
        shapes["classifier.2.bias"] = (cfg.num_labels,)
    else:
        raise ValueError(head)
    return shapes


def random_state_dict(cfg: Cfg, head: str, seed: int = 0, bf16_exact: bool = True) -> SD:
    """Random-init weights with the reference's statistics (N(0,0.02) linears /
    embeddings, LN ~ 1/0 perturbed so gamma/beta are exercised, biases small
    non-zero so bias paths are exercised).  Values are rounded to bf16-exact
    floats so that a bf16 CUDA model holds *identical* weights."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_keys(cfg, head).items():
        if k == "logit_scale":
            t = torch.tensor(math.log(1 / 0.07))
        elif k.endswith("LayerNorm.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("LayerNorm.bias") or k.endswith(".bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        elif k.endswith("_proj"):
            t = cfg.hidden_size ** -0.5 * torch.randn(shp, generator=g)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        sd[k] = t.to(torch.bfloat16).to(torch.float32) if bf16_exact else t
    return sd


def synthetic_batch(cfg: Cfg, B, La, Lt, R, seed, ragged=True, n_phrase_max=5, with_labels=False):
    """Synthetic inputs of SURVEY.md section 8(d): word ids U[1000,only_word),
    last <=5 valid text slots are phrase concepts U[only_word,vocab), tags
    U[1000,only_word), regions N(0,1) (rounded to bf16-exact), ragged valid lengths."""
    g = torch.Generator().manual_seed(seed)
    ow, V = cfg.only_word_size, cfg.vocab_size
    lo = min(1000, ow // 2)
    ids_a = torch.randint(lo, ow, (B, La), generator=g)
    ids_b = torch.randint(lo, ow, (B, Lt), generator=g)
    img = torch.randn(B, R, cfg.img_feature_dim, generator=g).to(torch.bfloat16).to(torch.float32)
    mask_a = torch.ones(B, La, dtype=torch.long)
    mask_b = torch.ones(B, Lt + R, dtype=torch.long)
    seg_a = torch.zeros(B, La, dtype=torch.long)
    seg_b = torch.ones(B, Lt, dtype=torch.long)
    phrase_index = torch.zeros(B, 2, dtype=torch.long)
    img_index = torch.zeros(B, 2, dtype=torch.long)
    for b in range(B):
        n_txt = La if not ragged else int(torch.randint(max(4, La // 2), La + 1, (1,), generator=g))
        n_ph = int(torch.randint(0, n_phrase_max + 1, (1,), generator=g)) if ragged else n_phrase_max
        n_ph = min(n_ph, n_txt - 3)
        n_tag = Lt if not ragged else int(torch.randint(max(1, Lt // 3), Lt + 1, (1,), generator=g))
        n_reg = R if not ragged else int(torch.randint(max(3, R // 2), R + 1, (1,), generator=g))
        mask_a[b, n_txt:] = 0
        ids_a[b, n_txt:] = 0
        if n_ph > 0 and V > ow:
            ids_a[b, n_txt - n_ph:n_txt] = torch.randint(ow, V, (n_ph,), generator=g)
        phrase_index[b] = torch.tensor([n_txt - n_ph, n_txt])
        mask_b[b, n_tag:Lt] = 0
        ids_b[b, n_tag:] = 0
        mask_b[b, Lt + n_reg:] = 0
        img[b, n_reg:] = 0
        img_index[b] = torch.tensor([La, La + n_reg])
    batch = dict(input_ids_a=ids_a, token_type_ids_a=seg_a, attention_mask_a=mask_a,
                 input_ids_b=ids_b, token_type_ids_b=seg_b, attention_mask_b=mask_b, img_feats=img)
    if with_labels:
        lab_a = torch.full((B, La), -1, dtype=torch.long)
        lab_b = torch.full((B, Lt + R), -1, dtype=torch.long)
        pick_a = (torch.rand(B, La, generator=g) < 0.15) & (mask_a > 0) & (ids_a < ow)
        pick_a[:, 1] = True  # at least one label per row
        pick_b = (torch.rand(B, Lt, generator=g) < 0.15) & (mask_b[:, :Lt] > 0)
        pick_b[:, 0] = True
        lab_a[pick_a] = torch.randint(lo, ow, (int(pick_a.sum()),), generator=g)
        lab_b[:, :Lt][pick_b] = torch.randint(lo, ow, (int(pick_b.sum()),), generator=g)
        rnd = random.Random(seed)
        neg_img = torch.tensor([rnd.choice([j for j in range(B) if j != b]) for b in range(B)])
        batch.update(masked_lm_labels_a=lab_a, masked_lm_labels_b=lab_b, phrase_index=phrase_index,
                     img_index=img_index, dice_index=torch.randperm(B, generator=g), neg_img=neg_img,
                     rand_pos=torch.randint(0, 3, (B, n_phrase_max), generator=g),
                     rand_neg=torch.randint(0, 3, (B, n_phrase_max), generator=g))
    return batch
