"""Drop-in model classes of the MVPTR two-stage encoder path.

Class names, constructor / ``forward`` signatures, output tuples, ``forward_mod`` switch,
error behaviour and ``state_dict`` keys mirror /root/reference/oscar/modeling/modeling_vlbert.py
(``BiBertImgModel`` :354-874, ``BiBertImgForPreTraining`` :1133-1311, ``BiImageBertForRetrieval``
:1598-1712, ``BiImageBertForSequenceClassification`` :1715-1798, ``BiImageBertForVQA`` :1801-1870,
``BiImageBertRep`` :2509-2557, ``BiBertImgForMLM`` :2559-2645) so run_pretrain_ml.py /
run_retrieval.py / run_vqa.py call them unchanged.  All arithmetic runs in hand-written
sm_100a CUDA behind the C-ABI (engine.py -> libmvptr_b200.so); there is no eager fallback.

Documented deviations:
* activations are bf16, so outputs are bf16 tensors (losses / logits fp32);
* rows of ``sequence_output`` at PADDED positions are computed like the reference computes
  them (same masked softmax) and are equally meaningless; parity is asserted on valid rows;
* ``head_mask``, ``encoder_history_states``, 3-D attention masks, ``output_attentions`` and
  the ``dis_code*`` feature types raise instead of silently taking another path.
"""
import copy
import logging
import random

import numpy as np
import torch
from torch import nn

from . import engine as E
from .modeling_bert import (BertEmbeddings, BertEncoder, BertLayer, BertLayerNorm, BertLMPredictionHead,
                            BertPooler, BertPreTrainedModel, BertQAPredictionHead, _Linear, _ParamOnly)
from .modeling_utils import BertConfig  # noqa: F401

logger = logging.getLogger(__name__)


class CaptionBertLayer(BertLayer):
    pass


class CaptionBertEncoder(BertEncoder):
    def __init__(self, config):
        super().__init__(config)
        self.num_layers = config.num_hidden_layers
        self.layer = nn.ModuleList([CaptionBertLayer(config) for _ in range(config.num_hidden_layers)])


def _check_unsupported(config, head_mask, encoder_history_states):
    if head_mask is not None:
        raise NotImplementedError("head_mask is not supported by the fused attention kernel")
    if encoder_history_states is not None:
        raise NotImplementedError("encoder_history_states is not supported by the fused attention kernel")
    if getattr(config, "output_attentions", False) or getattr(config, "output_hidden_states", False):
        raise NotImplementedError("output_attentions / output_hidden_states need the attention matrix that the "
                                  "fused kernel never materialises")


def _mask2d(mask, like):
    if mask is None:
        return torch.ones_like(like)
    if mask.dim() == 2:
        return mask.to(torch.int64).contiguous()
    if mask.dim() == 3:
        raise NotImplementedError("3-D attention masks are accepted by the reference (modeling_vlbert.py:433,449) "
                                  "but never produced by the in-scope scripts; the fused kernel takes key masks only")
    raise NotImplementedError  # same as the reference for any other rank (:435, :451)


def negative_sampling_probs(sim_mat, logit):
    """Row-wise sampling distributions of hn_mod='sample' (:535-540): softmax over logit * sim with the
    matched pair pushed down by 10000 -- (text -> image, image -> text).  [B, B] fp32, plumbing-sized."""
    n = sim_mat.shape[0]
    masked = (logit * sim_mat) - 10000 * torch.eye(n, dtype=sim_mat.dtype, device=sim_mat.device)
    return torch.softmax(masked, dim=1), torch.softmax(masked.t(), dim=1)


def sample_negatives(sim_mat, logit):
    """One multinomial draw per row from negative_sampling_probs: the same two torch.multinomial calls, in
    the same order, as the reference (:537-540)."""
    if logit is None:
        raise ValueError("hn_mod='sample' needs the temperature `logit` (the reference multiplies by it, :536)")
    p_t2i, p_i2t = negative_sampling_probs(sim_mat.float(), logit)
    hard_img_index = torch.multinomial(p_t2i, num_samples=1).squeeze(1)
    hard_txt_index = torch.multinomial(p_i2t, num_samples=1).squeeze(1)
    return hard_img_index, hard_txt_index


class BiBertImgModel(BertPreTrainedModel):
    """Two-stage encoder: uni-modal text / visual encoders, then the cross-modal encoder."""

    def __init__(self, config):
        super().__init__(config)
        self.embeddings = BertEmbeddings(config)
        half_config = copy.deepcopy(config)
        half_config.num_hidden_layers = half_config.num_hidden_layers // 2  # 2 phases (:361)
        self.vis_encoder = CaptionBertEncoder(half_config)
        self.txt_encoder = CaptionBertEncoder(half_config)
        self.mul_encoder = CaptionBertEncoder(half_config)
        self.pooler = BertPooler(config)
        scale = config.hidden_size ** -0.5
        self.txt_proj = nn.Parameter(scale * torch.randn(config.hidden_size, config.hidden_size))
        self.vis_proj = nn.Parameter(scale * torch.randn(config.hidden_size, config.hidden_size))
        self.img_dim = config.img_feature_dim
        self.img_feature_type = config.img_feature_type
        self.use_img_layernorm = getattr(config, "use_img_layernorm", None)
        if str(config.img_feature_type).startswith("dis_code"):
            raise NotImplementedError("dis_code* image feature types are out of scope of the CUDA path")
        self.img_embedding = _Linear(self.img_dim, config.hidden_size, bias=True)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        if self.use_img_layernorm:
            self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.img_layer_norm_eps)
        self.apply(self.init_weights)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    # ---- engine glue -----------------------------------------------------------------
    def _ctx(self):
        rt = self._rt if (self._rt is not None and self._rt_prefix and self._rt.arena.valid()) else self.runtime()
        return rt, self._rt_prefix

    def _stage1(self, rt, pf, input_ids_a, token_type_ids_a, attention_mask_a, position_ids_a, input_ids_b,
                token_type_ids_b, attention_mask_b, position_ids_b, img_feats):
        nl = self.config.num_hidden_layers // 2
        anchor = rt.anchor(self.txt_proj)
        save = torch.is_grad_enabled() and anchor.requires_grad
        ids_a = input_ids_a.to(torch.int64).contiguous()
        ids_b = input_ids_b.to(torch.int64).contiguous()
        seg_a = token_type_ids_a.to(torch.int64).contiguous() if token_type_ids_a is not None else None
        seg_b = token_type_ids_b.to(torch.int64).contiguous() if token_type_ids_b is not None else None
        pos_a = position_ids_a.to(torch.int64).contiguous() if position_ids_a is not None else None
        pos_b = position_ids_b.to(torch.int64).contiguous() if position_ids_b is not None else None
        mask_a = _mask2d(attention_mask_a, ids_a)
        if img_feats is None:
            raise NotImplementedError("the visual stream needs img_feats (text-only use is out of scope)")
        if attention_mask_b is None:
            attention_mask_b = torch.ones(ids_b.shape[0], ids_b.shape[1] + img_feats.shape[1], dtype=torch.int64,
                                          device=ids_b.device)
        mask_b = _mask2d(attention_mask_b, ids_b)
        if mask_b.shape[1] != ids_b.shape[1] + img_feats.shape[1]:
            raise ValueError(f"attention_mask_b covers {mask_b.shape[1]} positions, tags+regions = "
                             f"{ids_b.shape[1] + img_feats.shape[1]}")
        ma = E.mask_additive(rt, mask_a)
        mb = E.mask_additive(rt, mask_b)
        ea = E.EmbedFn.apply(ids_a, seg_a, pos_a, rt, pf + "embeddings", save, anchor)
        eb = E.VisInputFn.apply(ids_b, seg_b, pos_b, img_feats, rt, pf, save, anchor)
        txt = E.encoder(rt, pf + "txt_encoder", ea, ma, nl, anchor)
        vis = E.encoder(rt, pf + "vis_encoder", eb, mb, nl, anchor)
        return txt, vis, mask_a, mask_b

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, max_tag_length=None, use_b=False,
                position_ids_a=None, input_ids_b=None, token_type_ids_b=None, attention_mask_b=None,
                phrase_layer=None, position_ids_b=None, head_mask=None, img_feats=None,
                encoder_history_states=None, encode_hn=False, hn_mod='hard', logit=None):
        _check_unsupported(self.config, head_mask, encoder_history_states)
        rt, pf = self._ctx()
        rt.begin_forward(self.training)
        nl = self.config.num_hidden_layers // 2
        anchor = rt.anchor(self.txt_proj)
        txt, vis, mask_a, mask_b = self._stage1(rt, pf, input_ids_a, token_type_ids_a, attention_mask_a,
                                                position_ids_a, input_ids_b, token_type_ids_b, attention_mask_b,
                                                position_ids_b, img_feats)
        B = txt.shape[0]
        col0 = 1 if use_b else int(max_tag_length)  # :514-519
        global_txt = E.ClsProjNormFn.apply(txt, rt, pf + "txt_proj", anchor)
        global_img = E.ClsProjNormFn.apply(vis, rt, pf + "vis_proj", anchor)
        sim_mat = E.SimFn.apply(global_txt, global_img, rt)

        row_a = row_b = None
        hard_txt_index_full = hard_img_index_full = None
        if encode_hn:
            if hn_mod == 'hard':
                hard_img_index, hard_txt_index = E.hard_negatives(rt, sim_mat.detach())
            elif hn_mod == 'sample':
                hard_img_index, hard_txt_index = sample_negatives(sim_mat.detach(), logit)
            else:
                raise NotImplementedError
            n = B
            dice_index = torch.randperm(n, device=txt.device)  # same draw as :556
            ar = torch.arange(n, device=txt.device)
            first, second = dice_index[: n // 2], dice_index[n // 2:]
            hard_txt_index_full = torch.cat([ar[first], hard_txt_index[second]])
            hard_img_index_full = torch.cat([hard_img_index[first], ar[second]])
            # stage 2 runs ONCE over [joint pairs ; hard-negative pairs]
            row_a = torch.cat([ar, hard_txt_index_full]).contiguous()
            row_b = torch.cat([ar, hard_img_index_full]).contiguous()

        joint = E.ConcatRowsFn.apply(txt, vis, col0, row_a, row_b, rt)
        joint_mask = E.mask_additive(rt, mask_a, mask_b, col0, row_a, row_b)
        if phrase_layer is not None:
            mid = E.encoder(rt, pf + "mul_encoder", joint, joint_mask, nl, anchor, 0, phrase_layer + 1)
            seq_all = E.encoder(rt, pf + "mul_encoder", mid, joint_mask, nl, anchor, phrase_layer + 1, nl)
        else:
            mid = None
            seq_all = E.encoder(rt, pf + "mul_encoder", joint, joint_mask, nl, anchor)
        pooled_all = E.ClsDenseFn.apply(seq_all, rt, pf + "pooler.dense.weight", pf + "pooler.dense.bias", "tanh",
                                        anchor)
        if encode_hn:
            sequence_output, hard_encoder_outputs = seq_all[:B], seq_all[B:]
            pooled_output, hard_pooled_output = pooled_all[:B], pooled_all[B:]
            mid_joint, mid_hard = (mid[:B], mid[B:]) if mid is not None else (None, None)
        else:
            sequence_output, pooled_output, hard_encoder_outputs, hard_pooled_output = seq_all, pooled_all, None, None
            mid_joint, mid_hard = mid, None

        outputs = (sequence_output, pooled_output, hard_encoder_outputs, hard_pooled_output)
        single_stream_output = (txt, vis, sim_mat)
        hard_indexes = (hard_txt_index_full, hard_img_index_full)
        if phrase_layer is not None:
            return outputs, single_stream_output, hard_indexes, (mid_joint, mid_hard)
        return outputs, single_stream_output, hard_indexes

    def forward_single(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, max_tag_length=None,
                       position_ids_a=None, input_ids_b=None, token_type_ids_b=None, attention_mask_b=None,
                       position_ids_b=None, head_mask=None, img_feats=None, encoder_history_states=None):
        """Stage 1 only -> (normalised text embedding, normalised image embedding) (:611-723)."""
        _check_unsupported(self.config, head_mask, encoder_history_states)
        rt, pf = self._ctx()
        rt.begin_forward(self.training)
        txt, vis, _, _ = self._stage1(rt, pf, input_ids_a, token_type_ids_a, attention_mask_a, position_ids_a,
                                      input_ids_b, token_type_ids_b, attention_mask_b, position_ids_b, img_feats)
        anchor = rt.anchor(self.txt_proj)
        return (E.ClsProjNormFn.apply(txt, rt, pf + "txt_proj", anchor),
                E.ClsProjNormFn.apply(vis, rt, pf + "vis_proj", anchor))

    # ---- single-modality entry points used by the sharded retrieval scorer ---------------------
    def encode_text(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, position_ids_a=None):
        """Text stream only (embeddings + txt_encoder, :479, :509) -> (tokens [B,La,H], mask, normalised
        global embedding [B,H] fp32).  Stage 1 is pair independent, so a caption is
        encoded once however many images it is later paired with."""
        rt, pf = self._ctx()
        rt.begin_forward(self.training)
        nl = self.config.num_hidden_layers // 2
        anchor = rt.anchor(self.txt_proj)
        save = torch.is_grad_enabled() and anchor.requires_grad
        ids = input_ids_a.to(torch.int64).contiguous()
        seg = token_type_ids_a.to(torch.int64).contiguous() if token_type_ids_a is not None else None
        pos = position_ids_a.to(torch.int64).contiguous() if position_ids_a is not None else None
        mask = _mask2d(attention_mask_a, ids)
        emb = E.EmbedFn.apply(ids, seg, pos, rt, pf + "embeddings", save, anchor)
        txt = E.encoder(rt, pf + "txt_encoder", emb, E.mask_additive(rt, mask), nl, anchor)
        return txt, mask, E.ClsProjNormFn.apply(txt, rt, pf + "txt_proj", anchor)

    def encode_image(self, input_ids_b, token_type_ids_b=None, attention_mask_b=None, img_feats=None,
                     position_ids_b=None):
        """Visual stream only (tag embeddings + region projection + vis_encoder, :481-512)."""
        rt, pf = self._ctx()
        rt.begin_forward(self.training)
        nl = self.config.num_hidden_layers // 2
        anchor = rt.anchor(self.txt_proj)
        save = torch.is_grad_enabled() and anchor.requires_grad
        ids = input_ids_b.to(torch.int64).contiguous()
        seg = token_type_ids_b.to(torch.int64).contiguous() if token_type_ids_b is not None else None
        pos = position_ids_b.to(torch.int64).contiguous() if position_ids_b is not None else None
        if attention_mask_b is None:
            attention_mask_b = torch.ones(ids.shape[0], ids.shape[1] + img_feats.shape[1], dtype=torch.int64,
                                          device=ids.device)
        mask = _mask2d(attention_mask_b, ids)
        emb = E.VisInputFn.apply(ids, seg, pos, img_feats, rt, pf, save, anchor)
        vis = E.encoder(rt, pf + "vis_encoder", emb, E.mask_additive(rt, mask), nl, anchor)
        return vis, mask, E.ClsProjNormFn.apply(vis, rt, pf + "vis_proj", anchor)

    def forward_joint(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, max_tag_length=None,
                      position_ids_a=None, input_ids_b=None, token_type_ids_b=None, attention_mask_b=None,
                      position_ids_b2=None, input_ids_b2=None, token_type_ids_b2=None, attention_mask_b2=None,
                      img_feats2=None, position_ids_b=None, head_mask=None, img_feats=None,
                      encoder_history_states=None):
        """Two-image input (:725-869): the text and BOTH images go through stage 1 (the two visual streams
        as one batch of 2B sequences through the shared vis_encoder), stage 2 runs over
        [text | regions of image 1 | regions of image 2] -> (sequence_output, pooled_output)."""
        _check_unsupported(self.config, head_mask, encoder_history_states)
        rt, pf = self._ctx()
        rt.begin_forward(self.training)
        nl = self.config.num_hidden_layers // 2
        anchor = rt.anchor(self.txt_proj)
        Lt = int(max_tag_length)
        txt, mask_a, _ = self.encode_text(input_ids_a, token_type_ids_a, attention_mask_a, position_ids_a)
        B = txt.shape[0]
        if token_type_ids_b is None:   # the reference defaults BOTH tag segments to 0 (:745-749)
            token_type_ids_b = torch.zeros_like(input_ids_b)
        if token_type_ids_b2 is None:
            token_type_ids_b2 = torch.zeros_like(input_ids_b2)
        both = lambda u, v: None if u is None and v is None else torch.cat([u, v], dim=0)
        if (position_ids_b is None) != (position_ids_b2 is None):
            raise ValueError("forward_joint: give position ids for both images or for neither")
        def full_mask(m, ids, feats):  # default: every tag and every region is valid
            return m if m is not None else torch.ones(ids.shape[0], ids.shape[1] + feats.shape[1], dtype=torch.int64,
                                                      device=ids.device)
        mb1, mb2 = full_mask(attention_mask_b, input_ids_b, img_feats), full_mask(attention_mask_b2, input_ids_b2, img_feats2)
        vis12, mask12, _ = self.encode_image(torch.cat([input_ids_b, input_ids_b2], 0), both(token_type_ids_b, token_type_ids_b2),
                                             torch.cat([mb1, mb2], 0), torch.cat([img_feats, img_feats2], 0),
                                             both(position_ids_b, position_ids_b2))
        vis1, vis2 = vis12[:B], vis12[B:]
        m1, m2 = mask12[:B].contiguous(), mask12[B:].contiguous()
        j1 = E.ConcatRowsFn.apply(txt, vis1.contiguous(), Lt, None, None, rt)          # [text | regions 1]
        joint = E.ConcatRowsFn.apply(j1, vis2.contiguous(), Lt, None, None, rt)        # [... | regions 2]
        jmask_int = torch.cat([mask_a, m1[:, Lt:], m2[:, Lt:]], dim=1).contiguous()
        seq = E.encoder(rt, pf + "mul_encoder", joint, E.mask_additive(rt, jmask_int), nl, anchor)
        pooled = E.ClsDenseFn.apply(seq, rt, pf + "pooler.dense.weight", pf + "pooler.dense.bias", "tanh", anchor)
        return (seq, pooled)

    # ---- stage-2-only entry used by the retrieval scorer ---------------------------------
    def forward_stage2(self, txt, vis, mask_a, mask_b, max_tag_length, row_a, row_b):
        """mul_encoder + pooler over (text row_a[i], image row_b[i]) pairs built from CACHED
        stage-1 outputs.  Stage 1 is pair independent, so this reproduces forward_fine
         with 2.2x fewer FLOPs."""
        rt, pf = self._ctx()
        rt.begin_forward(self.training)
        nl = self.config.num_hidden_layers // 2
        anchor = rt.anchor(self.txt_proj)
        joint = E.ConcatRowsFn.apply(txt, vis, int(max_tag_length), row_a, row_b, rt)
        jm = E.mask_additive(rt, mask_a, mask_b, int(max_tag_length), row_a, row_b)
        seq = E.encoder(rt, pf + "mul_encoder", joint, jm, nl, anchor)
        pooled = E.ClsDenseFn.apply(seq, rt, pf + "pooler.dense.weight", pf + "pooler.dense.bias", "tanh", anchor)
        return seq, pooled


# --------------------------------------------------------------------------------------
# heads
# --------------------------------------------------------------------------------------
class BertPreTrainingHeads(_ParamOnly):
    def __init__(self, config, only_vocab=False, tied=True):
        super().__init__()
        self.predictions = BertLMPredictionHead(config, only_vocab=only_vocab, tied=tied)
        num_seq_relations = config.num_contrast_classes if hasattr(config, "num_contrast_classes") else 2
        self.seq_relationship = _Linear(config.hidden_size, num_seq_relations)


class BertVQAHeads(_ParamOnly):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertQAPredictionHead(config)


def _mlm_rows(labels2d, capacity=None, overflow=None):
    """Flat indexes and labels of the positions with a label > -1 (the masked_select of :1231-1235 /
    :1244-1249).

    Exact mode (capacity None): the dynamic size costs one host sync, so callers do this on the
    INPUT labels before any encoder work is queued and the host then runs ahead of the GPU.
    Capacity mode (CUDA-graph capture, graphs.py): a fixed number of rows, no sync -- unused slots
    point at row 0 with label -1 (ignored by the loss, zero gradient); if more labels than slots
    ever show up, `overflow` is raised on the device and the graphed step reports it."""
    flat = labels2d.reshape(-1)
    if capacity is None:
        idx = torch.nonzero(flat > -1).reshape(-1)
        return idx, flat[idx].contiguous()
    sel = flat > -1
    idx = torch.nonzero_static(sel, size=int(capacity), fill_value=0).reshape(-1)
    count = sel.sum()
    lab = torch.where(torch.arange(int(capacity), device=flat.device) < count, flat[idx], -1)
    if overflow is not None:
        overflow.logical_or_(count > capacity)
    return idx, lab.contiguous()


def _mlm_loss(rt, top, seq2d, rows, head_prefix, anchor):
    """selected rows -> LM head -> CE(ignore_index=-1) (:1231-1235, :1244-1249)."""
    cfg = top.config
    idx, labels = rows
    if idx.numel() == 0:
        return torch.full((), float("nan"), device=seq2d.device)  # CrossEntropy of an empty selection
    x = E.GatherRowsFn.apply(seq2d, idx, rt)
    t = E.HeadTransformFn.apply(x, rt, head_prefix + ".transform", anchor)
    return E.VocabCEFn.apply(t, labels, rt, "bert.embeddings.word_embeddings.weight", cfg.only_word_size,
                             head_prefix + ".bias", anchor)


def _draw_wra_choices(n_samples, max_phrases, device):
    """Random choices of the 'sample' phrase-grounding loss, drawn on the device without a
    host sync: per sample one OTHER image (the reference's random.choice over j != b,
    modeling_vlbert.py:1572-1573) and, per phrase, which of the top-3 regions to keep for the
    positive and the negative image (torch.randint(0, 3), :1548).  Same distributions as the
    reference; the batched kernel consumes them as index tensors."""
    r = torch.randint(0, max(n_samples - 1, 1), (n_samples,), device=device)
    neg_img = r + (r >= torch.arange(n_samples, device=device)).to(r.dtype)  # uniform over j != b
    if n_samples == 1:
        neg_img = torch.zeros(1, dtype=torch.int64, device=device)
    rand_pos = torch.randint(0, 3, (n_samples, max_phrases), device=device)
    rand_neg = torch.randint(0, 3, (n_samples, max_phrases), device=device)
    return neg_img, rand_pos, rand_neg


class BiBertImgForPreTraining(BertPreTrainedModel):
    """MLM + visual-tag MLM + VSC + ITM (+QA) + weakly-supervised phrase grounding (:1133-1311)."""

    def __init__(self, config):
        super().__init__(config)
        self.bert = BiBertImgModel(config)
        self.cls = BertPreTrainingHeads(config, only_vocab=True)
        self.half_mlm = BertLMPredictionHead(config, only_vocab=True)
        self.qa_head = _Linear(config.hidden_size, config.qa_answer_size)
        self.only_vocab_size = config.only_word_size
        self.num_seq_relations = config.num_contrast_classes if hasattr(config, "num_contrast_classes") else 2
        self.max_text_seq_length = config.max_text_seq_length if hasattr(config, "max_text_seq_length") else None
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.apply(self.init_weights)
        self.tie_weights()

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, masked_lm_labels_a=None, qa_ans=None,
                input_ids_b=None, token_type_ids_b=None, attention_mask_b=None, masked_lm_labels_b=None,
                max_tag_length=20, position_ids_a=None, position_ids_b=None, head_mask=None, img_feats=None,
                is_img_match=None, img_index=None, phrase_index=None, phrase_mod='sample', wra_choices=None):
        rt = self.runtime()
        self._adopt(self.bert, "bert.")
        anchor = rt.anchor(self.logit_scale)
        B, La = input_ids_a.shape
        Ltot = La + (attention_mask_b.shape[1] - max_tag_length if attention_mask_b is not None
                     else img_feats.shape[1])
        # label-only bookkeeping first (its host sync must not sit behind the queued forward)
        cap = getattr(self, "mlm_capacity", None)  # (tag rows, text rows) or None = exact
        ovf = getattr(self, "mlm_overflow", None)
        vis_rows = _mlm_rows(masked_lm_labels_b, cap[0] if cap else None, ovf)
        lab = torch.full((B, Ltot), -1, dtype=torch.int64, device=input_ids_a.device)
        lab[:, :La] = masked_lm_labels_a
        txt_rows = _mlm_rows(lab, cap[1] if cap else None, ovf)
        outputs, single_stream_output, hard_indexes = self.bert(
            input_ids_a=input_ids_a, position_ids_a=position_ids_a, token_type_ids_a=token_type_ids_a,
            attention_mask_a=attention_mask_a, head_mask=head_mask, img_feats=img_feats, input_ids_b=input_ids_b,
            position_ids_b=position_ids_b, token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
            max_tag_length=max_tag_length, encode_hn=True)
        txt, vis, sim_mat = single_stream_output
        sequence_output, pooled_output, hard_sequence_output, hard_pooled_output = outputs
        H = self.config.hidden_size
        assert sequence_output.shape[1] == Ltot

        # visual-tag MLM on the visual encoder output (:1231-1235)
        vis_mlm_loss = _mlm_loss(rt, self, vis.reshape(-1, H), vis_rows, "half_mlm", anchor)
        # VSC (:1238-1241)
        retrieval_loss, _ = E.VSCFn.apply(sim_mat, rt, "logit_scale", anchor)
        # MLM on the text part of the joint sequence (:1244-1249)
        masked_lm_loss = _mlm_loss(rt, self, sequence_output.reshape(-1, H), txt_rows, "cls.predictions", anchor)
        # ITM: 0 = matched, 1 = hard negative (:1247-1251)
        pooled_all = torch.cat([pooled_output, hard_pooled_output], dim=0)
        rel = E.SmallHeadFn.apply(pooled_all, rt, "cls.seq_relationship.weight", "cls.seq_relationship.bias", anchor)
        itm_labels = torch.cat([torch.zeros(B, dtype=torch.int64, device=rel.device),
                                torch.ones(hard_pooled_output.shape[0], dtype=torch.int64, device=rel.device)])
        next_sentence_loss = E.SmallCEFn.apply(rel, itm_labels, rt)

        total_loss = vis_mlm_loss + retrieval_loss + masked_lm_loss + next_sentence_loss
        outputs = (vis_mlm_loss, retrieval_loss, masked_lm_loss, next_sentence_loss)
        if qa_ans is not None:
            # qa_head + CrossEntropyLoss(ignore_index=-1) (:1224, :1262-1264): the same fused decoder + CE kernels
            # as the MLM heads, for any answer count
            qa_loss = E.VocabCEFn.apply(pooled_output, qa_ans.to(torch.int64).contiguous(), rt, "qa_head.weight",
                                        self.config.qa_answer_size, "qa_head.bias", anchor)
            total_loss = total_loss + qa_loss
            outputs = outputs + (qa_loss,)

        if phrase_index is not None:
            if phrase_mod == 'sample':
                maxp = E._lib.lib().mvptr_wra_max_phrases()
                if wra_choices is None:
                    wra_choices = _draw_wra_choices(B, maxp, sequence_output.device)
                neg_img, rand_pos, rand_neg = wra_choices
                pos_sims, neg_sims = E.WRAFn.apply(sequence_output, phrase_index.to(torch.int64).contiguous(),
                                                   img_index.to(torch.int64).contiguous(), neg_img.contiguous(),
                                                   rand_pos.contiguous(), rand_neg.contiguous(), rt)
                hinge = torch.clamp(neg_sims + 0.2 - pos_sims, min=0)
                valid = ((phrase_index[:, 1] - phrase_index[:, 0]) > 0).to(hinge.dtype)
                wra_loss = (hinge * valid).sum() / valid.sum()
                total_loss = total_loss + wra_loss
                outputs = (total_loss,) + outputs + (wra_loss,)
            elif phrase_mod == 'hard':
                # :1270-1283 -- positives from the matched sequences, negatives from the hard-negative sequences
                # (phrases of text hard_txt_index[i] against the regions of image hard_img_index[i]); both are the
                # "own image" half of the batched WRA kernel (negative image = the row itself, its output unused)
                maxp = E._lib.lib().mvptr_wra_max_phrases()
                if wra_choices is None:
                    wra_choices = _draw_wra_choices(B, maxp, sequence_output.device)
                _, rand_pos, rand_neg = wra_choices
                hard_txt_index, hard_img_index = hard_indexes
                pidx = phrase_index.to(torch.int64).contiguous()
                iidx = img_index.to(torch.int64).contiguous()
                hard_phrase_index = torch.index_select(pidx, 0, hard_txt_index).contiguous()
                hard_object_index = torch.index_select(iidx, 0, hard_img_index).contiguous()
                own = torch.arange(B, device=sequence_output.device)
                pos_sims, _ = E.WRAFn.apply(sequence_output, pidx, iidx, own, rand_pos.contiguous(),
                                            rand_pos.contiguous(), rt)
                neg_sims, _ = E.WRAFn.apply(hard_sequence_output, hard_phrase_index, hard_object_index, own,
                                            rand_neg.contiguous(), rand_neg.contiguous(), rt)
                hinge = torch.clamp(neg_sims + 0.2 - pos_sims, min=0)
                valid = (((pidx[:, 1] - pidx[:, 0]) > 0)
                         & ((hard_phrase_index[:, 1] - hard_phrase_index[:, 0]) > 0)).to(hinge.dtype)
                wra_loss = (hinge * valid).sum() / valid.sum()
                total_loss = total_loss + wra_loss
                outputs = (total_loss,) + outputs + (wra_loss,)
            else:
                raise NotImplementedError
        else:
            outputs = (total_loss,) + outputs
        return outputs


class BiImageBertForRetrieval(BertPreTrainedModel):
    """train / coarse / fine switch of :1598-1712."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = 2
        self.loss_type = config.loss_type
        self.bert = BiBertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.forward_mod = 'train'
        self.classifier = _make_classifier(config, config.hidden_size, config.num_labels)  # linear | mlp (:1615-1629)
        self.apply(self.init_weights)

    def reinit_cls_head(self):
        self.classifier.apply(self.init_weights)
        self.mark_weights_changed()

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, input_ids_b=None,
                token_type_ids_b=None, attention_mask_b=None, max_tag_length=20, position_ids_a=None,
                position_ids_b=None, head_mask=None, img_feats=None):
        kw = dict(input_ids_a=input_ids_a, token_type_ids_a=token_type_ids_a, attention_mask_a=attention_mask_a,
                  input_ids_b=input_ids_b, token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                  img_feats=img_feats, max_tag_length=max_tag_length, position_ids_a=position_ids_a,
                  position_ids_b=position_ids_b, head_mask=head_mask)
        if self.forward_mod == 'train':
            return self.forward_train(**kw)
        elif self.forward_mod == 'coarse':
            return self.forward_emb(**kw)
        elif self.forward_mod == 'fine':
            return self.forward_fine(**kw)
        else:
            raise NotImplementedError

    def _prep(self):
        rt = self.runtime()
        self._adopt(self.bert, "bert.")
        return rt

    def forward_train(self, **kw):
        rt = self._prep()
        anchor = rt.anchor(self.logit_scale)
        outputs, single_stream_output, _ = self.bert(encode_hn=True, **kw)
        sim_mat = single_stream_output[2]
        retrieval_loss, _ = E.VSCFn.apply(sim_mat, rt, "logit_scale", anchor)
        sequence_output, pooled_output, hard_sequence_output, hard_pooled_output = outputs
        pooled_all = torch.cat([pooled_output, hard_pooled_output], dim=0)
        if self.training and self.config.hidden_dropout_prob > 0:
            pooled_all = nn.functional.dropout(pooled_all, self.config.hidden_dropout_prob, True)  # :1680
        logits = _apply_classifier(self, rt, pooled_all, anchor)
        B = pooled_output.shape[0]
        labels = torch.cat([torch.ones(B, dtype=torch.int64, device=logits.device),
                            torch.zeros(hard_pooled_output.shape[0], dtype=torch.int64, device=logits.device)])
        next_sentence_loss = E.SmallCEFn.apply(logits, labels, rt)
        total_loss = retrieval_loss + next_sentence_loss
        return (total_loss, logits, retrieval_loss, next_sentence_loss, labels)

    def forward_emb(self, **kw):
        self._prep()
        kw.pop("max_tag_length", None)
        return self.bert.forward_single(**kw)

    def forward_fine(self, **kw):
        rt = self._prep()
        outputs, _, _ = self.bert(encode_hn=False, **kw)
        return _apply_classifier(self, rt, outputs[1], rt.anchor(self.logit_scale))


def _cls_loss(self, logits, labels, soft_label, num_labels):
    """Loss switch shared by the classification heads (:1778-1796 / :1850-1868) for the
    cases that are not fused with the decoder: tiny fp32 tensors, plumbing-level torch."""
    if num_labels == 1:
        return nn.functional.mse_loss(logits.view(-1), labels.to(torch.float).view(-1))
    if soft_label:
        lp = nn.functional.log_softmax(logits, dim=1)
        tgt = torch.stack([1 - labels.float(), labels.float()], dim=1)
        return -(tgt.view(tgt.shape[0], -1) * lp).sum(1).mean()
    if self.loss_type == 'kl':
        return nn.functional.kl_div(nn.functional.log_softmax(logits.contiguous().view(-1, 3129), dim=-1),
                                    labels.contiguous(), reduction="batchmean")
    if self.loss_type == 'bce':
        return nn.functional.binary_cross_entropy_with_logits(logits, labels) * labels.size(1)
    return nn.functional.cross_entropy(logits.view(-1, num_labels), labels.view(-1))


def _make_classifier(config, in_features, num_labels, default_in=None):
    """The classifier variants of :1730-1744 / :1996-2010 as parameter containers with the reference's keys
    (classifier.weight | classifier.0.weight + classifier.2.weight)."""
    if hasattr(config, 'classifier'):
        if not hasattr(config, 'cls_hidden_scale'):
            config.cls_hidden_scale = 2
        if config.classifier == 'linear':
            return _Linear(in_features, num_labels)
        if config.classifier == 'mlp':
            hid = config.hidden_size * config.cls_hidden_scale
            return nn.Sequential(_Linear(in_features, hid), nn.ReLU(), _Linear(hid, num_labels))
        raise NotImplementedError(f"classifier={config.classifier!r}")
    return _Linear(default_in if default_in is not None else in_features, num_labels)


def _apply_classifier(model, rt, x, anchor, prefix="classifier"):
    """logits = classifier(x) for the linear / mlp variants (fp32 [n, num_labels])."""
    mod = getattr(model, prefix)
    last = prefix
    if isinstance(mod, nn.Sequential):
        x = E.LinearFn.apply(x, rt, prefix + ".0.weight", prefix + ".0.bias", "relu", False, anchor)
        last = prefix + ".2"
    n_out = rt.arena.w(last + ".weight").shape[0]
    if n_out <= 64:
        return E.SmallHeadFn.apply(x.to(rt.adt), rt, last + ".weight", last + ".bias", anchor)
    return E.DecoderFn.apply(x.to(rt.adt), rt, last + ".weight", n_out, last + ".bias", anchor)


class BiImageBertForSequenceClassificationPlus(BertPreTrainedModel):
    """Visual entailment head (:1975-2070, run_ve.py:921): classifier(dropout([pooled ; single_mapping(dropout(
    [t ; v ; v - t ; v * t]))])) with t / v the un-normalised projections of the two stage-1 CLS tokens."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.loss_type = config.loss_type
        self.bert = BiBertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        H = config.hidden_size
        self.single_mapping = nn.Sequential(_Linear(4 * H, 2 * H), nn.ReLU(), _Linear(2 * H, H))
        self.classifier = _make_classifier(config, 2 * H, config.num_labels, default_in=H)
        self.apply(self.init_weights)

    def reinit_cls_head(self):
        self.classifier.apply(self.init_weights)
        self.mark_weights_changed()

    def freeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = False

    def unfreeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = True

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, labels=None, input_ids_b=None,
                token_type_ids_b=None, attention_mask_b=None, max_tag_length=20, position_ids_a=None,
                position_ids_b=None, head_mask=None, img_feats=None, soft_label=False):
        rt = self.runtime()
        self._adopt(self.bert, "bert.")
        anchor = rt.anchor(self.single_mapping[0].weight)
        outputs, (txt, vis, _), _ = self.bert(
            input_ids_a=input_ids_a, position_ids_a=position_ids_a, token_type_ids_a=token_type_ids_a,
            attention_mask_a=attention_mask_a, head_mask=head_mask, img_feats=img_feats, input_ids_b=input_ids_b,
            position_ids_b=position_ids_b, token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
            max_tag_length=max_tag_length, encode_hn=False)
        pooled = outputs[1]
        p_drop = self.config.hidden_dropout_prob if self.training else 0.0
        drop = (lambda t: nn.functional.dropout(t, p_drop, True)) if p_drop > 0 else (lambda t: t)
        gt = E.LinearFn.apply(txt[:, 0], rt, "bert.txt_proj", None, None, True, anchor).float()
        gi = E.LinearFn.apply(vis[:, 0], rt, "bert.vis_proj", None, None, True, anchor).float()
        single_out = torch.cat([gt, gi, gi - gt, gi * gt], dim=1)                               # :2041
        hid = E.LinearFn.apply(drop(single_out), rt, "single_mapping.0.weight", "single_mapping.0.bias", "relu",
                               False, anchor)
        single_hidden = E.LinearFn.apply(hid, rt, "single_mapping.2.weight", "single_mapping.2.bias", None, False,
                                         anchor)
        feats = drop(torch.cat([pooled, single_hidden], dim=1))                                 # :2044
        logits = _apply_classifier(self, rt, feats, anchor)
        out = (logits,) + outputs[2:]
        if labels is not None:
            out = (_cls_loss(self, logits, labels, soft_label, self.num_labels),) + out
        return out


class BiImageBertForSequenceClassification(BertPreTrainedModel):
    """classifier(dropout(pooled)) with the reference loss switch (:1715-1798)."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.loss_type = config.loss_type
        self.bert = BiBertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.classifier = _make_classifier(config, config.hidden_size, self.config.num_labels)
        self.apply(self.init_weights)

    def freeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = False

    def unfreeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = True

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, labels=None, input_ids_b=None,
                token_type_ids_b=None, attention_mask_b=None, max_tag_length=20, use_b=False, position_ids_a=None,
                position_ids_b=None, head_mask=None, img_feats=None, soft_label=False):
        rt = self.runtime()
        self._adopt(self.bert, "bert.")
        anchor = rt.anchor(next(self.classifier.parameters()))
        outputs, _, _ = self.bert(input_ids_a=input_ids_a, position_ids_a=position_ids_a,
                                  token_type_ids_a=token_type_ids_a, attention_mask_a=attention_mask_a,
                                  head_mask=head_mask, img_feats=img_feats, use_b=use_b, input_ids_b=input_ids_b,
                                  position_ids_b=position_ids_b, token_type_ids_b=token_type_ids_b,
                                  attention_mask_b=attention_mask_b, max_tag_length=max_tag_length, encode_hn=False)
        pooled = outputs[1]
        if self.training and self.config.hidden_dropout_prob > 0:
            pooled = nn.functional.dropout(pooled, self.config.hidden_dropout_prob, True)
        logits = _apply_classifier(self, rt, pooled, anchor)
        out = (logits,) + outputs[2:]
        if labels is not None:
            out = (_cls_loss(self, logits, labels, soft_label, self.num_labels),) + out
        return out


class BiImageBertForRE(BertPreTrainedModel):
    """Referring-expression head (:1873-1971, run_re.py:28): every region token is scored against [CLS] --
    mod 1 cosine similarity + MSE, mod 2 dot product + BCE on hard labels (returns sigmoid scores), mod 3 a
    linear classifier per region + BCE.  phrase_layer selects a mid-encoder output (:1921-1926)."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = 1
        self.loss_type = config.loss_type
        self.bert = BiBertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.classifier = _make_classifier(config, config.hidden_size, self.config.num_labels)
        self.apply(self.init_weights)

    def freeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = False

    def unfreeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = True

    def reinit_cls_head(self):
        self.classifier.apply(self.init_weights)
        self.mark_weights_changed()

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, labels=None, phrase_layer=None,
                input_ids_b=None, token_type_ids_b=None, attention_mask_b=None, max_tag_length=20, mod=1,
                position_ids_a=None, position_ids_b=None, head_mask=None, img_feats=None, soft_label=False):
        rt = self.runtime()
        self._adopt(self.bert, "bert.")
        anchor = rt.anchor(next(self.classifier.parameters()))
        res = self.bert(input_ids_a=input_ids_a, position_ids_a=position_ids_a, token_type_ids_a=token_type_ids_a,
                        attention_mask_a=attention_mask_a, head_mask=head_mask, img_feats=img_feats,
                        phrase_layer=phrase_layer, input_ids_b=input_ids_b, position_ids_b=position_ids_b,
                        token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
                        max_tag_length=max_tag_length, encode_hn=False)
        sequence_output = res[0][0] if phrase_layer is None else res[3][0]
        La = input_ids_a.shape[1]
        p_drop = self.config.hidden_dropout_prob if self.training else 0.0
        label_mask = labels >= 0
        if mod == 1:      # MSE on the cosine similarity (:1937-1944)
            logits = E.ClsRegionScoreFn.apply(sequence_output, La, True, p_drop, rt)
            loss = nn.functional.mse_loss(torch.masked_select(labels, label_mask).float(),
                                          torch.masked_select(logits, label_mask))
        elif mod == 2:    # BCE with [CLS] as the classifier (:1946-1952)
            raw = E.ClsRegionScoreFn.apply(sequence_output, La, False, p_drop, rt)
            hard_labels = (labels >= 0.5).float()
            loss = nn.functional.binary_cross_entropy_with_logits(torch.masked_select(raw, label_mask),
                                                                  torch.masked_select(hard_labels, label_mask))
            logits = torch.sigmoid(raw)
        elif mod == 3:    # per-region linear classifier (:1954-1958); the loss uses the SOFT labels, as the reference
            vis = sequence_output[:, La:]
            if p_drop > 0:
                vis = nn.functional.dropout(vis, p_drop, True)
            B, R, H = vis.shape
            logits = _apply_classifier(self, rt, vis.reshape(B * R, H), anchor).view(B, R)
            loss = nn.functional.binary_cross_entropy_with_logits(torch.masked_select(logits, label_mask),
                                                                  torch.masked_select(labels, label_mask).float())
        else:
            raise NotImplementedError
        return (loss, logits)


class BiImageBertForVQA(BertPreTrainedModel):
    """dropout(sequence_output[:,0]) -> BertQAPredictionHead -> loss switch (:1801-1870)."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.loss_type = config.loss_type
        self.bert = BiBertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.cls = BertVQAHeads(config)
        self.apply(self.init_weights)

    def freeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = False

    def unfreeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = True

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, labels=None, input_ids_b=None,
                token_type_ids_b=None, attention_mask_b=None, max_tag_length=20, position_ids_a=None,
                position_ids_b=None, head_mask=None, img_feats=None, soft_label=False):
        rt = self.runtime()
        self._adopt(self.bert, "bert.")
        anchor = rt.anchor(self.cls.predictions.bias)
        outputs, _, _ = self.bert(input_ids_a=input_ids_a, position_ids_a=position_ids_a,
                                  token_type_ids_a=token_type_ids_a, attention_mask_a=attention_mask_a,
                                  head_mask=head_mask, img_feats=img_feats, input_ids_b=input_ids_b,
                                  position_ids_b=position_ids_b, token_type_ids_b=token_type_ids_b,
                                  attention_mask_b=attention_mask_b, max_tag_length=max_tag_length, encode_hn=False)
        sequence_output = outputs[0]
        cls_tok = sequence_output[:, 0]
        if self.training and self.config.hidden_dropout_prob > 0:
            cls_tok = nn.functional.dropout(cls_tok, self.config.hidden_dropout_prob, True)  # :1844
        t = E.HeadTransformFn.apply(cls_tok.contiguous(), rt, "cls.predictions.transform", anchor)
        if labels is not None and self.num_labels != 1 and not soft_label and self.loss_type == 'bce':
            loss, logits_p = E.BCEFn.apply(t, labels, rt, "cls.predictions.decoder.weight", self.num_labels,
                                           "cls.predictions.bias", anchor)
            return (loss, logits_p[:, : self.num_labels]) + outputs[2:]
        logits = E.DecoderFn.apply(t, rt, "cls.predictions.decoder.weight", self.num_labels, "cls.predictions.bias",
                                   anchor)
        out = (logits,) + outputs[2:]
        if labels is not None:
            out = (_cls_loss(self, logits, labels, soft_label, self.num_labels),) + out
        return out


class BiImageBertRep(BertPreTrainedModel):
    """All-token contextual outputs for the inference pipeline (:2509-2557)."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = 1
        self.bert = BiBertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.apply(self.init_weights)

    def freeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = False

    def unfreeze_backbone(self):
        for param in self.bert.parameters():
            param.requires_grad = True

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, input_ids_b=None,
                token_type_ids_b=None, attention_mask_b=None, max_tag_length=20, position_ids_a=None,
                position_ids_b=None, head_mask=None, img_feats=None):
        self.runtime()
        self._adopt(self.bert, "bert.")
        outputs, single_stream_output, _ = self.bert(
            input_ids_a=input_ids_a, position_ids_a=position_ids_a, token_type_ids_a=token_type_ids_a,
            attention_mask_a=attention_mask_a, head_mask=head_mask, img_feats=img_feats, input_ids_b=input_ids_b,
            position_ids_b=position_ids_b, token_type_ids_b=token_type_ids_b, attention_mask_b=attention_mask_b,
            max_tag_length=max_tag_length, encode_hn=False)
        return outputs[0], outputs[1], single_stream_output[:2]


class BiBertImgForMLM(BertPreTrainedModel):
    """MLM logits at [MASK] (id 103) positions + ITM logits (:2559-2645); the decoder is NOT tied (:2616)."""

    def __init__(self, config):
        super().__init__(config)
        self.bert = BiBertImgModel(config)
        self.cls = BertPreTrainingHeads(config, only_vocab=True, tied=False)
        self.half_mlm = BertLMPredictionHead(config, only_vocab=True, tied=False)
        self.only_vocab_size = config.only_word_size
        self.num_seq_relations = config.num_contrast_classes if hasattr(config, "num_contrast_classes") else 2
        self.max_text_seq_length = config.max_text_seq_length if hasattr(config, "max_text_seq_length") else None
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.apply(self.init_weights)

    def forward(self, input_ids_a, token_type_ids_a=None, attention_mask_a=None, input_ids_b=None,
                token_type_ids_b=None, attention_mask_b=None, max_tag_length=20, position_ids_a=None,
                position_ids_b=None, head_mask=None, img_feats=None):
        rt = self.runtime()
        self._adopt(self.bert, "bert.")
        anchor = rt.anchor(self.logit_scale)
        outputs, _, _ = self.bert(input_ids_a=input_ids_a, position_ids_a=position_ids_a,
                                  token_type_ids_a=token_type_ids_a, attention_mask_a=attention_mask_a,
                                  head_mask=head_mask, img_feats=img_feats, input_ids_b=input_ids_b,
                                  position_ids_b=position_ids_b, token_type_ids_b=token_type_ids_b,
                                  attention_mask_b=attention_mask_b, max_tag_length=max_tag_length, encode_hn=False)
        sequence_output, pooled_output = outputs[0], outputs[1]
        B, La = input_ids_a.shape
        H = self.config.hidden_size
        Ltot = sequence_output.shape[1]
        pick = torch.zeros(B, Ltot, dtype=torch.bool, device=input_ids_a.device)
        pick[:, :La] = input_ids_a == 103
        idx = torch.nonzero(pick.reshape(-1)).reshape(-1)
        if idx.numel() == 0:  # no [MASK] in the batch: the reference's masked_select gives [0, V] scores
            scores = torch.empty(0, self.only_vocab_size, device=input_ids_a.device, dtype=torch.float32)
        else:
            rows = E.GatherRowsFn.apply(sequence_output.reshape(-1, H), idx, rt)
            t = E.HeadTransformFn.apply(rows, rt, "cls.predictions.transform", anchor)
            scores = E.DecoderFn.apply(t, rt, "cls.predictions.decoder.weight", self.only_vocab_size,
                                       "cls.predictions.bias", anchor)
        rel = E.SmallHeadFn.apply(pooled_output, rt, "cls.seq_relationship.weight", "cls.seq_relationship.bias", anchor)
        return scores, rel
