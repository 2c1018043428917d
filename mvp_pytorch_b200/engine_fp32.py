"""Host side of the fp32 VERIFICATION tier (csrc/fp32_tier.cu; enable with ``model.set_precision("fp32")``).

BASELINE.json north_star asks for "bit-exact ... masking and top-k ranking order under fp32" and "1e-4 in fp32":
the reference itself computes in fp32 (oscar/tmp_config_FP32.json; run_retrieval.py:1047 halves only on a flag).
This module mirrors engine.py Function for Function -- same names, same argument lists, so modeling_vlbert.py is
untouched -- with fp32 activations, the fp32 master weights as operands, and every contraction evaluated on the
tcgen05 GEMM as six bf16 products of 3-way operand splits (x = hi + mid + lo, 24 significand bits; the dropped
cross terms are <= 2^-24 of |a||b|).  Forward AND backward are implemented (gradients go to the same flat fp32
arena); dropout must be 0 (a verification tier compares against the reference's deterministic arithmetic).
It is not a performance path: ~6x the tensor work plus plain one-warp-per-row kernels.
"""
import torch
from torch.autograd import Function

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32
_PAIRS = ((0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1))  # hi.hi, hi.mid, mid.hi, hi.lo, lo.hi, mid.mid


def _pad8(n):
    return (n + 7) // 8 * 8


def split3(t, rows, cols, ld):
    """fp32 2-D array (rows x cols, element pitch ld, starting at t's data pointer) -> (hi, mid, lo) bf16
    [rows, pad8(cols)]."""
    p = _pad8(cols)
    out = torch.empty(3, rows, p, device=t.device, dtype=BF16)
    _lib.call("mvptr_f32_split3", t, ld, rows, cols, out[0], out[1], out[2], p)
    return out, p


def _wsplit(rt, key, t, rows, cols, ld):
    """Weight splits are made once per forward and reused by its backward (begin_forward clears the cache)."""
    c = rt.f32_cache.get(key)
    if c is None:
        c = rt.f32_cache[key] = split3(t, rows, cols, ld)
    return c


def gemm(A, B, D, M, N, K, *, ldd, a_mn=False, b_mn=False, accumulate=False, bias=None):
    """D[M,N] (+)= A . B^T (+ bias) with A, B given as (split, pitch) pairs from split3: six bf16 tensor-core
    products accumulated in fp32."""
    (As, pa), (Bs, pb) = A, B
    for n, (i, j) in enumerate(_PAIRS):
        _lib.gemm(As[i], Bs[j], D, M, N, K, lda=pa, ldb=pb, ldd=ldd, a_mn=a_mn, b_mn=b_mn,
                  accumulate=accumulate or n > 0, bias=bias if n == 0 else None)
    return D


def _act(x, act):
    _lib.call("mvptr_f32_act", x, x.numel(), {"gelu": 1, "tanh": 2, "relu": 3}[act])
    return x


def _act_bwd(dy, saved, act):
    dx = torch.empty_like(dy)
    _lib.call("mvptr_f32_act_bwd", dy, saved, dx, dy.numel(), {"gelu": 1, "tanh": 2, "relu": 3}[act])
    return dx


def _ln(x, residual, gamma, beta, y, rows_per_batch, batch_stride, save, rows, H, eps):
    dev = x.device
    pre = torch.empty(rows, H, device=dev, dtype=F32) if save else None
    st = torch.empty(2, rows, device=dev, dtype=F32) if save else None
    _lib.call("mvptr_f32_ln_fwd", x, residual, gamma, beta, y, rows_per_batch, batch_stride, pre,
              st[0] if save else None, st[1] if save else None, rows, H, eps)
    return pre, st


def _ln_bwd(rt, dy, rows_per_batch, batch_stride, pre, st, gname, bname, rows, H):
    a = rt.arena
    dx = torch.empty(rows, H, device=pre.device, dtype=F32)
    _lib.call("mvptr_f32_ln_bwd", dy, rows_per_batch, batch_stride, pre, st[0], st[1], a.master_of(gname), dx,
              a.g(gname), a.g(bname), rows, H)
    return dx


def _colsum(x, ldx, out, M, N):
    _lib.call("mvptr_f32_colsum", x, ldx, out, M, N)


def _linear_fwd(rt, x2d, ld_x, rows, wname, bias, n_out=None, w_rows=None):
    """y[rows, n_out] = x . W[:n_out]^T + bias (W = fp32 master [N, K])."""
    a = rt.arena
    W = a.master_of(wname)
    N, K = (W.shape[0] if n_out is None else n_out), W.shape[1]
    pitch = _pad8(N) if N % 4 else N
    y = torch.empty(rows, pitch, device=x2d.device, dtype=F32)
    gemm(split3(x2d, rows, K, ld_x), _wsplit(rt, wname, W, W.shape[0], K, K), y, rows, N, K, ldd=pitch, bias=bias)
    return y


def _linear_bwd(rt, dy, ld_dy, x2d, ld_x, rows, wname, bname_grad, n_out=None, need_dx=True):
    """db += colsum(dy); dW[:n_out] += dy^T x; returns dx = dy . W[:n_out]."""
    a = rt.arena
    W = a.master_of(wname)
    N, K = (W.shape[0] if n_out is None else n_out), W.shape[1]
    if bname_grad is not None:
        _colsum(dy, ld_dy, bname_grad, rows, N)
    dys = split3(dy, rows, N, ld_dy)
    gemm(dys, split3(x2d, rows, K, ld_x), a.g(wname), N, K, rows, ldd=K, a_mn=True, b_mn=True, accumulate=True)
    if not need_dx:
        return None
    dx = torch.empty(rows, K, device=dy.device, dtype=F32)
    gemm(dys, _wsplit(rt, wname, W, W.shape[0], K, K), dx, rows, K, N, ldd=K, b_mn=True)
    return dx


# ======================================================================================
# Encoder stack (modeling_vlbert.py:134-178, :191-199, :63-103; modeling_bert.py:348-352, 394-397, 407-411)
# ======================================================================================
class EncoderFn(Function):
    @staticmethod
    def forward(ctx, h, maskadd, rt, prefix, layer_lo, layer_hi, save, anchor):
        B, L, H = h.shape
        M, I, nh = B * L, rt.I, rt.nh
        a = rt.arena
        dev = h.device
        x = h.reshape(M, H).contiguous()
        saved = []
        for li in range(layer_lo, layer_hi):
            pf = f"{prefix}.layer.{li}."
            qo = a.offsets[pf + "attention.self.query.weight"][0]
            bo = a.offsets[pf + "attention.self.query.bias"][0]
            Wqkv = a.master[qo:qo + 3 * H * H].view(3 * H, H)  # q | k | v are adjacent in the arena
            bqkv = a.master[bo:bo + 3 * H]
            qkv = torch.empty(M, 3 * H, device=dev, dtype=F32)
            gemm(split3(x, M, H, H), _wsplit(rt, pf + "qkv", Wqkv, 3 * H, H, H), qkv, M, 3 * H, H, ldd=3 * H, bias=bqkv)
            att = torch.empty(M, H, device=dev, dtype=F32)
            probs = torch.empty(B, nh, L, L, device=dev, dtype=F32) if save else None
            _lib.call("mvptr_f32_attn_fwd", qkv, 3 * H, maskadd, att, H, probs, B, L, nh, H)
            t = _linear_fwd(rt, att, H, M, pf + "attention.output.dense.weight",
                            a.master_of(pf + "attention.output.dense.bias"))
            a1 = torch.empty(M, H, device=dev, dtype=F32)
            pre1, st1 = _ln(t, x, a.master_of(pf + "attention.output.LayerNorm.weight"),
                            a.master_of(pf + "attention.output.LayerNorm.bias"), a1, 0, 0, save, M, H, rt.eps)
            pre_g = _linear_fwd(rt, a1, H, M, pf + "intermediate.dense.weight", a.master_of(pf + "intermediate.dense.bias"))
            inter = _act(pre_g.clone() if save else pre_g, "gelu")
            t2 = _linear_fwd(rt, inter, I, M, pf + "output.dense.weight", a.master_of(pf + "output.dense.bias"))
            out = torch.empty(M, H, device=dev, dtype=F32)
            pre2, st2 = _ln(t2, a1, a.master_of(pf + "output.LayerNorm.weight"), a.master_of(pf + "output.LayerNorm.bias"),
                            out, 0, 0, save, M, H, rt.eps)
            if save:
                saved.append((pf, x, qkv, probs, att, pre1, st1, a1, pre_g, inter, pre2, st2))
            x = out
        ctx.rt, ctx.saved, ctx.dims = rt, saved, (B, L, H)
        return x.view(B, L, H)

    @staticmethod
    def backward(ctx, dout):
        rt, (B, L, H) = ctx.rt, ctx.dims
        M, I, nh = B * L, rt.I, rt.nh
        a = rt.arena
        dy = dout.reshape(M, H).contiguous()
        for (pf, x, qkv, probs, att, pre1, st1, a1, pre_g, inter, pre2, st2) in reversed(ctx.saved):
            dpre2 = _ln_bwd(rt, dy, 0, 0, pre2, st2, pf + "output.LayerNorm.weight", pf + "output.LayerNorm.bias", M, H)
            dinter = _linear_bwd(rt, dpre2, H, inter, I, M, pf + "output.dense.weight", a.g(pf + "output.dense.bias"))
            dpre_g = _act_bwd(dinter, pre_g, "gelu")
            da1 = dpre2.clone()  # residual branch; the dgrad below accumulates onto it
            W = a.master_of(pf + "intermediate.dense.weight")
            _colsum(dpre_g, I, a.g(pf + "intermediate.dense.bias"), M, I)
            dgs = split3(dpre_g, M, I, I)
            gemm(dgs, split3(a1, M, H, H), a.g(pf + "intermediate.dense.weight"), I, H, M, ldd=H, a_mn=True, b_mn=True,
                 accumulate=True)
            gemm(dgs, _wsplit(rt, pf + "intermediate.dense.weight", W, I, H, H), da1, M, H, I, ldd=H, b_mn=True,
                 accumulate=True)
            dpre1 = _ln_bwd(rt, da1, 0, 0, pre1, st1, pf + "attention.output.LayerNorm.weight",
                            pf + "attention.output.LayerNorm.bias", M, H)
            datt = _linear_bwd(rt, dpre1, H, att, H, M, pf + "attention.output.dense.weight",
                               a.g(pf + "attention.output.dense.bias"))
            dqkv = torch.empty(M, 3 * H, device=dy.device, dtype=F32)
            _lib.call("mvptr_f32_attn_bwd", qkv, 3 * H, probs, datt, H, dqkv, B, L, nh, H)
            qo = a.offsets[pf + "attention.self.query.weight"][0]
            bo = a.offsets[pf + "attention.self.query.bias"][0]
            Wqkv = a.master[qo:qo + 3 * H * H].view(3 * H, H)
            g = a.ensure_grad()
            a.touched.update(k for k in a.offsets if k.startswith(pf + "attention.self."))
            _colsum(dqkv, 3 * H, g[bo:bo + 3 * H], M, 3 * H)
            dqs = split3(dqkv, M, 3 * H, 3 * H)
            gemm(dqs, split3(x, M, H, H), g[qo:qo + 3 * H * H].view(3 * H, H), 3 * H, H, M, ldd=H, a_mn=True, b_mn=True,
                 accumulate=True)
            dx = dpre1.clone()
            gemm(dqs, _wsplit(rt, pf + "qkv", Wqkv, 3 * H, H, H), dx, M, H, 3 * H, ldd=H, b_mn=True, accumulate=True)
            dy = dx
        ctx.saved = None
        return dy.view(B, L, H), None, None, None, None, None, None, None


def encoder(rt, prefix, h, maskadd, n_layers, anchor, layer_lo=0, layer_hi=None):
    hi = n_layers if layer_hi is None else layer_hi
    save = torch.is_grad_enabled() and (h.requires_grad or anchor.requires_grad)
    return EncoderFn.apply(h, maskadd, rt, prefix, layer_lo, hi, save, anchor)


# ======================================================================================
# Input embeddings
# ======================================================================================
def _embed_fwd(rt, prefix, ids, type_ids, pos_ids, y, rows_per_batch, batch_stride, save):
    B, L = ids.shape
    a, cfg, H = rt.arena, rt.cfg, rt.H
    pre = torch.empty(B * L, H, device=ids.device, dtype=F32) if save else None
    st = torch.empty(2, B * L, device=ids.device, dtype=F32) if save else None
    m = a.master_of
    _lib.call("mvptr_f32_embed_ln_fwd", ids, type_ids, pos_ids, m(prefix + ".word_embeddings.weight"),
              m(prefix + ".position_embeddings.weight"), m(prefix + ".token_type_embeddings.weight"),
              m(prefix + ".LayerNorm.weight"), m(prefix + ".LayerNorm.bias"), y, rows_per_batch, batch_stride, pre,
              st[0] if save else None, st[1] if save else None, B, L, H, rt.eps, cfg.vocab_size,
              cfg.max_position_embeddings, cfg.type_vocab_size)
    return pre, st


def _embed_bwd(rt, prefix, dy, rows_per_batch, batch_stride, ids, type_ids, pos_ids, pre, st):
    if pos_ids is not None:
        raise NotImplementedError("backward through explicit position_ids is not supported")
    B, L = ids.shape
    a, H = rt.arena, rt.H
    dpre = _ln_bwd(rt, dy, rows_per_batch, batch_stride, pre, st, prefix + ".LayerNorm.weight", prefix + ".LayerNorm.bias",
                   B * L, H)
    _lib.call("mvptr_f32_embed_bwd", dpre, ids, type_ids, a.g(prefix + ".word_embeddings.weight"),
              a.g(prefix + ".position_embeddings.weight"), a.g(prefix + ".token_type_embeddings.weight"), B, L, H)


class EmbedFn(Function):
    @staticmethod
    def forward(ctx, ids, type_ids, pos_ids, rt, prefix, save, anchor):
        B, L = ids.shape
        y = torch.empty(B, L, rt.H, device=ids.device, dtype=F32)
        pre, st = _embed_fwd(rt, prefix, ids, type_ids, pos_ids, y, 0, 0, save)
        ctx.rt, ctx.s = rt, (prefix, ids, type_ids, pos_ids, pre, st)
        return y

    @staticmethod
    def backward(ctx, dy):
        prefix, ids, type_ids, pos_ids, pre, st = ctx.s
        _embed_bwd(ctx.rt, prefix, dy.contiguous(), 0, 0, ids, type_ids, pos_ids, pre, st)
        return (None,) * 7


class VisInputFn(Function):
    """modeling_vlbert.py:481-482, 498-506."""

    @staticmethod
    def forward(ctx, ids_b, type_b, pos_b, img_feats, rt, bert, save, anchor):
        B, Lt = ids_b.shape
        R, Kimg = img_feats.shape[1], img_feats.shape[2]
        H, a, cfg = rt.H, rt.arena, rt.cfg
        Lv = Lt + R
        out = torch.empty(B, Lv, H, device=ids_b.device, dtype=F32)
        emb = bert + "embeddings"
        pre_t, st_t = _embed_fwd(rt, emb, ids_b, type_b, pos_b, out, Lt, Lv * H, save)
        feats = img_feats.to(F32).contiguous().view(B * R, Kimg)
        if not cfg.use_img_layernorm:
            raise NotImplementedError("use_img_layernorm=0 is not supported by the CUDA path")
        xs = split3(feats, B * R, Kimg, Kimg)
        W = a.master_of(bert + "img_embedding.weight")
        pre_i = torch.empty(B * R, H, device=out.device, dtype=F32)
        gemm(xs, _wsplit(rt, bert + "img_embedding.weight", W, H, Kimg, Kimg), pre_i, B * R, H, Kimg, ldd=H,
             bias=a.master_of(bert + "img_embedding.bias"))
        img_rows = out.view(B * Lv, H)[Lt:]
        _, st_i = _ln(pre_i, None, a.master_of(bert + "LayerNorm.weight"), a.master_of(bert + "LayerNorm.bias"), img_rows, R,
                      Lv * H, save, B * R, H, float(cfg.img_layer_norm_eps))
        ctx.rt = rt
        ctx.s = (bert, ids_b, type_b, pos_b, pre_t, st_t, xs, pre_i, st_i, (B, Lt, R, Kimg))
        return out

    @staticmethod
    def backward(ctx, dout):
        rt = ctx.rt
        bert, ids_b, type_b, pos_b, pre_t, st_t, xs, pre_i, st_i, (B, Lt, R, Kimg) = ctx.s
        a, H = rt.arena, rt.H
        Lv = Lt + R
        dout = dout.contiguous()
        _embed_bwd(rt, bert + "embeddings", dout, Lt, Lv * H, ids_b, type_b, pos_b, pre_t, st_t)
        d_img_rows = dout.view(B * Lv, H)[Lt:]
        dpre = _ln_bwd(rt, d_img_rows, R, Lv * H, pre_i, st_i, bert + "LayerNorm.weight", bert + "LayerNorm.bias", B * R, H)
        _colsum(dpre, H, a.g(bert + "img_embedding.bias"), B * R, H)
        Kp = xs[1]
        scratch = torch.zeros(H, Kp, device=dout.device, dtype=F32)  # fp32 pitch 2054 * 4 B is not 16-byte aligned
        gemm(split3(dpre, B * R, H, H), xs, scratch, H, Kimg, B * R, ldd=Kp, a_mn=True, b_mn=True, accumulate=True)
        a.g(bert + "img_embedding.weight").add_(scratch[:, :Kimg])
        return (None,) * 8


# ======================================================================================
# Heads on the [CLS] rows
# ======================================================================================
class ClsProjNormFn(Function):
    """normalize(seq[:,0] @ proj), modeling_vlbert.py:525-526 / :717-718."""

    @staticmethod
    def forward(ctx, seq, rt, proj_name, anchor):
        seq = seq.contiguous()
        B, L, H = seq.shape
        P = rt.arena.master_of(proj_name)
        x32 = torch.empty(B, H, device=seq.device, dtype=F32)
        gemm(split3(seq, B, H, L * H), _wsplit(rt, proj_name, P, H, H, H), x32, B, H, H, ldd=H, b_mn=True)
        y32 = torch.empty(B, H, device=seq.device, dtype=F32)
        norm = torch.empty(B, device=seq.device, dtype=F32)
        _lib.call("mvptr_l2norm_fwd", x32, y32, None, norm, B, H)
        ctx.rt, ctx.s = rt, (seq, proj_name, y32, norm)
        return y32

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        seq, proj_name, y32, norm = ctx.s
        B, L, H = seq.shape
        a = rt.arena
        P = a.master_of(proj_name)
        dx = torch.empty(B, H, device=dy.device, dtype=F32)
        _lib.call("mvptr_f32_l2norm_bwd", dy.contiguous(), y32, norm, dx, B, H)
        dxs = split3(dx, B, H, H)
        gemm(split3(seq, B, H, L * H), dxs, a.g(proj_name), H, H, B, ldd=H, a_mn=True, b_mn=True, accumulate=True)
        dcls = torch.empty(B, H, device=dy.device, dtype=F32)
        gemm(dxs, _wsplit(rt, proj_name, P, H, H, H), dcls, B, H, H, ldd=H)
        dseq = torch.zeros_like(seq)
        dseq[:, 0] = dcls
        return dseq, None, None, None


class ClsDenseFn(Function):
    """act(seq[:, 0] W^T + b): BertPooler, modeling_bert.py:468-474."""

    @staticmethod
    def forward(ctx, seq, rt, wname, bname, act, anchor):
        seq = seq.contiguous()
        B, L, H = seq.shape
        a = rt.arena
        y = _linear_fwd(rt, seq, L * H, B, wname, a.master_of(bname))
        if act is not None:
            _act(y, act)
        ctx.rt, ctx.s = rt, (seq, wname, bname, act, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        seq, wname, bname, act, y = ctx.s
        B, L, H = seq.shape
        dpre = _act_bwd(dy.contiguous(), y, act) if act is not None else dy.contiguous()
        dcls = _linear_bwd(rt, dpre, y.shape[1], seq, L * H, B, wname, rt.arena.g(bname))
        dseq = torch.zeros_like(seq)
        dseq[:, 0] = dcls
        return dseq, None, None, None, None, None


class HeadTransformFn(Function):
    """LN(gelu(x W^T + b)), modeling_bert.py:487-491."""

    @staticmethod
    def forward(ctx, x, rt, prefix, anchor):
        x = x.contiguous()
        n, H = x.shape
        a = rt.arena
        pre = _linear_fwd(rt, x, H, n, prefix + ".dense.weight", a.master_of(prefix + ".dense.bias"))
        t = _act(pre.clone(), "gelu")
        y = torch.empty(n, H, device=x.device, dtype=F32)
        _, st = _ln(t, None, a.master_of(prefix + ".LayerNorm.weight"), a.master_of(prefix + ".LayerNorm.bias"), y, 0, 0,
                    True, n, H, rt.eps)
        ctx.rt, ctx.s = rt, (x, prefix, t, pre, st)
        return y

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        x, prefix, t, pre, st = ctx.s
        n, H = x.shape
        dt = _ln_bwd(rt, dy.contiguous(), 0, 0, t, st, prefix + ".LayerNorm.weight", prefix + ".LayerNorm.bias", n, H)
        dpre = _act_bwd(dt, pre, "gelu")
        dx = _linear_bwd(rt, dpre, H, x, H, n, prefix + ".dense.weight", rt.arena.g(prefix + ".dense.bias"))
        return dx, None, None, None


def _master_span(a, name, n):
    o = a.offsets[name][0]
    return a.master[o:o + n]


def decoder_logits(rt, t, wname, n_out, bias_name):
    """fp32 logits [n, pad8(n_out)] = t W[:n_out]^T + bias  (modeling_bert.py:514 / :531)."""
    n, H = t.shape
    a = rt.arena
    pitch = _pad8(n_out)
    W = a.master_of(wname)
    logits = torch.zeros(n, pitch, device=t.device, dtype=F32)
    gemm(split3(t, n, H, H), _wsplit(rt, wname, W, W.shape[0], H, H), logits, n, n_out, H, ldd=pitch,
         bias=_master_span(a, bias_name, n_out))
    return logits


def decoder_backward(rt, t, dlogits, wname, n_out, bias_name):
    n, H = t.shape
    a = rt.arena
    pitch = dlogits.shape[1]
    W = a.master_of(wname)
    _colsum(dlogits, pitch, a.g_span(bias_name, n_out), n, n_out)
    ds = split3(dlogits, n, n_out, pitch)
    gemm(ds, split3(t, n, H, H), a.g(wname), n_out, H, n, ldd=H, a_mn=True, b_mn=True, accumulate=True)
    dt = torch.empty(n, H, device=t.device, dtype=F32)
    gemm(ds, _wsplit(rt, wname, W, W.shape[0], H, H), dt, n, H, n_out, ldd=H, b_mn=True)
    return dt


class VocabCEFn(Function):
    @staticmethod
    def forward(ctx, t, labels, rt, wname, n_out, bias_name, anchor):
        t = t.contiguous()
        n = t.shape[0]
        logits = decoder_logits(rt, t, wname, n_out, bias_name)
        lse = torch.empty(n, device=t.device, dtype=F32)
        acc = torch.zeros(2, device=t.device, dtype=F32)
        _lib.call("mvptr_ce_fwd", logits, logits.shape[1], labels, n, n_out, -1, lse, acc[0:1], acc[1:2])
        ctx.rt, ctx.s = rt, (t, labels, wname, n_out, bias_name, logits, lse, acc)
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, g):
        rt = ctx.rt
        t, labels, wname, n_out, bias_name, logits, lse, acc = ctx.s
        n, pitch = logits.shape
        dlogits = torch.empty(n, pitch, device=t.device, dtype=F32)
        _lib.call("mvptr_f32_ce_bwd", logits, pitch, labels, n, n_out, -1, lse, acc[1:2], g.reshape(1).to(F32).contiguous(),
                  dlogits, pitch)
        return decoder_backward(rt, t, dlogits, wname, n_out, bias_name), None, None, None, None, None, None


class DecoderFn(Function):
    @staticmethod
    def forward(ctx, t, rt, wname, n_out, bias_name, anchor):
        t = t.contiguous()
        logits = decoder_logits(rt, t, wname, n_out, bias_name)
        ctx.rt, ctx.s = rt, (t, wname, n_out, bias_name, logits.shape[1])
        return logits[:, :n_out]

    @staticmethod
    def backward(ctx, dl):
        rt = ctx.rt
        t, wname, n_out, bias_name, pitch = ctx.s
        d = torch.zeros(t.shape[0], pitch, device=t.device, dtype=F32)
        d[:, :n_out] = dl
        return decoder_backward(rt, t, d, wname, n_out, bias_name), None, None, None, None, None


class BCEFn(Function):
    """instance_bce_with_logits (modeling_vlbert.py:878-883) on the answer decoder."""

    @staticmethod
    def forward(ctx, t, labels, rt, wname, n_out, bias_name, anchor):
        t = t.contiguous()
        n = t.shape[0]
        logits = decoder_logits(rt, t, wname, n_out, bias_name)
        lab = labels.to(F32).contiguous()
        loss = torch.zeros(1, device=t.device, dtype=F32)
        _lib.call("mvptr_bce_fwd", logits, logits.shape[1], lab, n, n_out, loss)
        ctx.rt, ctx.s = rt, (t, lab, wname, n_out, bias_name, logits)
        ctx.mark_non_differentiable(logits)
        return loss[0], logits

    @staticmethod
    def backward(ctx, g, _):
        rt = ctx.rt
        t, lab, wname, n_out, bias_name, logits = ctx.s
        n, pitch = logits.shape
        dlogits = torch.empty(n, pitch, device=t.device, dtype=F32)
        _lib.call("mvptr_f32_bce_bwd", logits, pitch, lab, n, n_out, g.reshape(1).to(F32).contiguous(), dlogits, pitch)
        return decoder_backward(rt, t, dlogits, wname, n_out, bias_name), None, None, None, None, None, None


class LinearFn(Function):
    """y = act(x W^T + b), or y = x P for a projection stored [K, N] (kn=True)."""

    @staticmethod
    def forward(ctx, x, rt, wname, bname, act, kn, anchor):
        x = x.to(F32).contiguous()
        n, K = x.shape
        a = rt.arena
        W = a.master_of(wname)
        bias = a.master_of(bname) if bname is not None else None
        if kn:
            N = W.shape[1]
            y = torch.empty(n, N, device=x.device, dtype=F32)
            gemm(split3(x, n, K, K), _wsplit(rt, wname, W, K, N, N), y, n, N, K, ldd=N, b_mn=True, bias=bias)
        else:
            y = _linear_fwd(rt, x, K, n, wname, bias)
        if act is not None:
            _act(y, act)
        ctx.rt, ctx.s = rt, (x, wname, bname, act, kn, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        x, wname, bname, act, kn, y = ctx.s
        n, K = x.shape
        N = y.shape[1]
        a = rt.arena
        dpre = _act_bwd(dy.contiguous(), y, act) if act is not None else dy.to(F32).contiguous()
        if not kn:
            dx = _linear_bwd(rt, dpre, N, x, K, n, wname, a.g(bname) if bname is not None else None)
            return dx, None, None, None, None, None, None
        W = a.master_of(wname)
        if bname is not None:
            _colsum(dpre, N, a.g(bname), n, N)
        ds = split3(dpre, n, N, N)
        gemm(split3(x, n, K, K), ds, a.g(wname), K, N, n, ldd=N, a_mn=True, b_mn=True, accumulate=True)
        dx = torch.empty(n, K, device=dy.device, dtype=F32)
        gemm(ds, _wsplit(rt, wname, W, K, N, N), dx, n, K, N, ldd=K)
        return dx, None, None, None, None, None, None


class SmallHeadFn(Function):
    """x W^T + b with a handful of outputs (ITM / retrieval classifier) -> fp32 logits, exact fp32 FMA."""

    @staticmethod
    def forward(ctx, x, rt, wname, bname, anchor):
        x = x.to(F32).contiguous()
        n, H = x.shape
        a = rt.arena
        C = a.master_of(wname).shape[0]
        logits = torch.empty(n, C, device=x.device, dtype=F32)
        _lib.call("mvptr_f32_small_head_fwd", x, H, a.master_of(wname), a.master_of(bname), logits, n, H, C)
        ctx.rt, ctx.s = rt, (x, wname, bname, C)
        return logits

    @staticmethod
    def backward(ctx, dl):
        rt = ctx.rt
        x, wname, bname, C = ctx.s
        n, H = x.shape
        a = rt.arena
        dx = torch.empty(n, H, device=x.device, dtype=F32)
        _lib.call("mvptr_f32_small_head_bwd", dl.to(F32).contiguous(), x, H, a.master_of(wname), dx, a.g(wname), a.g(bname),
                  n, H, C)
        return dx, None, None, None, None


# ======================================================================================
# Similarity, gathers, WRA
# ======================================================================================
def sim_matrix(rt, a32, b32):
    """a . b^T [n, pad8(m)] (modeling_vlbert.py:527, run_retrieval.py:739)."""
    n, H = a32.shape
    m = b32.shape[0]
    pitch = _pad8(m)
    out = torch.empty(n, pitch, device=a32.device, dtype=F32)
    gemm(split3(a32.contiguous(), n, H, H), split3(b32.contiguous(), m, H, H), out, n, m, H, ldd=pitch)
    return out


class SimFn(Function):
    @staticmethod
    def forward(ctx, gt, gi, rt):
        sim = sim_matrix(rt, gt, gi)
        ctx.rt, ctx.s = rt, (gt, gi)
        return sim[:, : gi.shape[0]]

    @staticmethod
    def backward(ctx, dsim):
        gt, gi = ctx.s
        n, H = gt.shape
        m = gi.shape[0]
        pm = _pad8(m)
        d = torch.zeros(n, pm, device=gt.device, dtype=F32)
        d[:, :m] = dsim
        ds = split3(d, n, m, pm)
        dgt = torch.empty(n, H, device=gt.device, dtype=F32)
        dgi = torch.empty(m, H, device=gt.device, dtype=F32)
        gemm(ds, split3(gi.contiguous(), m, H, H), dgt, n, H, m, ldd=H, b_mn=True)
        gemm(ds, split3(gt.contiguous(), n, H, H), dgi, m, H, n, ldd=H, a_mn=True, b_mn=True)
        return dgt, dgi, None


class ConcatRowsFn(Function):
    """out[r] = cat(a[row_a[r]], b[row_b[r], col0:]) -- the bf16 gather kernel moves bytes, so an fp32 row is
    2H of its elements."""

    @staticmethod
    def forward(ctx, a3, b3, col0, row_a, row_b, rt):
        a3, b3 = a3.contiguous(), b3.contiguous()
        Ba, La, H = a3.shape
        Lb = b3.shape[1]
        rows = row_a.shape[0] if row_a is not None else Ba
        out = torch.empty(rows, La + Lb - col0, H, device=a3.device, dtype=F32)
        _lib.call("mvptr_concat_rows", a3, La, b3, Lb, col0, row_a, row_b, out, rows, 2 * H)
        ctx.s = (a3.shape, b3.shape, col0, row_a, row_b, rows)
        return out

    @staticmethod
    def backward(ctx, dout):
        sa, sb, col0, row_a, row_b, rows = ctx.s
        da = torch.zeros(sa, device=dout.device, dtype=F32)
        db = torch.zeros(sb, device=dout.device, dtype=F32)
        _lib.call("mvptr_f32_concat_rows_bwd", dout.contiguous(), sa[1], sb[1], col0, row_a, row_b, da, db, rows, sa[2])
        return da, db, None, None, None, None


class GatherRowsFn(Function):
    @staticmethod
    def forward(ctx, x2d, idx, rt):
        x2d = x2d.contiguous()
        n, H = idx.shape[0], x2d.shape[1]
        out = torch.empty(n, H, device=x2d.device, dtype=F32)
        _lib.call("mvptr_gather_rows", x2d, idx, out, n, 2 * H)
        ctx.s = (x2d.shape, idx)
        return out

    @staticmethod
    def backward(ctx, dout):
        shape, idx = ctx.s
        dx = torch.zeros(shape, device=dout.device, dtype=F32)
        _lib.call("mvptr_f32_scatter_rows_add", dout.contiguous(), idx, dx, idx.shape[0], shape[1])
        return dx, None, None


class WRAFn(Function):
    @staticmethod
    def forward(ctx, seq, phrase_index, img_index, neg_img, rand_pos, rand_neg, rt):
        seq = seq.contiguous()
        B, Lt, H = seq.shape
        dev = seq.device
        maxp = _lib.lib().mvptr_wra_max_phrases()
        out = torch.empty(2, B, device=dev, dtype=F32)
        sel = torch.full((2, B, maxp), -1, device=dev, dtype=torch.int32)
        _lib.call("mvptr_f32_wra_fwd", seq, B, Lt, H, phrase_index, img_index, neg_img, rand_pos, rand_neg,
                  rand_pos.shape[1], out[0], out[1], sel[0], sel[1])
        ctx.s = (seq, phrase_index, neg_img, sel)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, dpos, dneg):
        seq, phrase_index, neg_img, sel = ctx.s
        B, Lt, H = seq.shape
        dseq = torch.zeros(B, Lt, H, device=seq.device, dtype=F32)
        _lib.call("mvptr_f32_wra_bwd", seq, B, Lt, H, phrase_index, neg_img, sel[0], sel[1], dpos.contiguous(),
                  dneg.contiguous(), dseq)
        return dseq, None, None, None, None, None, None


class ClsRegionScoreFn(Function):
    @staticmethod
    def forward(ctx, *a):
        raise NotImplementedError("the referring-expression head (SURVEY f-3) is not part of the fp32 verification tier")
