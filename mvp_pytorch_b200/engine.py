"""Host-side orchestration of the CUDA hot path: torch.autograd.Functions whose forward and
backward are sequences of libmvptr_b200.so calls.  PyTorch supplies device memory, streams
and the autograd tape; all arithmetic on activations happens in the C-ABI kernels.

Weight gradients are never returned to autograd: the wgrad GEMMs reduce-add directly into
the model's flat fp32 gradient arena (arena.py), whose slices are the parameters' ``.grad``.
"""
import torch
from torch.autograd import Function

from . import _lib
from .arena import ParamArena

BF16 = torch.bfloat16
F32 = torch.float32
NUM_SMS = 148


def _pad8(n):
    return (n + 7) // 8 * 8


class Runtime:
    """Per-model state shared by all Functions of one forward/backward."""

    def __init__(self, model, config, precision="bf16"):
        self.model = model
        self.cfg = config
        self.arena = ParamArena(model)
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision {precision!r}: 'bf16' (the product path) or 'fp32' (verification tier)")
        # fp32 verification tier (engine_fp32.py): fp32 activations, fp32 master weights as operands,
        # 3-way-split tensor-core contractions
        self.fp32 = precision == "fp32"
        self.adt = F32 if self.fp32 else BF16   # activation storage dtype
        self.f32_cache = {}
        if self.fp32 and self.arena.dtype != F32:
            raise _lib.MvptrError("the fp32 verification tier needs float32 parameters")
        self.H = config.hidden_size
        self.I = config.intermediate_size
        self.nh = config.num_attention_heads
        if self.H != self.nh * 64:
            raise _lib.MvptrError(f"hidden_size {self.H} / heads {self.nh}: the attention kernel needs head_dim 64")
        self.eps = float(config.layer_norm_eps)
        self.seed_base = int(torch.initial_seed()) & 0x7FFFFFFF
        self.seed_ctr = 0
        self.img_w16 = None
        self._img_w_version = -1
        self._launch_base = _lib.launch_count()
        self.layer_protos = {}

    # ---- bookkeeping -------------------------------------------------------------------
    def next_seed(self):
        self.seed_ctr += 1
        return (self.seed_base * 2654435761 + self.seed_ctr * 974711) & 0xFFFFFFFF

    @property
    def launches(self):
        """Kernels launched by the library since this runtime was created (the library's own count)."""
        return _lib.launch_count() - self._launch_base

    def call(self, name, *args):
        _lib.call(name, *args)

    def gemm(self, *a, **kw):
        return _lib.gemm(*a, **kw)

    def anchor(self, param):
        """A fresh scalar leaf that ties this forward's Functions into the autograd graph in place of
        the nn.Parameter `param` (whose requires_grad it inherits).  Parameter gradients are written
        straight into the arena by the kernels, never by autograd, and a Parameter's AccumulateGrad
        node stays bound to the stream of its FIRST backward -- which makes a later CUDA-graph capture
        on another stream fail with a cross-stream dependency.  A per-forward leaf has no history."""
        return torch.zeros((), device=self.arena.device, dtype=F32,
                           requires_grad=bool(param.requires_grad) and torch.is_grad_enabled())

    def begin_forward(self, training):
        a = self.arena
        if not a.valid():
            raise _lib.MvptrError("parameters were moved or re-typed after the first forward; call "
                                  "model.rebuild_arena() after .to()/.half()/.bfloat16()")
        if self.fp32:
            self.f32_cache = {}  # operand splits of the weights: valid for this forward and its backward
            if training and (self.cfg.hidden_dropout_prob > 0 or self.cfg.attention_probs_dropout_prob > 0):
                raise _lib.MvptrError("the fp32 verification tier is deterministic: set the dropout probabilities to 0 "
                                      "(or call model.eval())")
            self.training = training
            return
        # optimizers that update through p.data do not bump version counters -> recast every training step
        a.refresh_shadow(force=training and a.shadow is not a.master and not getattr(self, "shadow_managed", False))
        self.training = training

    def img_weight(self, name):
        """bf16 copy of the region projection weight with the K=2054 pitch padded to 16 bytes."""
        a = self.arena
        m = a.master_of(name)
        K = m.shape[1]
        Kp = _pad8(K)
        if self.img_w16 is None:
            self.img_w16 = torch.zeros(m.shape[0], Kp, device=m.device, dtype=BF16)
        if self.training or self._img_w_version != a.master._version:
            self.call("mvptr_pad_cast", m, int(m.dtype == F32), K, self.img_w16, Kp, m.shape[0], K)
            self._img_w_version = a.master._version
        return self.img_w16


def split_k_for(m_out, n_out, k, bn=256):
    """Enough K-splits to put ~2 waves of CTAs on 148 SMs (wgrad outputs are small, K is the token count)."""
    tiles = ((m_out + 127) // 128) * ((n_out + bn - 1) // bn)
    kb = (k + 63) // 64
    s = max(1, (2 * NUM_SMS) // tiles)
    return max(1, min(s, kb // 4 if kb >= 4 else 1))


def wgrad(rt, dY, ld_dy, X, ld_x, n_out, k_in, tokens, dW, ldw=None):
    """dW[n_out, k_in] += dY[tokens, n_out]^T . X[tokens, k_in]   (fp32 reduce-add, split-K)"""
    rt.gemm(dY, X, dW, n_out, k_in, tokens, lda=ld_dy, ldb=ld_x, ldd=ldw or k_in, a_mn=True, b_mn=True,
            accumulate=True, split_k=split_k_for(n_out, k_in, tokens))


def dgrad(rt, dY, ld_dy, W, ld_w, tokens, n_out, k_in, dX, **epi):
    """dX[tokens, k_in] = dY[tokens, n_out] . W[n_out, k_in]"""
    rt.gemm(dY, W, dX, tokens, k_in, n_out, lda=ld_dy, ldb=ld_w, ldd=k_in, b_mn=True, **epi)


# ======================================================================================
# Encoder stack: CaptionBertEncoder.forward (modeling_vlbert.py:134-178) over
# CaptionBertLayer (:191-199) = attention (:63-103) + BertSelfOutput / BertIntermediate /
# BertOutput (modeling_bert.py:348-352, 394-397, 407-411)
# ======================================================================================
_W_FIELDS = (("w_qkv", "attention.self.query.weight"), ("b_qkv", "attention.self.query.bias"),
             ("w_o", "attention.output.dense.weight"), ("b_o", "attention.output.dense.bias"),
             ("ln1_g", "attention.output.LayerNorm.weight"), ("ln1_b", "attention.output.LayerNorm.bias"),
             ("w_i", "intermediate.dense.weight"), ("b_i", "intermediate.dense.bias"),
             ("w_o2", "output.dense.weight"), ("b_o2", "output.dense.bias"),
             ("ln2_g", "output.LayerNorm.weight"), ("ln2_b", "output.LayerNorm.bias"))


def _layer_proto(rt, pf, with_grads):
    """LayerArgs with the (stable) weight / gradient arena pointers of one layer filled in."""
    key = (pf, with_grads, id(rt.arena.shadow), id(rt.arena.grad))
    proto = rt.layer_protos.get(key)
    if proto is None:
        a = rt.arena
        offs = a.offsets
        # fused QKV relies on query/key/value tensors being adjacent in the arena
        for kind in ("weight", "bias"):
            q, k, v = (offs[pf + f"attention.self.{n}.{kind}"] for n in ("query", "key", "value"))
            assert k[0] == q[0] + q[1] and v[0] == k[0] + k[1], "q/k/v must be adjacent in the parameter arena"
        proto = _lib.LayerArgs()
        proto.H, proto.I, proto.nh, proto.eps = rt.H, rt.I, rt.nh, rt.eps
        sh = a.shadow.data_ptr()
        for f, n in _W_FIELDS:
            setattr(proto, f, sh + 2 * offs[pf + n][0])
        if with_grads:
            g = a.ensure_grad().data_ptr()
            for f, n in _W_FIELDS:
                setattr(proto, "g_" + f, g + 4 * offs[pf + n][0])
            a.touched.update(k for k in offs if k.startswith(pf))
        rt.layer_protos[key] = proto
    return _lib.LayerArgs.from_buffer_copy(proto)


class EncoderFn(Function):
    """One ctypes call per layer (csrc/layer.cu launches the layer's kernels from C++)."""

    @staticmethod
    def forward(ctx, h, maskadd, rt, prefix, layer_lo, layer_hi, save, anchor):
        B, L, H = h.shape
        M, I, nh = B * L, rt.I, rt.nh
        dev = h.device
        p_h = rt.cfg.hidden_dropout_prob if rt.training else 0.0
        p_a = rt.cfg.attention_probs_dropout_prob if rt.training else 0.0
        x = h.reshape(M, H)
        if not x.is_contiguous():
            x = x.contiguous()
        saved = []
        n_bf16 = M * (8 * H + 2 * I)  # qkv 3H | att | pre1 | a1 | pre2 | pre_g I | inter I | tmp
        n_f32 = B * nh * L + 4 * M
        ws = None
        for li in range(layer_lo, layer_hi):
            pf = f"{prefix}.layer.{li}."
            la = _layer_proto(rt, pf, False)
            if save or ws is None:
                blk = torch.empty(n_bf16, device=dev, dtype=BF16)
                f32 = torch.empty(n_f32, device=dev, dtype=F32)
                ws = (blk, f32)
            else:
                blk, f32 = ws
            out = torch.empty(M, H, device=dev, dtype=BF16)
            b0, f0 = blk.data_ptr(), f32.data_ptr()
            la.B, la.L, la.save = B, L, int(save)
            la.p_hidden, la.p_attn = p_h, p_a
            la.seed_attn, la.seed1, la.seed2 = rt.next_seed(), rt.next_seed(), rt.next_seed()
            la.maskadd = maskadd.data_ptr()
            la.x = x.data_ptr()
            la.qkv = b0
            la.att = b0 + 2 * M * 3 * H
            la.pre1 = b0 + 2 * M * 4 * H
            la.a1 = b0 + 2 * M * 5 * H
            la.pre2 = b0 + 2 * M * 6 * H
            la.pre_g = b0 + 2 * M * 7 * H
            la.inter = b0 + 2 * M * (7 * H + I)
            la.tmp = b0 + 2 * M * (7 * H + 2 * I)
            la.out = out.data_ptr()
            la.lse = f0
            la.st1 = f0 + 4 * B * nh * L
            la.st2 = f0 + 4 * (B * nh * L + 2 * M)
            _lib.layer_call("mvptr_layer_fwd", la, 7)
            if save:
                saved.append((pf, la, x, blk, f32, out))
            x = out
        ctx.rt, ctx.saved, ctx.dims, ctx.maskadd = rt, saved, (B, L, H), maskadd
        return x.view(B, L, H)

    @staticmethod
    def backward(ctx, dout):
        rt, (B, L, H), maskadd = ctx.rt, ctx.dims, ctx.maskadd
        M, I = B * L, rt.I
        dev = dout.device
        dy = dout.reshape(M, H)
        if not dy.is_contiguous():
            dy = dy.contiguous()
        # scratch shared by all layers: dpre2 | dpre2d | da1 | dpre1 | dpre1d | datt | dqkv 3H | dpre_g I
        scratch = torch.empty(M * (9 * H + I), device=dev, dtype=BF16)
        s0 = scratch.data_ptr()
        for (pf, la_f, x, blk, f32, out) in reversed(ctx.saved):
            la = _layer_proto(rt, pf, True)
            for f in ("B", "L", "save", "p_hidden", "p_attn", "seed_attn", "seed1", "seed2", "maskadd", "x", "qkv",
                      "att", "pre1", "a1", "pre2", "pre_g", "inter", "out", "lse", "st1", "st2"):
                setattr(la, f, getattr(la_f, f))
            dx = torch.empty(M, H, device=dev, dtype=BF16)
            la.dout, la.dx = dy.data_ptr(), dx.data_ptr()
            la.dpre2 = s0
            la.dpre2d = s0 + 2 * M * H
            la.da1 = s0 + 2 * M * 2 * H
            la.dpre1 = s0 + 2 * M * 3 * H
            la.dpre1d = s0 + 2 * M * 4 * H
            la.datt = s0 + 2 * M * 5 * H
            la.dqkv = s0 + 2 * M * 6 * H
            la.dpre_g = s0 + 2 * M * 9 * H
            sync = getattr(rt, "grad_sync", None)
            if sync is not None and sync.enabled:
                offs = rt.arena.offsets
                w0 = offs[pf + "attention.self.query.weight"][0]
                w1 = offs[pf + "output.dense.weight"]
                b0 = offs[pf + "attention.self.query.bias"][0]
                b1 = offs[pf + "output.LayerNorm.bias"]
                sync.will_write(w0, w1[0] + w1[1])
                sync.will_write(b0, b1[0] + b1[1])
            _lib.layer_call("mvptr_layer_bwd", la, 12)
            if sync is not None and sync.enabled:  # this layer's gradients are final: reduce them now
                sync.layer_done(w0, w1[0] + w1[1])
                sync.layer_done(b0, b1[0] + b1[1])
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
class EmbedFn(Function):
    """BertEmbeddings.forward, modeling_bert.py:262-277."""

    @staticmethod
    def forward(ctx, ids, type_ids, pos_ids, rt, prefix, save, anchor):
        B, L = ids.shape
        H, a, cfg = rt.H, rt.arena, rt.cfg
        dev = ids.device
        p = cfg.hidden_dropout_prob if rt.training else 0.0
        seed = rt.next_seed()
        y = torch.empty(B, L, H, device=dev, dtype=BF16)
        pre = torch.empty(B * L, H, device=dev, dtype=BF16) if save else None
        st = torch.empty(2, B * L, device=dev, dtype=F32) if save else None
        rt.call("mvptr_embed_ln_fwd", ids, type_ids, pos_ids, a.w(prefix + ".word_embeddings.weight"),
                a.w(prefix + ".position_embeddings.weight"), a.w(prefix + ".token_type_embeddings.weight"),
                a.w(prefix + ".LayerNorm.weight"), a.w(prefix + ".LayerNorm.bias"), y, 0, 0, pre,
                st[0] if save else None, st[1] if save else None, B, L, H, rt.eps, cfg.vocab_size,
                cfg.max_position_embeddings, cfg.type_vocab_size, p, seed)
        ctx.rt, ctx.s = rt, (prefix, ids, type_ids, pos_ids, pre, st, p, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        prefix, ids, type_ids, pos_ids, pre, st, p, seed = ctx.s
        B, L = ids.shape
        embed_backward(rt, prefix, dy.contiguous(), 0, 0, ids, type_ids, pos_ids, pre, st, p, seed, B, L)
        return (None,) * 7


def embed_backward(rt, prefix, dy, rows_per_batch, batch_stride, ids, type_ids, pos_ids, pre, st, p, seed, B, L):
    a, H, cfg = rt.arena, rt.H, rt.cfg
    if pos_ids is not None:
        raise NotImplementedError("backward through explicit position_ids is not supported")
    dpre = torch.empty(B * L, H, device=dy.device, dtype=BF16)
    rt.call("mvptr_ln_bwd", dy, rows_per_batch, batch_stride, pre, st[0], st[1], a.w(prefix + ".LayerNorm.weight"),
            dpre, None, a.g(prefix + ".LayerNorm.weight"), a.g(prefix + ".LayerNorm.bias"), None, B * L, H, p, seed,
            0.0, 0)
    rt.call("mvptr_embed_bwd", dpre, ids, type_ids, a.g(prefix + ".word_embeddings.weight"),
            a.g(prefix + ".position_embeddings.weight"), a.g(prefix + ".token_type_embeddings.weight"), B, L, H,
            cfg.vocab_size, cfg.type_vocab_size, 0)


class VisInputFn(Function):
    """Tag embeddings + region projection + LN + dropout, written into one [B, Lt+R, H]
    buffer (modeling_vlbert.py:481-482, 498-506: embeddings(b), img_embedding, LayerNorm,
    dropout, torch.cat)."""

    @staticmethod
    def forward(ctx, ids_b, type_b, pos_b, img_feats, rt, bert, save, anchor):
        B, Lt = ids_b.shape
        R, Kimg = img_feats.shape[1], img_feats.shape[2]
        H, a, cfg = rt.H, rt.arena, rt.cfg
        dev = ids_b.device
        Lv = Lt + R
        p = cfg.hidden_dropout_prob if rt.training else 0.0
        s_tag, s_img = rt.next_seed(), rt.next_seed()
        out = torch.empty(B, Lv, H, device=dev, dtype=BF16)
        emb = bert + "embeddings"
        pre_t = torch.empty(B * Lt, H, device=dev, dtype=BF16) if save else None
        st_t = torch.empty(2, B * Lt, device=dev, dtype=F32) if save else None
        rt.call("mvptr_embed_ln_fwd", ids_b, type_b, pos_b, a.w(emb + ".word_embeddings.weight"),
                a.w(emb + ".position_embeddings.weight"), a.w(emb + ".token_type_embeddings.weight"),
                a.w(emb + ".LayerNorm.weight"), a.w(emb + ".LayerNorm.bias"), out, Lt, Lv * H, pre_t,
                st_t[0] if save else None, st_t[1] if save else None, B, Lt, H, rt.eps, cfg.vocab_size,
                cfg.max_position_embeddings, cfg.type_vocab_size, p, s_tag)
        # region features -> bf16 with a 16-byte aligned pitch, then the K=2054 projection
        Kp = _pad8(Kimg)
        kind = {BF16: 0, F32: 1, torch.float16: 2}.get(img_feats.dtype)
        if kind is None:
            raise _lib.MvptrError(f"img_feats dtype {img_feats.dtype} unsupported (float32, bfloat16 or float16)")
        feats = img_feats if img_feats.is_contiguous() else img_feats.contiguous()
        x16 = torch.empty(B * R, Kp, device=dev, dtype=BF16)
        rt.call("mvptr_pad_cast", feats, kind, Kimg, x16, Kp, B * R, Kimg)
        w16 = rt.img_weight(bert + "img_embedding.weight")
        pre_i = torch.empty(B * R, H, device=dev, dtype=BF16)
        rt.gemm(x16, w16, pre_i, B * R, H, Kimg, lda=Kp, ldb=Kp, ldd=H, bias=a.w(bert + "img_embedding.bias"))
        st_i = torch.empty(2, B * R, device=dev, dtype=F32) if save else None
        img_rows = out.view(B * Lv, H)[Lt:]  # row (b, r) lives at b*Lv*H + (Lt + r)*H
        if cfg.use_img_layernorm:
            rt.call("mvptr_ln_fwd", pre_i, a.w(bert + "LayerNorm.weight"), a.w(bert + "LayerNorm.bias"), img_rows, R,
                    Lv * H, st_i[0] if save else None, st_i[1] if save else None, B * R, H,
                    float(cfg.img_layer_norm_eps), p, s_img)
        else:
            raise NotImplementedError("use_img_layernorm=0 is not supported by the CUDA path")
        ctx.rt = rt
        ctx.s = (bert, ids_b, type_b, pos_b, pre_t, st_t, x16, pre_i, st_i, p, s_tag, s_img, (B, Lt, R, Kimg, Kp))
        return out

    @staticmethod
    def backward(ctx, dout):
        rt = ctx.rt
        bert, ids_b, type_b, pos_b, pre_t, st_t, x16, pre_i, st_i, p, s_tag, s_img, (B, Lt, R, Kimg, Kp) = ctx.s
        a, H = rt.arena, rt.H
        Lv = Lt + R
        dout = dout.contiguous()
        embed_backward(rt, bert + "embeddings", dout, Lt, Lv * H, ids_b, type_b, pos_b, pre_t, st_t, p, s_tag, B, Lt)
        d_img_rows = dout.view(B * Lv, H)[Lt:]
        dpre = torch.empty(B * R, H, device=dout.device, dtype=BF16)
        rt.call("mvptr_ln_bwd", d_img_rows, R, Lv * H, pre_i, st_i[0], st_i[1], a.w(bert + "LayerNorm.weight"), dpre,
                None, a.g(bert + "LayerNorm.weight"), a.g(bert + "LayerNorm.bias"), a.g(bert + "img_embedding.bias"),
                B * R, H, p, s_img, 0.0, 0)
        # dW[H, 2054]: fp32 pitch 2054*4 B is not 16-byte aligned -> padded scratch, then add
        scratch = torch.zeros(H, Kp, device=dout.device, dtype=F32)
        wgrad(rt, dpre, H, x16, Kp, H, Kp, B * R, scratch, ldw=Kp)
        a.g(bert + "img_embedding.weight").add_(scratch[:, :Kimg])
        return (None,) * 8


# ======================================================================================
# Small dense heads on the [CLS] rows
# ======================================================================================
class ClsProjNormFn(Function):
    """normalize(seq[:,0] @ proj): modeling_vlbert.py:525-526 / :717-718 -> fp32 [B,H]."""

    @staticmethod
    def forward(ctx, seq, rt, proj_name, anchor):
        B, L, H = seq.shape
        a = rt.arena
        x32 = torch.empty(B, H, device=seq.device, dtype=F32)
        # x @ P : B operand is P[k, n] with n contiguous -> MN-major
        rt.gemm(seq, a.w(proj_name), x32, B, H, H, lda=L * H, ldb=H, ldd=H, b_mn=True)
        y32 = torch.empty(B, H, device=seq.device, dtype=F32)
        norm = torch.empty(B, device=seq.device, dtype=F32)
        rt.call("mvptr_l2norm_fwd", x32, y32, None, norm, B, H)
        ctx.rt, ctx.s = rt, (seq, proj_name, y32, norm)
        return y32

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        seq, proj_name, y32, norm = ctx.s
        B, L, H = seq.shape
        a = rt.arena
        dx16 = torch.empty(B, H, device=dy.device, dtype=BF16)
        rt.call("mvptr_l2norm_bwd", dy.contiguous(), y32, norm, dx16, B, H)
        # dP[k, n] += sum_b x[b,k] dx[b,n]  ->  "out rows" = k (x is the A operand)
        rt.gemm(seq, dx16, a.g(proj_name), H, H, B, lda=L * H, ldb=H, ldd=H, a_mn=True, b_mn=True, accumulate=True)
        dcls = torch.empty(B, H, device=dy.device, dtype=BF16)
        # dx_cls[b, k] = sum_n dx[b,n] P[k,n]  -> B operand P is K-major here
        rt.gemm(dx16, a.w(proj_name), dcls, B, H, H, lda=H, ldb=H, ldd=H)
        dseq = torch.zeros_like(seq)
        dseq[:, 0] = dcls
        return dseq, None, None, None


class ClsDenseFn(Function):
    """act(seq[:, 0] W^T + b) on the first token of every sequence: BertPooler
    (modeling_bert.py:468-474, act=tanh) without materialising the gathered rows."""

    @staticmethod
    def forward(ctx, seq, rt, wname, bname, act, anchor):
        B, L, H = seq.shape
        a = rt.arena
        N = a.w(wname).shape[0]
        y = torch.empty(B, N, device=seq.device, dtype=BF16)
        rt.gemm(seq, a.w(wname), y, B, N, H, lda=L * H, ldb=H, ldd=N, bias=a.w(bname), act=act)
        ctx.rt, ctx.s = rt, (seq, wname, bname, act, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        seq, wname, bname, act, y = ctx.s
        B, L, H = seq.shape
        a = rt.arena
        N = y.shape[1]
        if act == "tanh":
            yf = y.float()
            dpre = (dy.float() * (1.0 - yf * yf)).to(BF16)
        elif act is None:
            dpre = dy.contiguous()
        else:
            raise NotImplementedError(act)
        rt.call("mvptr_colsum", dpre, N, a.g(bname), B, N)
        rt.gemm(dpre, seq, a.g(wname), N, H, B, lda=N, ldb=L * H, ldd=H, a_mn=True, b_mn=True, accumulate=True)
        dcls = torch.empty(B, H, device=dy.device, dtype=BF16)
        dgrad(rt, dpre, N, a.w(wname), H, B, N, H, dcls)
        dseq = torch.zeros_like(seq)
        dseq[:, 0] = dcls
        return dseq, None, None, None, None, None


class HeadTransformFn(Function):
    """BertPredictionHeadTransform: LN(gelu(x W^T + b)), modeling_bert.py:487-491."""

    @staticmethod
    def forward(ctx, x, rt, prefix, anchor):
        n, H = x.shape
        a = rt.arena
        dev = x.device
        x = x.contiguous()
        t = torch.empty(n, H, device=dev, dtype=BF16)
        pre = torch.empty(n, H, device=dev, dtype=BF16)
        rt.gemm(x, a.w(prefix + ".dense.weight"), t, n, H, H, lda=H, ldb=H, ldd=H, bias=a.w(prefix + ".dense.bias"),
                act="gelu", pre_act=pre, ld_aux=H)
        y = torch.empty(n, H, device=dev, dtype=BF16)
        st = torch.empty(2, n, device=dev, dtype=F32)
        rt.call("mvptr_ln_fwd", t, a.w(prefix + ".LayerNorm.weight"), a.w(prefix + ".LayerNorm.bias"), y, 0, 0, st[0],
                st[1], n, H, rt.eps, 0.0, 0)
        ctx.rt, ctx.s = rt, (x, prefix, t, pre, st)
        return y

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        x, prefix, t, pre, st = ctx.s
        n, H = x.shape
        a = rt.arena
        dev = dy.device
        dt = torch.empty(n, H, device=dev, dtype=BF16)
        rt.call("mvptr_ln_bwd", dy.contiguous(), 0, 0, t, st[0], st[1], a.w(prefix + ".LayerNorm.weight"), dt, None,
                a.g(prefix + ".LayerNorm.weight"), a.g(prefix + ".LayerNorm.bias"), None, n, H, 0.0, 0, 0.0, 0)
        dpre = torch.empty(n, H, device=dev, dtype=BF16)
        rt.call("mvptr_gelu_bwd", dt, pre, dpre, n * H)
        rt.call("mvptr_colsum", dpre, H, a.g(prefix + ".dense.bias"), n, H)
        wgrad(rt, dpre, H, x, H, H, H, n, a.g(prefix + ".dense.weight"))
        dx = torch.empty(n, H, device=dev, dtype=BF16)
        dgrad(rt, dpre, H, a.w(prefix + ".dense.weight"), H, n, H, H, dx)
        return dx, None, None, None


def decoder_logits(rt, t, wname, n_out, bias_name):
    """fp32 logits [n, pad8(n_out)] = t W[:n_out]^T + bias  (modeling_bert.py:514 / :531)."""
    n, H = t.shape
    a = rt.arena
    pitch = _pad8(n_out)
    # row count rounded up so that the data-dependent number of masked positions maps onto a few
    # recurring allocation sizes (a fresh cudaMalloc inside a step costs tens of ms)
    logits = torch.empty((n + 255) // 256 * 256, pitch, device=t.device, dtype=F32)[:n]
    rt.gemm(t, a.w(wname), logits, n, n_out, H, lda=H, ldb=H, ldd=pitch, bias=a.w_span(bias_name, n_out))
    return logits


def decoder_backward(rt, t, dlogits, wname, n_out, bias_name):
    n, H = t.shape
    a = rt.arena
    pitch = dlogits.shape[1]
    # bias slot in the arena is padded to 8 elements, so the padded columns (zeros) are harmless
    rt.call("mvptr_colsum", dlogits, pitch, a.g_span(bias_name, pitch), n, pitch)
    wgrad(rt, dlogits, pitch, t, H, n_out, H, n, a.g(wname))
    # dt[n, H] = dlogits[n, n_out] . W[n_out, H]: a few thousand rows x 768 columns is only ~54 output tiles
    # for a 30 522-deep contraction (18-36 CTAs busy for 155 us in the ncu launch list) -> split K over the
    # idle SMs, fp32 TMA reduce-add, one cast
    split = split_k_for(n, H, n_out)
    if split > 1:
        dt32 = torch.zeros(n, H, device=t.device, dtype=F32)
        dgrad(rt, dlogits, pitch, a.w(wname), H, n, n_out, H, dt32, accumulate=True, split_k=split)
        dt = torch.empty(n, H, device=t.device, dtype=BF16)
        rt.call("mvptr_cast_f32_bf16", dt32, dt, n * H)
        return dt
    dt = torch.empty(n, H, device=t.device, dtype=BF16)
    dgrad(rt, dlogits, pitch, a.w(wname), H, n, n_out, H, dt)
    return dt


class VocabCEFn(Function):
    """decoder (tied to word_embeddings[:only_word_size]) + bias + CrossEntropy(ignore_index=-1):
    modeling_bert.py:513-516 with modeling_vlbert.py:1212-1216 (tie), :1235 / :1249 (loss)."""

    @staticmethod
    def forward(ctx, t, labels, rt, wname, n_out, bias_name, anchor):
        t = t.contiguous()
        n = t.shape[0]
        logits = decoder_logits(rt, t, wname, n_out, bias_name)
        lse = torch.empty(n, device=t.device, dtype=F32)
        acc = torch.zeros(2, device=t.device, dtype=F32)
        rt.call("mvptr_ce_fwd", logits, logits.shape[1], labels, n, n_out, -1, lse, acc[0:1], acc[1:2])
        ctx.rt, ctx.s = rt, (t, labels, wname, n_out, bias_name, logits, lse, acc)
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, g):
        rt = ctx.rt
        t, labels, wname, n_out, bias_name, logits, lse, acc = ctx.s
        n = t.shape[0]
        pitch = logits.shape[1]
        dlogits = torch.empty((n + 255) // 256 * 256, pitch, device=t.device, dtype=BF16)[:n]
        gs = g.reshape(1).to(F32).contiguous()
        rt.call("mvptr_ce_bwd", logits, pitch, labels, n, n_out, -1, lse, acc[1:2], gs, dlogits, pitch)
        dt = decoder_backward(rt, t, dlogits, wname, n_out, bias_name)
        ctx.s = None
        return dt, None, None, None, None, None, None


class DecoderFn(Function):
    """Plain decoder logits (VQA answers / MLM inference): modeling_bert.py:514, :531."""

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
        d16 = torch.zeros(t.shape[0], pitch, device=t.device, dtype=BF16)
        d16[:, :n_out] = dl
        return decoder_backward(rt, t, d16, wname, n_out, bias_name), None, None, None, None, None


class BCEFn(Function):
    """instance_bce_with_logits (modeling_vlbert.py:878-883) fused with the answer decoder."""

    @staticmethod
    def forward(ctx, t, labels, rt, wname, n_out, bias_name, anchor):
        t = t.contiguous()
        n = t.shape[0]
        logits = decoder_logits(rt, t, wname, n_out, bias_name)
        lab = labels.to(F32).contiguous()
        loss = torch.zeros(1, device=t.device, dtype=F32)
        rt.call("mvptr_bce_fwd", logits, logits.shape[1], lab, n, n_out, loss)
        ctx.rt, ctx.s = rt, (t, lab, wname, n_out, bias_name, logits)
        ctx.mark_non_differentiable(logits)
        return loss[0], logits

    @staticmethod
    def backward(ctx, g, _):
        rt = ctx.rt
        t, lab, wname, n_out, bias_name, logits = ctx.s
        n, pitch = logits.shape
        dlogits = torch.empty(n, pitch, device=t.device, dtype=BF16)
        rt.call("mvptr_bce_bwd", logits, pitch, lab, n, n_out, g.reshape(1).to(F32).contiguous(), dlogits, pitch)
        return decoder_backward(rt, t, dlogits, wname, n_out, bias_name), None, None, None, None, None, None


class LinearFn(Function):
    """nn.Linear on an arbitrary [n, K] activation: y = act(x W^T + b), act in {None, 'relu', 'tanh'}; or, with
    kn=True, y = x P for a projection stored [K, N] (txt_proj / vis_proj, modeling_vlbert.py:525-526).  The thin
    task heads (classifier MLPs :1730-1744, single_mapping :1991-1995) are built from it; N and K must be
    multiples of 8 (narrower outputs go through SmallHeadFn)."""

    @staticmethod
    def forward(ctx, x, rt, wname, bname, act, kn, anchor):
        x = x.to(BF16).contiguous()
        n, K = x.shape
        a = rt.arena
        W = a.w(wname)
        N = W.shape[1] if kn else W.shape[0]
        y = torch.empty(n, N, device=x.device, dtype=BF16)
        bias = a.w(bname) if bname is not None else None
        if kn:
            rt.gemm(x, W, y, n, N, K, lda=K, ldb=N, ldd=N, b_mn=True, bias=bias)
        else:
            rt.gemm(x, W, y, n, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="tanh" if act == "tanh" else None)
        if act == "relu":
            y = torch.relu_(y)
        ctx.rt, ctx.s = rt, (x, wname, bname, act, kn, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        rt = ctx.rt
        x, wname, bname, act, kn, y = ctx.s
        n, K = x.shape
        N = y.shape[1]
        a = rt.arena
        W = a.w(wname)
        if act == "relu":
            dpre = (dy * (y > 0)).to(BF16).contiguous()
        elif act == "tanh":
            yf = y.float()
            dpre = (dy.float() * (1.0 - yf * yf)).to(BF16)
        else:
            dpre = dy.to(BF16).contiguous()
        if bname is not None:
            rt.call("mvptr_colsum", dpre, N, a.g(bname), n, N)
        dx = torch.empty(n, K, device=dy.device, dtype=BF16)
        if kn:
            # dP[k, n'] += sum_r x[r, k] dpre[r, n'] ; dx[r, k] = sum_n' dpre[r, n'] P[k, n']
            rt.gemm(x, dpre, a.g(wname), K, N, n, lda=K, ldb=N, ldd=N, a_mn=True, b_mn=True, accumulate=True)
            rt.gemm(dpre, W, dx, n, K, N, lda=N, ldb=N, ldd=K)
        else:
            wgrad(rt, dpre, N, x, K, N, K, n, a.g(wname))
            dgrad(rt, dpre, N, W, K, n, N, K, dx)
        return dx, None, None, None, None, None, None


class ClsRegionScoreFn(Function):
    """logits[b, j] = cos / dot of region token (first + j) with the [CLS] token, after an optional dropout of the
    sequence output: BiImageBertForRE mod 1 / mod 2 (modeling_vlbert.py:1931-1951) -> fp32 [B, R]."""

    @staticmethod
    def forward(ctx, seq, first, normalize, p_drop, rt):
        seq = seq.contiguous()
        B, Ltot, H = seq.shape
        R = Ltot - first
        logits = torch.empty(B, R, device=seq.device, dtype=F32)
        inv = torch.empty(B * R, 2, device=seq.device, dtype=F32)
        seed = rt.next_seed() if p_drop > 0 else 0
        rt.call("mvptr_cls_region_score_fwd", seq, B, Ltot, H, first, R, int(normalize), logits, inv, float(p_drop), seed)
        ctx.rt, ctx.s = rt, (seq, first, normalize, p_drop, seed, logits, inv)
        return logits

    @staticmethod
    def backward(ctx, dl):
        rt = ctx.rt
        seq, first, normalize, p_drop, seed, logits, inv = ctx.s
        B, Ltot, H = seq.shape
        dseq = torch.zeros_like(seq)
        rt.call("mvptr_cls_region_score_bwd", seq, B, Ltot, H, first, Ltot - first, int(normalize), logits, inv,
                dl.to(F32).contiguous(), dseq, float(p_drop), seed)
        return dseq, None, None, None, None


class SmallHeadFn(Function):
    """x W^T + b with a handful of outputs (ITM / retrieval classifier, modeling_vlbert.py:1247,
    :1680, :1708) -> fp32 logits."""

    @staticmethod
    def forward(ctx, x, rt, wname, bname, anchor):
        x = x.contiguous()
        n, H = x.shape
        a = rt.arena
        C = a.w(wname).shape[0]
        logits = torch.empty(n, C, device=x.device, dtype=F32)
        rt.call("mvptr_small_head_fwd", x, H, a.w(wname), a.w(bname), logits, n, H, C)
        ctx.rt, ctx.s = rt, (x, wname, bname, C)
        return logits

    @staticmethod
    def backward(ctx, dl):
        rt = ctx.rt
        x, wname, bname, C = ctx.s
        n, H = x.shape
        a = rt.arena
        dx = torch.empty(n, H, device=x.device, dtype=BF16)
        rt.call("mvptr_small_head_bwd", dl.to(F32).contiguous(), x, H, a.w(wname), dx, H, a.g(wname), a.g(bname), n, H, C)
        return dx, None, None, None, None


class SmallCEFn(Function):
    """CrossEntropyLoss(ignore_index=-1) over [n, C] fp32 logits, C small (modeling_vlbert.py:1251, :1262-1264,
    :1682): labels outside [0, C) are ignored, the mean runs over the valid rows."""

    @staticmethod
    def forward(ctx, logits, labels, rt):
        n, C = logits.shape
        logits = logits.contiguous()
        acc = torch.zeros(2, device=logits.device, dtype=F32)  # {sum of row losses, valid rows}
        rt.call("mvptr_small_ce", logits, labels, n, C, acc, None, None)
        ctx.rt, ctx.s = rt, (logits, labels, acc)
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, g):
        rt = ctx.rt
        logits, labels, acc = ctx.s
        n, C = logits.shape
        dl = torch.empty_like(logits)
        rt.call("mvptr_small_ce", logits, labels, n, C, acc, dl, g.reshape(1).to(F32).contiguous())
        return dl, None, None


# ======================================================================================
# Contrastive similarity, VSC loss, hard negatives
# ======================================================================================
def _split_bf16(x32):
    hi = x32.to(BF16)
    lo = (x32 - hi.float()).to(BF16)
    return hi, lo


def sim_matrix(rt, a32, b32):
    """fp32-accurate a . b^T from bf16 tensor-core passes (hi*hi + hi*lo + lo*hi): keeps the
    ranking / arg-max decisions taken on it (modeling_vlbert.py:527-534, run_retrieval.py:739)
    stable against bf16 operand rounding."""
    n, H = a32.shape
    m = b32.shape[0]
    pitch = _pad8(m)
    out = torch.empty(n, pitch, device=a32.device, dtype=F32)
    ah, al = _split_bf16(a32)
    bh, bl = _split_bf16(b32)
    rt.gemm(ah, bh, out, n, m, H, lda=H, ldb=H, ldd=pitch)
    rt.gemm(ah, bl, out, n, m, H, lda=H, ldb=H, ldd=pitch, accumulate=True)
    rt.gemm(al, bh, out, n, m, H, lda=H, ldb=H, ldd=pitch, accumulate=True)
    return out


class SimFn(Function):
    @staticmethod
    def forward(ctx, gt, gi, rt):
        sim = sim_matrix(rt, gt, gi)
        ctx.rt, ctx.s = rt, (gt, gi)
        return sim[:, : gi.shape[0]]

    @staticmethod
    def backward(ctx, dsim):
        rt = ctx.rt
        gt, gi = ctx.s
        n, H = gt.shape
        m = gi.shape[0]
        pn, pm = _pad8(n), _pad8(m)
        d16 = torch.zeros(n, pm, device=gt.device, dtype=BF16)
        d16[:, :m] = dsim
        gt16, gi16 = gt.to(BF16), gi.to(BF16)
        dgt = torch.empty(n, H, device=gt.device, dtype=F32)
        dgi = torch.empty(m, H, device=gt.device, dtype=F32)
        rt.gemm(d16, gi16, dgt, n, H, m, lda=pm, ldb=H, ldd=H, b_mn=True)
        rt.gemm(d16, gt16, dgi, m, H, n, lda=pm, ldb=H, ldd=H, a_mn=True, b_mn=True)
        return dgt, dgi, None


class VSCFn(Function):
    """(CE(s*sim, arange) + CE(s*sim^T, arange))/2 with s = exp(logit_scale)
    (modeling_vlbert.py:1238-1241) + in-batch hardest negatives (:530-534)."""

    @staticmethod
    def forward(ctx, sim, rt, ls_name, anchor):
        B = sim.shape[0]
        simc = sim.contiguous()
        dev = sim.device
        lse = torch.empty(2, B, device=dev, dtype=F32)
        loss = torch.zeros(1, device=dev, dtype=F32)
        hard = torch.empty(2, B, device=dev, dtype=torch.int64)
        ls = rt.arena.master_of(ls_name)
        if ls.dtype != F32:
            ls = ls.float()
        rt.call("mvptr_vsc_fwd", simc, B, ls, lse[0], lse[1], loss, hard[0], hard[1])
        ctx.rt, ctx.s = rt, (simc, ls, lse, ls_name)
        ctx.mark_non_differentiable(hard)
        return loss[0], hard

    @staticmethod
    def backward(ctx, g, _):
        rt = ctx.rt
        simc, ls, lse, ls_name = ctx.s
        B = simc.shape[0]
        dsim = torch.empty_like(simc)
        rt.call("mvptr_vsc_bwd", simc, B, ls, lse[0], lse[1], g.reshape(1).to(F32).contiguous(), dsim,
                rt.arena.g(ls_name).reshape(1))
        return dsim, None, None, None


def hard_negatives(rt, sim):
    """arg-max only (no loss): hn_mod='hard' of modeling_vlbert.py:530-534."""
    B = sim.shape[0]
    dev = sim.device
    lse = torch.empty(2, B, device=dev, dtype=F32)
    loss = torch.zeros(1, device=dev, dtype=F32)
    hard = torch.empty(2, B, device=dev, dtype=torch.int64)
    zero = torch.zeros(1, device=dev, dtype=F32)
    rt.call("mvptr_vsc_fwd", sim.contiguous(), B, zero, lse[0], lse[1], loss, hard[0], hard[1])
    return hard[0], hard[1]


# ======================================================================================
# Gathers
# ======================================================================================
class ConcatRowsFn(Function):
    """out[r] = cat(a[row_a[r]], b[row_b[r], col0:]): the torch.cat / index_select assembly of the
    joint and hard-negative stage-2 inputs, modeling_vlbert.py:542-566 and :586."""

    @staticmethod
    def forward(ctx, a3, b3, col0, row_a, row_b, rt):
        a3, b3 = a3.contiguous(), b3.contiguous()
        Ba, La, H = a3.shape
        Lb = b3.shape[1]
        rows = row_a.shape[0] if row_a is not None else Ba
        out = torch.empty(rows, La + Lb - col0, H, device=a3.device, dtype=BF16)
        rt.call("mvptr_concat_rows", a3, La, b3, Lb, col0, row_a, row_b, out, rows, H)
        ctx.rt, ctx.s = rt, (a3.shape, b3.shape, col0, row_a, row_b, rows)
        return out

    @staticmethod
    def backward(ctx, dout):
        rt = ctx.rt
        sa, sb, col0, row_a, row_b, rows = ctx.s
        dev = dout.device
        da = torch.zeros(sa, device=dev, dtype=F32)
        db = torch.zeros(sb, device=dev, dtype=F32)
        rt.call("mvptr_concat_rows_bwd", dout.contiguous(), sa[1], sb[1], col0, row_a, row_b, da, db, rows, sa[2])
        da16 = torch.empty(sa, device=dev, dtype=BF16)
        db16 = torch.empty(sb, device=dev, dtype=BF16)
        rt.call("mvptr_add_cast", da, None, da16, da.numel())
        rt.call("mvptr_add_cast", db, None, db16, db.numel())
        return da16, db16, None, None, None, None


class GatherRowsFn(Function):
    """torch.masked_select(...).reshape(-1, H) of the masked-LM rows (modeling_vlbert.py:1232, :1246)."""

    @staticmethod
    def forward(ctx, x2d, idx, rt):
        n, H = idx.shape[0], x2d.shape[1]
        out = torch.empty(n, H, device=x2d.device, dtype=BF16)
        rt.call("mvptr_gather_rows", x2d, idx, out, n, H)
        ctx.s = (x2d.shape, idx)
        return out

    @staticmethod
    def backward(ctx, dout):
        shape, idx = ctx.s
        dx = torch.zeros(shape, device=dout.device, dtype=BF16)
        # real indexes are unique; capacity-mode padding slots repeat row 0 with zero gradient, so add
        dx.index_add_(0, idx, dout)
        return dx, None, None


class WRAFn(Function):
    """Batched weakly-supervised phrase grounding: mean top-3-sampled cosine similarity of every
    phrase against the regions of its own image and of one sampled other image
    (modeling_vlbert.py:1288-1300, helpers :1502-1508, :1543-1596)."""

    @staticmethod
    def forward(ctx, seq, phrase_index, img_index, neg_img, rand_pos, rand_neg, rt):
        seq = seq.contiguous()
        B, Lt, H = seq.shape
        dev = seq.device
        maxp = _lib.lib().mvptr_wra_max_phrases()
        P = rand_pos.shape[1]
        out = torch.empty(2, B, device=dev, dtype=F32)
        sel = torch.full((2, B, maxp), -1, device=dev, dtype=torch.int32)
        rt.call("mvptr_wra_fwd", seq, B, Lt, H, phrase_index, img_index, neg_img, rand_pos, rand_neg, P, out[0],
                out[1], sel[0], sel[1])
        ctx.rt, ctx.s = rt, (seq, phrase_index, neg_img, sel)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, dpos, dneg):
        rt = ctx.rt
        seq, phrase_index, neg_img, sel = ctx.s
        B, Lt, H = seq.shape
        dseq = torch.zeros(B, Lt, H, device=seq.device, dtype=F32)
        rt.call("mvptr_wra_bwd", seq, B, Lt, H, phrase_index, neg_img, sel[0], sel[1], dpos.contiguous(),
                dneg.contiguous(), dseq)
        d16 = torch.empty(B, Lt, H, device=seq.device, dtype=BF16)
        rt.call("mvptr_add_cast", dseq, None, d16, dseq.numel())
        return d16, None, None, None, None, None, None


def mask_additive(rt, mask_a, mask_b=None, col0=0, row_a=None, row_b=None):
    """(1 - mask) * -10000 for [a | b[:, col0:]] rows (modeling_vlbert.py:430-460, 542-566, 587)."""
    La = mask_a.shape[1]
    Lb = mask_b.shape[1] if mask_b is not None else 0
    rows = row_a.shape[0] if row_a is not None else mask_a.shape[0]
    out = torch.empty(rows, La + Lb - col0, device=mask_a.device, dtype=F32)
    rt.call("mvptr_mask_prepare", mask_a, La, mask_b, Lb, col0, row_a, row_b, out, rows)
    return out


# ======================================================================================
# Precision dispatch: with Runtime.fp32 (model.set_precision("fp32")) every Function / helper above is
# replaced by its namesake in engine_fp32.py (same argument lists), so modeling_vlbert.py is precision agnostic.
# ======================================================================================
def _install_precision_dispatch():
    from . import engine_fp32 as F

    def runtime_of(args):
        for x in args:
            if isinstance(x, Runtime):
                return x
        raise TypeError("no Runtime among the arguments")

    def wrap_function(name, bf16_cls):
        f32_cls = getattr(F, name)

        class Dispatch:
            __doc__ = bf16_cls.__doc__
            bf16, fp32 = bf16_cls, f32_cls

            @staticmethod
            def apply(*args):
                return (f32_cls if runtime_of(args).fp32 else bf16_cls).apply(*args)

        Dispatch.__name__ = name
        return Dispatch

    g = globals()
    for name in ("EncoderFn", "EmbedFn", "VisInputFn", "ClsProjNormFn", "ClsDenseFn", "HeadTransformFn", "VocabCEFn",
                 "DecoderFn", "BCEFn", "LinearFn", "SmallHeadFn", "SimFn", "ConcatRowsFn", "GatherRowsFn", "WRAFn",
                 "ClsRegionScoreFn"):
        g[name] = wrap_function(name, g[name])
    for name in ("encoder", "sim_matrix", "decoder_logits", "decoder_backward"):
        def make(bf16_fn, f32_fn):
            def dispatch(rt, *a, **kw):
                return (f32_fn if rt.fp32 else bf16_fn)(rt, *a, **kw)
            dispatch.__doc__ = bf16_fn.__doc__
            return dispatch
        g[name] = make(g[name], getattr(F, name))


_install_precision_dispatch()
