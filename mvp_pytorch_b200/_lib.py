"""ctypes binding of libmvptr_b200.so (the C-ABI declared in include/mvptr_b200.h).

There is no CPU fallback: importing succeeds without the library so that CPU-only
tooling can introspect the package, but any compute call raises if the CUDA
library cannot be loaded.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmvptr_b200.so")
_lib = None


class MvptrError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("D", ctypes.c_void_p),
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("lda", ctypes.c_int), ("ldb", ctypes.c_int), ("ldd", ctypes.c_int),
        ("a_mn", ctypes.c_int), ("b_mn", ctypes.c_int),
        ("d_is_f32", ctypes.c_int), ("accumulate", ctypes.c_int), ("split_k", ctypes.c_int),
        ("alpha", ctypes.c_float),
        ("bias", ctypes.c_void_p), ("bias_is_bf16", ctypes.c_int),
        ("pre_act", ctypes.c_void_p), ("act", ctypes.c_int),
        ("gelu_grad_of", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("ld_aux", ctypes.c_int),
        ("p_drop", ctypes.c_float), ("seed", ctypes.c_uint32), ("block_n", ctypes.c_int), ("cta_pair", ctypes.c_int),
        ("colsum", ctypes.c_void_p), ("aux_is_gelu_grad", ctypes.c_int),
    ]


class LayerArgs(ctypes.Structure):
    _P = ctypes.c_void_p
    _fields_ = (
        [(n, ctypes.c_int) for n in ("B", "L", "H", "I", "nh", "save")]
        + [(n, ctypes.c_float) for n in ("eps", "p_hidden", "p_attn")]
        + [(n, ctypes.c_uint32) for n in ("seed_attn", "seed1", "seed2")]
        + [(n, ctypes.c_void_p) for n in (
            "maskadd",
            "w_qkv", "b_qkv", "w_o", "b_o", "ln1_g", "ln1_b", "w_i", "b_i", "w_o2", "b_o2", "ln2_g", "ln2_b",
            "x", "qkv", "att", "lse", "pre1", "st1", "a1", "pre_g", "inter", "pre2", "st2", "out", "tmp",
            "dout", "dx", "dpre2", "dpre2d", "dpre_g", "da1", "dpre1", "dpre1d", "datt", "dqkv",
            "g_w_qkv", "g_b_qkv", "g_w_o", "g_b_o", "g_ln1_g", "g_ln1_b", "g_w_i", "g_b_i", "g_w_o2", "g_b_o2",
            "g_ln2_g", "g_ln2_b")])


_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float,
       "u": ctypes.c_uint32, "z": ctypes.c_size_t}
# argument kinds of every entry point declared in include/mvptr_b200.h (stream last)
SIGNATURES = {
    "mvptr_gemm": "pp",
    "mvptr_layer_fwd": "pp",
    "mvptr_layer_bwd": "pp",
    "mvptr_embed_ln_fwd": "ppppppppp" + "il" + "ppp" + "iiifiii" + "fup",
    "mvptr_embed_bwd": "pppppp" + "iiiiii" + "p",
    "mvptr_ln_fwd": "pppp" + "il" + "pp" + "iif" + "fup",
    "mvptr_add_ln_fwd": "ppfu" + "pppp" + "pp" + "iifp",
    "mvptr_gelu_fwd": "ppzp",
    "mvptr_gelu_bwd_colsum": "ppppiip",
    "mvptr_ln_bwd": "p" + "il" + "pppp" + "pp" + "ppp" + "ii" + "fufu" + "p",
    "mvptr_colsum": "pipiip",
    "mvptr_pad_cast": "pilpiiip",
    "mvptr_mask_prepare": "pipiipppip",
    "mvptr_concat_rows": "pipiipppiip",
    "mvptr_concat_rows_bwd": "piiipppp" + "iip",
    "mvptr_gather_rows": "pppiip",
    "mvptr_scatter_rows_add": "pppiip",
    "mvptr_cast_f32_bf16": "ppzp",
    "mvptr_add_cast": "pppzp",
    "mvptr_cast_bf16_f32": "ppzip",
    "mvptr_attn_set_path": "ii",
    "mvptr_attn_fwd": "pippip" + "iiii" + "fup",
    "mvptr_attn_bwd": "pipppippp" + "iiii" + "fup",
    "mvptr_ce_fwd": "pipiiipppp",
    "mvptr_ce_bwd": "pipiii" + "ppp" + "pip",
    "mvptr_l2norm_fwd": "ppppiip",
    "mvptr_l2norm_bwd": "ppppiip",
    "mvptr_vsc_fwd": "pipppppp" + "p",
    "mvptr_vsc_bwd": "pippppppp",
    "mvptr_small_head_fwd": "pipppiiip",
    "mvptr_small_head_bwd": "ppippippiiip",
    "mvptr_small_ce": "ppiipppp",
    "mvptr_gemm_set_max_ctas": "i",
    "mvptr_adamw": "ppppp" + "zz" + "fffff" + "ii" + "pf" + "pp",
    "mvptr_set_dropout_epoch": "pp",
    "mvptr_step_params": "pippp",
    "mvptr_sumsq": "pzpp",
    "mvptr_topk_rows": "pliiippp",
    "mvptr_match_prob": "ppip",
    "mvptr_wra_fwd": "piiippppp" + "ipppp" + "p",
    "mvptr_wra_bwd": "piiipppppp" + "p" + "p",
    "mvptr_cls_region_score_fwd": "piiiiii" + "pp" + "fu" + "p",
    "mvptr_cls_region_score_bwd": "piiiiii" + "pppp" + "fu" + "p",
    "mvptr_gelu_bwd": "pppzp",
    "mvptr_bce_fwd": "pipiipp",
    "mvptr_bce_bwd": "pipiippip",
    "mvptr_b64_decode_features": "pppp" + "iiiiii" + "pp",
    "mvptr_mlm_mask": "ppppppp" + "i" + "pp" + "ii" + "llll" + "up",
    # fp32 verification tier (csrc/fp32_tier.cu)
    "mvptr_f32_split3": "pliipppip",
    "mvptr_f32_ln_fwd": "ppppp" + "il" + "ppp" + "iif" + "p",
    "mvptr_f32_ln_bwd": "p" + "il" + "ppppppp" + "ii" + "p",
    "mvptr_f32_embed_ln_fwd": "ppppppppp" + "il" + "ppp" + "iiifiii" + "p",
    "mvptr_f32_embed_bwd": "pppppp" + "iii" + "p",
    "mvptr_f32_act": "pzip",
    "mvptr_f32_act_bwd": "pppzip",
    "mvptr_f32_attn_fwd": "pippip" + "iiii" + "p",
    "mvptr_f32_attn_bwd": "pippip" + "iiii" + "p",
    "mvptr_f32_colsum": "pipiip",
    "mvptr_f32_small_head_fwd": "plpppiiip",
    "mvptr_f32_small_head_bwd": "pplppppiiip",
    "mvptr_f32_concat_rows_bwd": "piiipppp" + "iip",
    "mvptr_f32_scatter_rows_add": "pppiip",
    "mvptr_f32_l2norm_bwd": "ppppiip",
    "mvptr_f32_ce_bwd": "pipiii" + "ppp" + "pip",
    "mvptr_f32_bce_bwd": "pipiippip",
    "mvptr_f32_wra_fwd": "piiippppp" + "ipppp" + "p",
    "mvptr_f32_wra_bwd": "piiipppppp" + "p" + "p",
}


class LaunchProfiler:
    """Brackets every C-ABI call with CUDA events on the launching stream (bench.py uses it for the
    per-kernel roofline numbers).  Adds host overhead, so it is only enabled for dedicated
    instrumented steps, never inside a timed region."""

    def __init__(self):
        self.records = []  # (name, work, start_event, end_event)

    def wrap(self, name, work, fn):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        self.records.append((name, work, s, e))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, work, s, e in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["work"] += work
        return out


PROFILER = None


def set_attention_path(fwd="auto", bwd="auto"):
    """Attention kernel family per direction for L <= 128: 'auto' (the faster one per shape, as measured), 'tc'
    (tcgen05 / TMA / TMEM, csrc/attention_tc.cu), 'mma' (mma.sync, csrc/attention.cu), 'env' (MVPTR_ATTN_*_TC)."""
    m = {"env": -1, "auto": 0, "tc": 1, "mma": 2}
    check(lib().mvptr_attn_set_path(m[fwd], m[bwd]), "mvptr_attn_set_path")


def set_gemm_max_ctas(n):
    """Persistent CTAs a GEMM may occupy (0 = all SMs); see include/mvptr_b200.h."""
    check(lib().mvptr_gemm_set_max_ctas(int(n)), "mvptr_gemm_set_max_ctas")


def launch_count():
    """Kernels launched by libmvptr_b200.so in this process so far (counted at the launch sites)."""
    return int(lib().mvptr_launch_count())


def profile_enable(on):
    lib().mvptr_profile_enable(int(bool(on)))


def profile_collect(cap=100000, raw=False):
    """{kernel name: {launches, ms, work}} of the launches recorded since profile_enable(True)
    (raw=True: the list of (name, work, ms) per launch)."""
    L = lib()
    names = (ctypes.c_char_p * cap)()
    work = (ctypes.c_double * cap)()
    ms = (ctypes.c_float * cap)()
    L.mvptr_profile_collect.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    n = L.mvptr_profile_collect(names, work, ms, cap)
    if raw:
        return [(names[i].decode(), work[i], ms[i]) for i in range(n)]
    out = {}
    for i in range(n):
        d = out.setdefault(names[i].decode(), {"launches": 0, "ms": 0.0, "work": 0.0})
        d["launches"] += 1
        d["ms"] += ms[i]
        d["work"] += work[i]
    return out


def call(name, *args):
    """Call an entry point; torch tensors are passed as device pointers, None as NULL,
    and the current CUDA stream is appended."""
    L = lib()
    conv = []
    for a in args:
        if a is None:
            conv.append(None)
        elif isinstance(a, torch.Tensor):
            conv.append(a.data_ptr())
        else:
            conv.append(a)
    conv.append(_raw_stream())
    if PROFILER is not None:
        box = []
        PROFILER.wrap(name, 0.0, lambda: box.append(getattr(L, name)(*conv)))
        rc = box[0]
    else:
        rc = getattr(L, name)(*conv)
    if rc != 0:
        raise MvptrError(f"{name} failed (rc={rc}): {L.mvptr_last_error().decode()}")


def layer_call(name, args, n_kernels):
    """mvptr_layer_fwd / mvptr_layer_bwd with a filled LayerArgs."""
    L = lib()
    stream = _raw_stream()
    if PROFILER is not None:
        box = []
        PROFILER.wrap(name, 0.0, lambda: box.append(getattr(L, name)(ctypes.byref(args), stream)))
        rc = box[0]
    else:
        rc = getattr(L, name)(ctypes.byref(args), stream)
    if rc != 0:
        raise MvptrError(f"{name} failed (rc={rc}): {L.mvptr_last_error().decode()}")


def lib():
    """The loaded shared library; raises loudly when it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MvptrError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). mvp_pytorch_b200 has no CPU or eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.mvptr_last_error.restype = ctypes.c_char_p
        L.mvptr_abi_version.restype = ctypes.c_int
        L.mvptr_launch_count.restype = ctypes.c_ulonglong
        for name, spec in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch: fail loudly
            fn.argtypes = [_CT[c] for c in spec]
            fn.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise MvptrError(f"{what} failed (rc={rc}): {lib().mvptr_last_error().decode()}")


def _raw_stream():
    """Current CUDA stream handle of the current device (the fast C accessor, ~0.3 us)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(_raw_stream())


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


ACT = {None: 0, "none": 0, "gelu": 1, "tanh": 2}


def gemm(A, B, D, M, N, K, *, lda, ldb, ldd, a_mn=False, b_mn=False, accumulate=False, split_k=1, alpha=1.0,
         bias=None, pre_act=None, act=None, gelu_grad_of=None, residual=None, ld_aux=0, p_drop=0.0, seed=0,
         block_n=0, cta_pair=0, colsum=None, aux_is_gelu_grad=False):
    """D[M,N] (+)= epilogue(alpha * A . B^T); see include/mvptr_b200.h."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert D.dtype in (torch.bfloat16, torch.float32)
    g = GemmArgs()
    g.A, g.B, g.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = lda, ldb, ldd
    g.a_mn, g.b_mn = int(a_mn), int(b_mn)
    g.d_is_f32 = int(D.dtype == torch.float32)
    g.accumulate, g.split_k, g.alpha = int(accumulate), int(split_k), float(alpha)
    if bias is not None:
        g.bias, g.bias_is_bf16 = bias.data_ptr(), int(bias.dtype == torch.bfloat16)
        assert bias.dtype in (torch.bfloat16, torch.float32)
    for name, t in (("pre_act", pre_act), ("gelu_grad_of", gelu_grad_of), ("residual", residual)):
        if t is not None:
            assert t.dtype == torch.bfloat16
            setattr(g, name, t.data_ptr())
    g.act = ACT[act]
    g.ld_aux, g.p_drop, g.seed, g.block_n = ld_aux, float(p_drop), int(seed) & 0xFFFFFFFF, block_n
    g.cta_pair = cta_pair
    g.aux_is_gelu_grad = int(aux_is_gelu_grad)
    if colsum is not None:
        assert colsum.dtype == torch.float32
        g.colsum = colsum.data_ptr()
    if PROFILER is not None:
        kind = "mvptr_gemm[%s%s]" % ("mn" if a_mn else "k", "mn" if b_mn else "k")
        box = []
        PROFILER.wrap(kind, 2.0 * M * N * K, lambda: box.append(lib().mvptr_gemm(ctypes.byref(g), stream_ptr())))
        check(box[0], "mvptr_gemm")
    else:
        check(lib().mvptr_gemm(ctypes.byref(g), stream_ptr()), "mvptr_gemm")
    return D
