"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/mvptr_b200.h
declares, with the argument kinds the ctypes binding assumes.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from mvp_pytorch_b200 import _lib
    return _lib


def _header_functions():
    hdr = open(os.path.join(ROOT, "include", "mvptr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*|unsigned long long)\s+(mvptr_\w+)\(([^;{}]*?)\);", hdr, flags=re.S):
        out[m.group(1)] = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
    return out


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built.LIB_PATH)
    fns = _header_functions()
    assert len(fns) >= 30
    for name in fns:
        assert hasattr(L, name), f"{name} declared in include/mvptr_b200.h but not exported"
    assert built.lib().mvptr_abi_version() == 1
    assert built.lib().mvptr_last_error() is not None


def test_ctypes_signatures_match_header(built):
    fns = _header_functions()
    for name, spec in built.SIGNATURES.items():
        args = fns[name]
        kinds = ""
        for a in args:
            if "*" in a: kinds += "p"
            elif a.startswith("long long"): kinds += "l"
            elif a.startswith("float"): kinds += "f"
            elif a.startswith("uint32_t"): kinds += "u"
            elif a.startswith("size_t"): kinds += "z"
            else: kinds += "i"
        assert kinds == spec, f"{name}: header {kinds} vs binding {spec}"
    declared = set(fns) - {"mvptr_abi_version", "mvptr_last_error", "mvptr_launch_count", "mvptr_wra_max_phrases", "mvptr_profile_enable",
                               "mvptr_profile_collect"}
    assert declared == set(built.SIGNATURES), declared ^ set(built.SIGNATURES)


def test_gemm_args_struct_matches_header(built):
    hdr = open(os.path.join(ROOT, "include", "mvptr_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} mvptr_gemm_args;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(part.replace("*", " ").split()[-1])
    assert names == [f[0] for f in built.GemmArgs._fields_]


def test_layer_args_struct_matches_header(built):
    hdr = open(os.path.join(ROOT, "include", "mvptr_b200.h")).read()
    body = re.search(r"typedef struct \{([^}]*?)\} mvptr_layer_args;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(part.replace("*", " ").split()[-1])
    assert names == [f[0] for f in built.LayerArgs._fields_]


def test_argument_errors_do_not_need_a_gpu(built):
    g = built.GemmArgs()
    rc = built.lib().mvptr_gemm(ctypes.byref(g), None)
    assert rc == -1 and b"null operand" in built.lib().mvptr_last_error()


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mvp_pytorch_b200.modeling_vlbert as mv
    from mvp_pytorch_b200.modeling_utils import BertConfig
    from mvp_pytorch_b200._lib import MvptrError
    c = BertConfig(vocab_size_or_config_json_file=100, hidden_size=64, num_hidden_layers=2, num_attention_heads=1,
                   intermediate_size=128, max_position_embeddings=32)
    c.only_word_size, c.qa_answer_size, c.img_feature_dim, c.img_feature_type = 50, 7, 22, "faster_r-cnn"
    c.use_img_layernorm, c.img_layer_norm_eps, c.loss_type = 1, 1e-12, "sfmx"
    m = mv.BiImageBertRep(c)
    ids = torch.zeros(2, 4, dtype=torch.long)
    with pytest.raises(MvptrError):
        m(input_ids_a=ids, input_ids_b=ids, img_feats=torch.zeros(2, 3, 22))
    with pytest.raises(RuntimeError):  # parameter containers have no eager forward to fall back to
        m.bert.pooler(torch.zeros(2, 4, 64))


def test_every_kernel_file_that_hashes_dropout_masks_owns_a_registered_epoch_word():
    """The dropout epoch (common.cuh: g_dropout_epoch) is one device word PER translation unit.  A kernel file that
    calls site_seed() without its own MVPTR_DEFINE_EPOCH_SETTER, or whose setter api.cu does not drive from both
    mvptr_set_dropout_epoch and mvptr_step_params, hashes epoch 0 for ever: forward and backward kernels in different
    files then disagree about the keep mask as soon as a CUDA-graph replay advances the epoch."""
    import glob
    import re
    csrc = os.path.join(ROOT, "mvp_pytorch_b200", "csrc")
    api = open(os.path.join(csrc, "api.cu")).read()
    users = 0
    for path in sorted(glob.glob(os.path.join(csrc, "*.cu"))):
        src = open(path).read()
        if "site_seed(" not in src:
            continue
        users += 1
        m = re.search(r"MVPTR_DEFINE_EPOCH_SETTER\((\w+)\)", src)
        assert m, f"{os.path.basename(path)} hashes dropout masks but defines no epoch setter"
        fn = m.group(1)
        assert re.search(rf"\b{fn}\(src, stream\)", api), f"mvptr_set_dropout_epoch does not call {fn}"
        assert re.search(rf"\b{fn}_addr\(&a\[\d\]\)", api), f"mvptr_step_params does not write {fn}'s epoch word"
    assert users >= 5
