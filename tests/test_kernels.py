"""GPU: every C-ABI kernel against a plain PyTorch fp32 restatement of the same op on the same
seeded inputs (bf16 inputs are shared, so only accumulation order / output rounding differ)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def lib():
    from mvp_pytorch_b200 import _lib
    _lib.lib()
    return _lib


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(BF16)


def assert_close(got, ref, rtol, atol, what=""):
    err = (got.float() - ref.float()).abs()
    tol = atol + rtol * ref.float().abs()
    assert (err <= tol).all(), f"{what}: max err {err.max().item():.5f}, ref absmax {ref.abs().max().item():.4f}, " \
                               f"bad {(err > tol).float().mean().item():.4%}"


# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,f32", [
    (300, 768, 768, False, False, False), (1000, 3072, 768, False, False, False),
    (777, 768, 3072, False, True, False), (768, 3072, 5000, True, True, True), (130, 1002, 768, False, False, True),
    (129, 64, 2054, False, False, False)])
def test_gemm_layouts(lib, M, N, K, a_mn, b_mn, f32):
    kp = (K + 7) // 8 * 8
    A = rnd(M, K, seed=1); B = rnd(N, K, seed=2)
    if a_mn:
        A_st, lda = A.t().contiguous(), M
    else:
        A_st = torch.zeros(M, kp, device="cuda", dtype=BF16); A_st[:, :K] = A; lda = kp
    if b_mn:
        B_st, ldb = B.t().contiguous(), N
    else:
        B_st = torch.zeros(N, kp, device="cuda", dtype=BF16); B_st[:, :K] = B; ldb = kp
    pitch = (N + 7) // 8 * 8
    D = torch.zeros(M, pitch, device="cuda", dtype=F32 if f32 else BF16)
    bias = rnd(N, seed=3)
    lib.gemm(A_st, B_st, D, M, N, K, lda=lda, ldb=ldb, ldd=pitch, a_mn=a_mn, b_mn=b_mn, bias=bias,
             accumulate=a_mn and b_mn, split_k=5 if (a_mn and b_mn) else 1)
    ref = A.float() @ B.float().t() + bias.float()
    assert_close(D[:, :N], ref, 1e-2 if not f32 else 1e-4, 0.05 if not f32 else 2e-3, "gemm")


def test_gemm_epilogues(lib):
    M, N, K = 515, 3072, 768
    A, B, bias = rnd(M, K, scale=0.5, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3)
    out = torch.empty(M, N, device="cuda", dtype=BF16); pre = torch.empty_like(out)
    lib.gemm(A, B, out, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="gelu", pre_act=pre, ld_aux=N)
    z = A.float() @ B.float().t() + bias.float()
    assert_close(pre, z, 1e-2, 2e-2, "pre_act")
    assert_close(out, F.gelu(z), 1e-2, 2e-2, "gelu")
    # backward-style epilogue: acc * gelu'(pre) ; and residual add
    res = rnd(M, N, seed=4)
    out2 = torch.empty(M, N, device="cuda", dtype=BF16)
    lib.gemm(A, B, out2, M, N, K, lda=K, ldb=K, ldd=N, gelu_grad_of=pre, residual=res, ld_aux=N)
    x = pre.float().requires_grad_(True)
    F.gelu(x).sum().backward()
    assert_close(out2, (A.float() @ B.float().t()) * x.grad + res.float(), 1e-2, 3e-2, "gelu_grad+residual")
    # dropout: kept elements scaled by 1/keep, drop rate ~ p, identical mask for identical seed
    o3 = torch.empty(M, N, device="cuda", dtype=F32); o4 = torch.empty_like(o3)
    lib.gemm(A, B, o3, M, N, K, lda=K, ldb=K, ldd=N, p_drop=0.1, seed=77)
    lib.gemm(A, B, o4, M, N, K, lda=K, ldb=K, ldd=N, p_drop=0.1, seed=77)
    assert torch.equal(o3, o4)
    raw = A.float() @ B.float().t()
    dropped = (o3 == 0) & (raw.abs() > 1e-3)
    assert abs(dropped.float().mean().item() - 0.1) < 0.01
    assert_close(o3[~dropped], raw[~dropped] / 0.9, 3e-3, 3e-3, "kept")


@pytest.mark.parametrize("M,N", [(515, 3072), (1000, 320), (4100, 512)])
def test_gemm_lean_gelu_epilogues(lib, M, N):
    """The lean tcgen05 epilogues of the FFN: bias+GELU (single store), bias+GELU with the pre-activation
    saved (DUAL TMA stores, CTA pairs) and the dgrad epilogue acc * gelu'(pre) with the fused bias-gradient
    column sums (pre tiles fetched by TMA).  N = 320 leaves a half-empty 256-wide tile, M a ragged row tile."""
    K = 768
    A, B, bias = rnd(M, K, scale=0.5, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3)
    z = A.float() @ B.float().t() + bias.float()
    out = torch.empty(M, N, device="cuda", dtype=BF16)
    lib.gemm(A, B, out, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="gelu")
    assert_close(out, F.gelu(z), 1e-2, 2e-2, "lean gelu")
    for pair in (1, 2):
        out2 = torch.zeros(M, N, device="cuda", dtype=BF16); pre = torch.zeros_like(out2)
        lib.gemm(A, B, out2, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="gelu", pre_act=pre, ld_aux=N, cta_pair=pair)
        assert_close(pre, z, 1e-2, 2e-2, f"dual pre (cta_pair={pair})")
        assert torch.equal(out2, out), "dual and single-store GELU tiles must agree bit for bit"
    # dgrad: dX[M, N] = dY[M, K2] . W[K2, N] (W stored [K2][N]: MN-major B), times gelu'(pre), column sums
    K2 = 256
    dY, W = rnd(M, K2, scale=0.5, seed=5), rnd(K2, N, scale=0.05, seed=6)
    dx = torch.zeros(M, N, device="cuda", dtype=BF16)
    cs = torch.full((N,), 0.5, device="cuda", dtype=F32)  # accumulates (+=)
    lib.gemm(dY, W, dx, M, N, K2, lda=K2, ldb=N, ldd=N, b_mn=True, gelu_grad_of=pre, ld_aux=N, colsum=cs)
    x = pre.float().requires_grad_(True)
    F.gelu(x).sum().backward()
    ref = (dY.float() @ W.float()) * x.grad
    assert_close(dx, ref, 1e-2, 2e-2, "acc * gelu'(pre)")
    assert_close(cs, 0.5 + dx.float().sum(0), 1e-3, 2e-2, "fused bias-gradient column sums")


# ------------------------------------------------------------------------------------
def ln_ref(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return w * ((x - mu) / torch.sqrt(var + eps)) + b


@pytest.mark.parametrize("H", [768, 128])
def test_layernorm_fwd_bwd(lib, H):
    rows = 1000
    x, w, b, dy = rnd(rows, H, seed=1), (1 + 0.1 * rnd(H, seed=2).float()).to(BF16), rnd(H, scale=0.1, seed=3), rnd(rows, H, seed=4)
    y = torch.empty_like(x); st = torch.empty(2, rows, device="cuda", dtype=F32)
    lib.call("mvptr_ln_fwd", x, w, b, y, 0, 0, st[0], st[1], rows, H, 1e-12, 0.0, 0)
    xr = x.float().requires_grad_(True); wr = w.float().requires_grad_(True); br = b.float().requires_grad_(True)
    yr = ln_ref(xr, wr, br, 1e-12)
    assert_close(y, yr, 1e-2, 1e-2, "ln fwd")
    yr.backward(dy.float())
    dx = torch.empty_like(x); dg = torch.zeros(H, device="cuda"); db = torch.zeros(H, device="cuda"); dbias = torch.zeros(H, device="cuda")
    lib.call("mvptr_ln_bwd", dy, 0, 0, x, st[0], st[1], w, dx, None, dg, db, dbias, rows, H, 0.0, 0, 0.0, 0)
    assert_close(dx, xr.grad, 2e-2, 2e-2, "ln dx")
    assert_close(dg, wr.grad, 1e-2, 0.05, "ln dgamma")
    assert_close(db, br.grad, 1e-2, 0.05, "ln dbeta")
    assert_close(dbias, dx.float().sum(0), 1e-3, 0.05, "dbias = colsum(dx)")


def test_embed_ln_fwd_bwd(lib):
    B, L, H, V = 7, 13, 768, 500
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, V, (B, L), generator=g).cuda(); ids[0, :3] = 0
    seg = torch.randint(0, 2, (B, L), generator=g).cuda()
    word, pos, typ = rnd(V, H, scale=0.02, seed=1), rnd(64, H, scale=0.02, seed=2), rnd(2, H, scale=0.02, seed=3)
    w, b = (1 + 0.1 * rnd(H, seed=4).float()).to(BF16), rnd(H, scale=0.1, seed=5)
    y = torch.empty(B, L, H, device="cuda", dtype=BF16); pre = torch.empty(B * L, H, device="cuda", dtype=BF16)
    st = torch.empty(2, B * L, device="cuda", dtype=F32)
    lib.call("mvptr_embed_ln_fwd", ids, seg, None, word, pos, typ, w, b, y, 0, 0, pre, st[0], st[1], B, L, H, 1e-12, V, 64, 2, 0.0, 0)
    wf, pf, tf = word.float().requires_grad_(True), pos.float().requires_grad_(True), typ.float().requires_grad_(True)
    e = wf[ids] + pf[torch.arange(L, device="cuda")][None] + tf[seg]
    yr = ln_ref(e, w.float(), b.float(), 1e-12)
    assert_close(y, yr, 1e-2, 1e-2, "embed fwd")
    dy = rnd(B, L, H, seed=6)
    # reference: padding_idx=0 row receives no gradient (nn.Embedding(padding_idx=0))
    yr.backward(dy.float())
    dpre = torch.empty(B * L, H, device="cuda", dtype=BF16)
    dg = torch.zeros(H, device="cuda"); db = torch.zeros(H, device="cuda")
    lib.call("mvptr_ln_bwd", dy, 0, 0, pre, st[0], st[1], w, dpre, None, dg, db, None, B * L, H, 0.0, 0, 0.0, 0)
    dW = torch.zeros(V, H, device="cuda"); dP = torch.zeros(64, H, device="cuda"); dT = torch.zeros(2, H, device="cuda")
    lib.call("mvptr_embed_bwd", dpre, ids, seg, dW, dP, dT, B, L, H, V, 2, 0)
    ref_w = wf.grad.clone(); ref_w[0] = 0
    assert_close(dW, ref_w, 2e-2, 0.5, "dword")  # grads ~1e2; dpre is bf16
    assert_close(dP, pf.grad, 2e-2, 1.0, "dpos")  # sums of B bf16 rows, magnitudes ~1e2
    assert_close(dT, tf.grad, 2e-2, 2.0, "dtype")


# ------------------------------------------------------------------------------------
def attn_ref(qkv, maskadd, B, L, nh, H):
    q, k, v = qkv.float().view(B, L, 3, nh, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0 + maskadd[:, None, None, :]
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * L, H)


@pytest.fixture(params=["auto", "tc", "mma"])
def attn_path(request, lib):
    """Every attention test runs on the tcgen05 kernels (L <= 128), on the mma.sync kernels, and on the default mix."""
    lib.set_attention_path(request.param, request.param)
    yield request.param
    lib.set_attention_path("auto", "auto")


@pytest.mark.parametrize("L", [7, 16, 35, 40, 70, 90, 105, 128, 133, 183, 193])
def test_attention_fwd_bwd(lib, L, attn_path):
    B, nh = 3, 12
    H = nh * 64
    qkv = rnd(B * L, 3 * H, seed=L)
    mask = torch.ones(B, L, device="cuda")
    mask[0, L - 5:] = 0
    mask[1, 3:7] = 0  # disjoint valid segments
    maskadd = ((1 - mask) * -10000.0).contiguous()
    ctx = torch.empty(B * L, H, device="cuda", dtype=BF16); lse = torch.empty(B, nh, L, device="cuda", dtype=F32)
    lib.call("mvptr_attn_fwd", qkv, 3 * H, maskadd, ctx, H, lse, B, L, nh, H, 0.0, 0)
    x = qkv.float().requires_grad_(True)
    ref = attn_ref(x, maskadd, B, L, nh, H)
    assert_close(ctx, ref, 1e-2, 1e-2, "attn fwd")
    dctx = rnd(B * L, H, seed=5)
    ref.backward(dctx.float())
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros(3 * H, device="cuda")
    lib.call("mvptr_attn_bwd", qkv, 3 * H, maskadd, ctx, dctx, H, lse, dqkv, dbias, B, L, nh, H, 0.0, 0)
    assert_close(dbias, x.grad.sum(0), 2e-2, 0.3, "fused qkv bias grad")
    rel = ((dqkv.float() - x.grad).norm() / x.grad.norm()).item()
    assert rel < 2e-2, f"attn bwd rel l2 {rel}"
    assert_close(dqkv, x.grad, 3e-2, 3e-2, "attn bwd")


def test_attention_dropout_consistency(lib, attn_path):
    """fwd and bwd regenerate the same keep mask: check d(ctx)/dV against finite structure."""
    B, L, nh = 2, 40, 2
    H = nh * 64
    qkv = rnd(B * L, 3 * H, seed=1)
    maskadd = torch.zeros(B, L, device="cuda")
    ctx = torch.empty(B * L, H, device="cuda", dtype=BF16); lse = torch.empty(B, nh, L, device="cuda", dtype=F32)
    lib.call("mvptr_attn_fwd", qkv, 3 * H, maskadd, ctx, H, lse, B, L, nh, H, 0.3, 99)
    # recover the effective (dropped, rescaled) probabilities from ctx = P_drop V by solving against V
    q, k, v = qkv.float().view(B, L, 3, nh, 64).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
    dctx = rnd(B * L, H, seed=2)
    dqkv = torch.empty_like(qkv)
    lib.call("mvptr_attn_bwd", qkv, 3 * H, maskadd, ctx, dctx, H, lse, dqkv, None, B, L, nh, H, 0.3, 99)
    # dV = P_drop^T dO ; ctx = P_drop V  ->  <dV, V> == <dO, ctx> for every head (adjoint identity)
    dv = dqkv.float().view(B, L, 3, nh, 64)[:, :, 2]
    vv = qkv.float().view(B, L, 3, nh, 64)[:, :, 2]
    lhs = (dv * vv).sum(dim=(1, 3))
    rhs = (dctx.float().view(B, L, nh, 64) * ctx.float().view(B, L, nh, 64)).sum(dim=(1, 3))
    assert_close(lhs, rhs, 3e-2, 0.5, "dropout adjoint identity")
    assert (ctx.float() - (p @ v).permute(0, 2, 1, 3).reshape(B * L, H)).abs().max() > 0.05  # dropout did something


@pytest.fixture(params=[0, 7], ids=["epoch0", "epoch7"])
def dropout_epoch(request, lib):
    """The per-replay dropout epoch word (mvptr_set_dropout_epoch) lives once per translation unit of the library:
    a forward in one unit (attention_tc.cu) and a backward in another (attention.cu) must see the same value."""
    word = torch.tensor([request.param], dtype=torch.int32, device="cuda")
    lib.call("mvptr_set_dropout_epoch", word, None)
    torch.cuda.synchronize()
    yield request.param
    word.zero_()
    lib.call("mvptr_set_dropout_epoch", word, None)
    torch.cuda.synchronize()


@pytest.mark.parametrize("L", [40, 90, 150])
def test_attention_dropout_matches_autograd_with_the_extracted_mask(lib, L, attn_path, dropout_epoch):
    """The keep mask depends only on (seed, epoch, batch, head, query, key): extract it by pushing one-hot V
    columns through the forward kernel, then check forward AND backward (dQ, dK, dV) against torch
    autograd using that exact mask.  L=40/90 run the shared-memory backward, L=150 the recompute one."""
    B, nh, pd, seed = 2, 3, 0.25, 1234
    H = nh * 64
    maskadd = torch.zeros(B, L, device="cuda")
    maskadd[1, L - 4:] = -10000.0
    keep = torch.zeros(B, nh, L, L, device="cuda")
    for c in range((L + 63) // 64):
        probe = torch.zeros(B, L, 3, nh, 64, device="cuda")
        for k in range(c * 64, min(L, c * 64 + 64)):
            probe[:, k, 2, :, k - c * 64] = 1.0  # V[k, :] = e_(k - 64c); Q = K = 0 -> uniform P
        ctx = torch.empty(B * L, H, device="cuda", dtype=BF16)
        lib.call("mvptr_attn_fwd", probe.view(B * L, 3 * H).to(BF16), 3 * H, torch.zeros(B, L, device="cuda"), ctx, H,
                 None, B, L, nh, H, pd, seed)
        pdrop = ctx.float().view(B, L, nh, 64).permute(0, 2, 1, 3)  # [B, nh, q, k - 64c]
        n = min(64, L - c * 64)
        keep[:, :, :, c * 64:c * 64 + n] = (pdrop[..., :n] > 0).float()
    rate = keep.mean().item()
    assert abs(rate - (1 - pd)) < 0.02, rate
    qkv = rnd(B * L, 3 * H, seed=L + 1)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.view(B, L, 3, nh, 64).permute(2, 0, 3, 1, 4)
    pr = torch.softmax(q @ k.transpose(-1, -2) / 8.0 + maskadd[:, None, None, :], -1)
    ref = ((pr * keep / (1 - pd)) @ v).permute(0, 2, 1, 3).reshape(B * L, H)
    ctx = torch.empty(B * L, H, device="cuda", dtype=BF16); lse = torch.empty(B, nh, L, device="cuda", dtype=F32)
    lib.call("mvptr_attn_fwd", qkv, 3 * H, maskadd, ctx, H, lse, B, L, nh, H, pd, seed)
    assert_close(ctx, ref, 2e-2, 2e-2, "attn fwd with dropout")
    dctx = rnd(B * L, H, seed=6)
    ref.backward(dctx.float())
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros(3 * H, device="cuda")
    lib.call("mvptr_attn_bwd", qkv, 3 * H, maskadd, ctx, dctx, H, lse, dqkv, dbias, B, L, nh, H, pd, seed)
    rel = ((dqkv.float() - x.grad).norm() / x.grad.norm()).item()
    assert rel < 2e-2, f"attn bwd with dropout rel l2 {rel}"
    assert_close(dbias, x.grad.sum(0), 2e-2, 0.3, "fused qkv bias grad")


# ------------------------------------------------------------------------------------
def test_ce_fwd_bwd(lib):
    n, V = 37, 30522
    pitch = (V + 7) // 8 * 8
    g = torch.Generator().manual_seed(0)
    logits = torch.zeros(n, pitch, device="cuda"); logits[:, :V] = torch.randn(n, V, generator=g).cuda() * 3
    labels = torch.randint(0, V, (n,), generator=g).cuda(); labels[5] = -1
    lse = torch.empty(n, device="cuda"); acc = torch.zeros(2, device="cuda")
    lib.call("mvptr_ce_fwd", logits, pitch, labels, n, V, -1, lse, acc[0:1], acc[1:2])
    x = logits[:, :V].clone().requires_grad_(True)
    ref = F.cross_entropy(x, labels, ignore_index=-1)
    assert_close(acc[0] / acc[1], ref, 1e-5, 1e-5, "ce")
    ref.backward()
    d = torch.empty(n, pitch, device="cuda", dtype=BF16)
    lib.call("mvptr_ce_bwd", logits, pitch, labels, n, V, -1, lse, acc[1:2], None, d, pitch)
    assert_close(d[:, :V], x.grad, 1e-2, 1e-6, "dlogits")
    assert (d[:, V:] == 0).all()


def test_vsc_and_hard_negatives(lib):
    B = 50
    g = torch.Generator().manual_seed(1)
    sim = (torch.rand(B, B, generator=g) * 2 - 1).cuda()
    ls = torch.tensor(math.log(1 / 0.07), device="cuda")
    lse = torch.empty(2, B, device="cuda"); loss = torch.zeros(1, device="cuda")
    hard = torch.empty(2, B, device="cuda", dtype=torch.int64)
    lib.call("mvptr_vsc_fwd", sim, B, ls, lse[0], lse[1], loss, hard[0], hard[1])
    s = sim.clone().requires_grad_(True); l = ls.clone().requires_grad_(True)
    m = s * l.exp()
    lab = torch.arange(B, device="cuda")
    ref = (F.cross_entropy(m, lab) + F.cross_entropy(m.t(), lab)) / 2
    assert_close(loss[0], ref, 1e-5, 1e-5, "vsc")
    masked = sim - 2 * torch.eye(B, device="cuda")
    assert torch.equal(hard[0], masked.max(1)[1]) and torch.equal(hard[1], masked.max(0)[1])  # bit exact
    ref.backward()
    dsim = torch.empty_like(sim); dls = torch.zeros(1, device="cuda")
    lib.call("mvptr_vsc_bwd", sim, B, ls, lse[0], lse[1], None, dsim, dls)
    assert_close(dsim, s.grad, 1e-4, 1e-6, "dsim")
    assert_close(dls[0], l.grad, 1e-4, 1e-5, "dlogit_scale")


def test_small_head_l2norm_colsum(lib):
    n, H, C = 33, 768, 2
    x, W, b = rnd(n, H, seed=1), rnd(C, H, scale=0.05, seed=2), rnd(C, seed=3)
    logits = torch.empty(n, C, device="cuda")
    lib.call("mvptr_small_head_fwd", x, H, W, b, logits, n, H, C)
    assert_close(logits, x.float() @ W.float().t() + b.float(), 1e-4, 1e-3, "small head")
    labels = torch.randint(0, C, (n,)).cuda()
    labels[::5] = -1  # ignore_index rows (QA loss, modeling_vlbert.py:1262-1264): no loss, no gradient
    acc = torch.zeros(2, device="cuda"); dl = torch.empty_like(logits)
    lib.call("mvptr_small_ce", logits, labels, n, C, acc, None, None)
    lib.call("mvptr_small_ce", logits, labels, n, C, acc, dl, None)
    lg = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lg, labels, ignore_index=-1); ref.backward()
    assert float(acc[1]) == float((labels >= 0).sum())
    assert_close(acc[0] / acc[1], ref, 1e-5, 1e-5, "small ce"); assert_close(dl, lg.grad, 1e-4, 1e-6, "small ce grad")
    assert float(dl[::5].abs().max()) == 0.0
    dx = torch.empty(n, H, device="cuda", dtype=BF16); dW = torch.zeros(C, H, device="cuda"); db = torch.zeros(C, device="cuda")
    lib.call("mvptr_small_head_bwd", dl, x, H, W, dx, H, dW, db, n, H, C)
    assert_close(dx, dl @ W.float(), 1e-2, 1e-4, "small dx"); assert_close(dW, dl.t() @ x.float(), 1e-4, 1e-4, "small dW")
    assert_close(db, dl.sum(0), 1e-4, 1e-6, "small db")
    # l2norm
    v = torch.randn(n, H, device="cuda"); y = torch.empty_like(v); nr = torch.empty(n, device="cuda")
    lib.call("mvptr_l2norm_fwd", v, y, None, nr, n, H)
    vr = v.clone().requires_grad_(True); yr = F.normalize(vr, p=2, dim=-1)
    assert_close(y, yr, 1e-5, 1e-6, "l2norm")
    dy = torch.randn(n, H, device="cuda"); yr.backward(dy)
    dx16 = torch.empty(n, H, device="cuda", dtype=BF16)
    lib.call("mvptr_l2norm_bwd", dy, y, nr, dx16, n, H)
    assert_close(dx16, vr.grad, 1e-2, 1e-3, "l2norm bwd")
    # colsum
    big = rnd(3000, 2304, seed=9); out = torch.zeros(2304, device="cuda")
    lib.call("mvptr_colsum", big, 2304, out, 3000, 2304)
    assert_close(out, big.float().sum(0), 1e-4, 1e-2, "colsum")


def test_concat_gather_mask(lib):
    B, La, Lb, H, col0 = 5, 7, 9, 128, 3
    a, b = rnd(B, La, H, seed=1), rnd(B, Lb, H, seed=2)
    ra = torch.tensor([0, 1, 2, 3, 4, 2, 0], device="cuda"); rb = torch.tensor([0, 1, 2, 3, 4, 4, 1], device="cuda")
    out = torch.empty(7, La + Lb - col0, H, device="cuda", dtype=BF16)
    lib.call("mvptr_concat_rows", a, La, b, Lb, col0, ra, rb, out, 7, H)
    assert torch.equal(out, torch.cat([a[ra], b[rb][:, col0:]], 1))  # bit exact gather
    dout = rnd(7, La + Lb - col0, H, seed=3)
    da = torch.zeros(B, La, H, device="cuda"); db = torch.zeros(B, Lb, H, device="cuda")
    lib.call("mvptr_concat_rows_bwd", dout, La, Lb, col0, ra, rb, da, db, 7, H)
    ra_ref = torch.zeros_like(da).index_add_(0, ra, dout[:, :La].float())
    rb_ref = torch.zeros_like(db); rb_ref[:, col0:] = torch.zeros(B, Lb - col0, H, device="cuda").index_add_(0, rb, dout[:, La:].float())
    assert_close(da, ra_ref, 1e-6, 1e-6, "da"); assert_close(db, rb_ref, 1e-6, 1e-6, "db")
    ma = torch.randint(0, 2, (B, La)).cuda(); mb = torch.randint(0, 2, (B, Lb)).cuda()
    m = torch.empty(7, La + Lb - col0, device="cuda")
    lib.call("mvptr_mask_prepare", ma, La, mb, Lb, col0, ra, rb, m, 7)
    assert torch.equal(m, (1.0 - torch.cat([ma[ra], mb[rb][:, col0:]], 1).float()) * -10000.0)


def test_wra_fwd_bwd(lib):
    from oracle import mvptr_oracle as O
    B, Lt, H, P = 6, 30, 128, 16
    seq = rnd(B, Lt, H, seed=1)
    g = torch.Generator().manual_seed(3)
    phrase_index = torch.tensor([[5, 8], [6, 6], [2, 7], [9, 10], [4, 8], [3, 5]])
    img_index = torch.tensor([[12, 30], [12, 25], [12, 20], [12, 30], [12, 16], [12, 28]])
    neg_img = torch.tensor([1, 0, 5, 2, 0, 3])
    rp = torch.randint(0, 3, (B, P), generator=g); rn = torch.randint(0, 3, (B, P), generator=g)
    out = torch.empty(2, B, device="cuda"); sel = torch.full((2, B, 16), -1, device="cuda", dtype=torch.int32)
    lib.call("mvptr_wra_fwd", seq, B, Lt, H, phrase_index.cuda(), img_index.cuda(), neg_img.cuda(), rp.cuda(), rn.cuda(), P,
             out[0], out[1], sel[0], sel[1])
    x = seq.float().cpu().requires_grad_(True)
    pos, neg = [], []
    for b in range(B):
        p0, p1 = phrase_index[b].tolist(); i0, i1 = img_index[b].tolist(); nb = int(neg_img[b]); n0, n1 = img_index[nb].tolist()
        ph = F.normalize(x[b, p0:p1], dim=-1)
        pos.append(O.t2i_sim(ph @ F.normalize(x[b, i0:i1], dim=-1).t(), rp[b, : p1 - p0]))
        neg.append(O.t2i_sim(ph @ F.normalize(x[nb, n0:n1], dim=-1).t(), rn[b, : p1 - p0]))
    pos, neg = torch.stack(pos), torch.stack(neg)
    assert_close(out[0].cpu(), pos, 1e-4, 1e-5, "wra pos"); assert_close(out[1].cpu(), neg, 1e-4, 1e-5, "wra neg")
    gp, gn = torch.randn(B, generator=g), torch.randn(B, generator=g)
    (pos * gp + neg * gn).sum().backward()
    dseq = torch.zeros(B, Lt, H, device="cuda")
    lib.call("mvptr_wra_bwd", seq, B, Lt, H, phrase_index.cuda(), neg_img.cuda(), sel[0], sel[1], gp.cuda(), gn.cuda(), dseq)
    assert_close(dseq.cpu(), x.grad, 1e-3, 1e-5, "wra dseq")


def test_adamw_matches_reference_trajectory(lib, golden_dir):
    import os
    g = torch.load(os.path.join(golden_dir, "adamw_traj.pt"), weights_only=False)
    p = torch.tensor([0.1, -0.2, -0.1, 0.7], device="cuda"); m = torch.zeros(4, device="cuda"); v = torch.zeros(4, device="cuda")
    p16 = torch.zeros(4, device="cuda", dtype=BF16)
    tgt = torch.tensor([0.4, 0.2, -0.5, 0.1], device="cuda")
    used = [0.02 * x for x in (0.0, 0.5, 1.0, 0.875, 0.75)]
    for it in range(5):
        grad = (p - tgt) * 2
        lib.call("mvptr_adamw", p, grad, m, v, p16, 4, 4, used[it], 0.9, 0.999, 1e-6, 0.01, it + 1, 1, None, 0.0, None)
        assert_close(p.cpu(), g["traj"][it], 1e-5, 1e-6, f"adamw step {it}")
    assert torch.equal(p16, p.to(BF16))
    # no-decay tail + clipping
    n = 4096
    p = torch.randn(n, device="cuda"); p0 = p.clone(); gr = torch.randn(n, device="cuda") * 10
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda"); ss = torch.zeros(1, device="cuda")
    lib.call("mvptr_sumsq", gr, n, ss)
    assert_close(ss[0], (gr * gr).sum(), 1e-4, 1e-2, "sumsq")
    lib.call("mvptr_adamw", p, gr, m, v, None, n, n // 2, 0.01, 0.9, 0.999, 1e-6, 0.1, 1, 1, ss, 1.0, None)
    from oracle import mvptr_oracle as O
    scale = min(1.0, 1.0 / (float(gr.norm()) + 1e-6))
    pr, mr, vr = p0.clone().cpu(), torch.zeros(n), torch.zeros(n)
    a, b2 = pr[: n // 2], pr[n // 2:]
    O.adamw_step(a, gr.cpu()[: n // 2] * scale, mr[: n // 2], vr[: n // 2], 1, 0.01, weight_decay=0.1)
    O.adamw_step(b2, gr.cpu()[n // 2:] * scale, mr[n // 2:], vr[n // 2:], 1, 0.01, weight_decay=0.0)
    assert_close(p.cpu(), pr, 1e-4, 1e-5, "adamw clip + decay boundary")


@pytest.mark.parametrize("K,f32", [(2054, False), (2054, True), (70, False), (71, False), (71, True)])
def test_pad_cast_is_bit_exact(lib, K, f32):
    """Region features [rows, K] -> bf16 [rows, pad8(K)] (modeling_vlbert.py:498-506 input staging): the 8-wide
    pair-load kernel (even K) and the scalar fallback (odd K) both reproduce torch's cast bit for bit and zero
    the padding columns."""
    rows, Kp = 517, (K + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(K)
    src = torch.randn(rows, K, device="cuda", generator=g)
    src = src if f32 else src.to(BF16)
    dst = torch.full((rows, Kp), 7.0, device="cuda", dtype=BF16)
    lib.call("mvptr_pad_cast", src, int(f32), K, dst, Kp, rows, K)
    assert torch.equal(dst[:, :K], src.to(BF16))
    assert (dst[:, K:] == 0).all()


def test_small_head_dw_row_slices(lib):
    """ITM head weight gradient at the pre-training size (2B = 512 rows): the row-sliced reduction."""
    n, H, C = 512, 768, 2
    x, dl = rnd(n, H, seed=4), torch.randn(n, C, device="cuda") / n
    W = rnd(C, H, scale=0.05, seed=5)
    dx = torch.empty(n, H, device="cuda", dtype=BF16); dW = torch.zeros(C, H, device="cuda"); db = torch.zeros(C, device="cuda")
    lib.call("mvptr_small_head_bwd", dl, x, H, W, dx, H, dW, db, n, H, C)
    assert_close(dW, dl.t() @ x.float(), 1e-4, 1e-5, "small dW (512 rows)")
    assert_close(db, dl.sum(0), 1e-4, 1e-6, "small db (512 rows)")
    assert_close(dx, dl @ W.float(), 1e-2, 1e-5, "small dx (512 rows)")


@pytest.mark.parametrize("M,N", [(515, 3072), (1000, 320)])
def test_gemm_gelu_grad_factor_epilogues(lib, M, N):
    """aux_is_gelu_grad: the FFN1 epilogue saves gelu'(pre-activation) (shared erf evaluation) and the FFN2-dgrad
    epilogue multiplies by the saved factor -- lean tcgen05 epilogues (EPI 4 / 5) and the generic fallback."""
    K = 768
    A, B, bias = rnd(M, K, scale=0.5, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3)
    z = (A.float() @ B.float().t() + bias.float()).requires_grad_(True)
    F.gelu(z).sum().backward()
    gref = z.grad
    out = torch.empty(M, N, device="cuda", dtype=BF16)
    lib.gemm(A, B, out, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="gelu")
    auxs = []
    for kw in (dict(cta_pair=1), dict(cta_pair=2), dict(block_n=128)):  # lean single / lean pair / generic epilogue
        out2 = torch.zeros(M, N, device="cuda", dtype=BF16); aux = torch.zeros_like(out2)
        lib.gemm(A, B, out2, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="gelu", pre_act=aux, ld_aux=N,
                 aux_is_gelu_grad=True, **kw)
        assert_close(aux, gref, 1e-2, 1e-2, f"saved gelu' ({kw})")
        if "cta_pair" in kw:
            assert torch.equal(out2, out), "the activation must not depend on what the second store carries"
        else:
            assert_close(out2, F.gelu(z.detach()), 1e-2, 2e-2, "generic gelu")
        auxs.append(aux)
    assert torch.equal(auxs[0], auxs[1])
    K2 = 256
    dY, W = rnd(M, K2, scale=0.5, seed=5), rnd(K2, N, scale=0.05, seed=6)
    aux = auxs[0]
    ref = (dY.float() @ W.float()) * aux.float()
    dx = torch.zeros(M, N, device="cuda", dtype=BF16)
    cs = torch.full((N,), 0.25, device="cuda", dtype=F32)
    lib.gemm(dY, W, dx, M, N, K2, lda=K2, ldb=N, ldd=N, b_mn=True, gelu_grad_of=aux, ld_aux=N, colsum=cs,
             aux_is_gelu_grad=True)
    assert_close(dx, ref, 1e-2, 1e-2, "acc * saved factor (lean)")
    assert_close(cs, 0.25 + dx.float().sum(0), 1e-3, 2e-2, "fused bias-gradient column sums (factor mode)")
    dx2 = torch.zeros(M, N, device="cuda", dtype=BF16)
    lib.gemm(dY, W, dx2, M, N, K2, lda=K2, ldb=N, ldd=N, b_mn=True, gelu_grad_of=aux, ld_aux=N,
             aux_is_gelu_grad=True, cta_pair=1)  # single-CTA tiles -> generic epilogue
    assert_close(dx2, ref, 1e-2, 1e-2, "acc * saved factor (generic)")
