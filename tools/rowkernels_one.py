"""Launch each non-GEMM kernel once at the cross-modal encoder shape (M = 512 x 90 tokens) -- for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_pytorch_b200 import _lib
BF16, F32 = torch.bfloat16, torch.float32
B, L, nh, H, I = 512, 90, 12, 768, 3072
M = B * L
dev = "cuda"
x = torch.randn(M, H, device=dev).to(BF16); res = torch.randn(M, H, device=dev).to(BF16)
g = torch.ones(H, device=dev, dtype=BF16); b = torch.zeros(H, device=dev, dtype=BF16)
y = torch.empty_like(x); pre = torch.empty_like(x); st = torch.empty(2, M, device=dev, dtype=F32)
dx = torch.empty_like(x); dxd = torch.empty_like(x)
dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev); dbias = torch.zeros(H, device=dev)
big = torch.randn(M, I, device=dev).to(BF16); big2 = torch.empty_like(big); big3 = torch.randn(M, I, device=dev).to(BF16)
dbi = torch.zeros(I, device=dev)
qkv = torch.randn(M, 3 * H, device=dev).to(BF16); ctx = torch.empty(M, H, device=dev, dtype=BF16)
lse = torch.empty(B, nh, L, device=dev, dtype=F32); mask = torch.zeros(B, L, device=dev)
dctx = torch.randn(M, H, device=dev).to(BF16); dqkv = torch.empty_like(qkv); dbq = torch.zeros(3 * H, device=dev)
for it in range(2):
    _lib.call("mvptr_add_ln_fwd", x, res, 0.1, 7, pre, g, b, y, st[0], st[1], M, H, 1e-12)
    _lib.call("mvptr_ln_bwd", y, 0, 0, pre, st[0], st[1], g, dx, dxd, dg, db, dbias, M, H, 0.0, 0, 0.1, 7)
    _lib.call("mvptr_gelu_fwd", big, big2, M * I)
    _lib.call("mvptr_gelu_bwd_colsum", big3, big, big2, dbi, M, I)
    _lib.call("mvptr_attn_fwd", qkv, 3 * H, mask, ctx, H, lse, B, L, nh, H, 0.1, 5)
    _lib.call("mvptr_attn_bwd", qkv, 3 * H, mask, ctx, dctx, H, lse, dqkv, dbq, B, L, nh, H, 0.1, 5)
    _lib.call("mvptr_colsum", dqkv, 3 * H, dbq, M, 3 * H)
torch.cuda.synchronize()
