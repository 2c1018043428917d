"""One GEMM shape, a few launches (for ncu --set full)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_pytorch_b200 import _lib
M, N, K = (int(x) for x in sys.argv[1:4])
b_mn = "bmn" in sys.argv[4:]
act = "gelu" if "gelu" in sys.argv[4:] else None
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
B = (torch.randn(K, N, device="cuda") if b_mn else torch.randn(N, K, device="cuda")).to(torch.bfloat16)
D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(N, device="cuda").to(torch.bfloat16)
for _ in range(4):
    _lib.gemm(A, B, D, M, N, K, lda=K, ldb=N if b_mn else K, ldd=N, b_mn=b_mn, bias=bias, act=act)
torch.cuda.synchronize()
