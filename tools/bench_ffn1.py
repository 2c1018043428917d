"""FFN1 (bias+GELU) variants: plain GEMM + separate GELU kernel vs fused epilogue (single / dual output)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_pytorch_b200 import _lib
BF16 = torch.bfloat16
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timeit(fn, reps=7):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2] * 1e3

for M in (46080, 17920, 10240):
    N, K = 3072, 768
    A = torch.randn(M, K, device=dev).to(BF16); B = torch.randn(N, K, device=dev).to(BF16) * 0.05
    bias = torch.randn(N, device=dev).to(BF16)
    pre = torch.empty(M, N, device=dev, dtype=BF16); act = torch.empty(M, N, device=dev, dtype=BF16)
    def plain(pair): return lambda: _lib.gemm(A, B, pre, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, cta_pair=pair)
    def gelu_only(pair): return lambda: _lib.gemm(A, B, act, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="gelu", cta_pair=pair)
    def dual(pair): return lambda: _lib.gemm(A, B, act, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, act="gelu", pre_act=pre, ld_aux=N, cta_pair=pair)
    def sep():
        _lib.gemm(A, B, pre, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, cta_pair=1)
        _lib.call("mvptr_gelu_fwd", pre, act, M * N)
    print(f"M={M}: plain 1cta {timeit(plain(1)):7.1f}  pair {timeit(plain(2)):7.1f} | gemm+gelu kernel {timeit(sep):7.1f} | "
          f"fused gelu 1cta {timeit(gelu_only(1)):7.1f} pair {timeit(gelu_only(2)):7.1f} | dual 1cta {timeit(dual(1)):7.1f} pair {timeit(dual(2)):7.1f} us")
    # correctness of dual vs separate
    sep(); ref_pre, ref_act = pre.clone(), act.clone()
    pre.zero_(); act.zero_(); dual(2)(); torch.cuda.synchronize()
    print("   dual pair: pre equal", torch.equal(pre, ref_pre), " act max diff", float((act.float() - ref_act.float()).abs().max()))
    pre.zero_(); act.zero_(); dual(1)(); torch.cuda.synchronize()
    print("   dual 1cta: pre equal", torch.equal(pre, ref_pre), " act max diff", float((act.float() - ref_act.float()).abs().max()))
