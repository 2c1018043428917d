"""Where does a K=768 forward GEMM tile spend its 11-12 k cycles?  Times the QKV / out-projection / FFN1 shapes
of the cross-modal encoder with single-CTA and CTA-pair tiles; run once per MVPTR_GEMM_DEBUG setting
(0 = real kernel, 1 = no slab-reuse wait, 2 = no epilogue at all: mainloop + TMA only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_pytorch_b200 import _lib
BF16 = torch.bfloat16
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
M = 46080
dbg = os.environ.get("MVPTR_GEMM_DEBUG", "0")
for name, N, K in (("qkv", 2304, 768), ("o-proj", 768, 768), ("ffn1", 3072, 768), ("ffn2", 768, 3072)):
    A = torch.randn(M, K, device=dev).to(BF16); B = torch.randn(N, K, device=dev).to(BF16)
    D = torch.zeros(M, N, device=dev, dtype=BF16); bias = torch.randn(N, device=dev).to(BF16)
    for pair in (1, 2):
        ts = []
        for i in range(7):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.gemm(A, B, D, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, cta_pair=pair)
            e.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(s.elapsed_time(e))
        t = sorted(ts)[len(ts) // 2]
        # back to back (no flush, no idle gaps): the in-step regime
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(20):
            _lib.gemm(A, B, D, M, N, K, lda=K, ldb=K, ldd=N, bias=bias, cta_pair=pair)
        e.record(); torch.cuda.synchronize()
        tb = s.elapsed_time(e) / 20
        print(f"debug={dbg} {name:7s} {'pair' if pair == 2 else '1cta'}  isolated {t*1e3:7.1f} us {2.0*M*N*K/t/1e9:7.1f} TF/s"
              f"   back-to-back {tb*1e3:7.1f} us {2.0*M*N*K/tb/1e9:7.1f} TF/s", flush=True)
