"""GEMM micro-benchmark on the shapes of the pre-training step (CUDA events, L2 flushed between runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_pytorch_b200 import _lib

BF16 = torch.bfloat16
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run(name, M, N, K, reps=5, a_mn=False, b_mn=False, f32=False, **epi):
    A = (torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)).to(BF16)
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).to(BF16)
    D = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else BF16)
    kw = {}
    if epi.get("bias"): kw["bias"] = torch.randn(N, device=dev).to(BF16)
    if epi.get("res"): kw["residual"] = torch.randn(M, N, device=dev).to(BF16); kw["ld_aux"] = N
    if epi.get("pre"): kw["pre_act"] = torch.empty(M, N, device=dev, dtype=BF16); kw["ld_aux"] = N
    if epi.get("gelu"): kw["act"] = "gelu"
    if epi.get("ggrad"): kw["gelu_grad_of"] = torch.randn(M, N, device=dev).to(BF16); kw["ld_aux"] = N
    if epi.get("drop"): kw["p_drop"] = 0.1; kw["seed"] = 5
    if epi.get("split"): kw["split_k"] = epi["split"]; kw["accumulate"] = True
    if epi.get("bn"): kw["block_n"] = epi["bn"]
    kw["cta_pair"] = epi.get("pair", 0)
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _lib.gemm(A, B, D, M, N, K, lda=M if a_mn else K, ldb=N if b_mn else K, ldd=N, a_mn=a_mn, b_mn=b_mn, **kw)
        e.record()
        torch.cuda.synchronize()
        if i >= 2: ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:46s} M={M:6d} N={N:5d} K={K:5d}  {t*1e3:8.1f} us  {2.0*M*N*K/t/1e9:8.1f} TFLOP/s", flush=True)


M = 46080
print("== single-CTA vs CTA-pair tiles ==")
for nm, (m, n, k), kw in (("ffn1 bias", (M, 3072, 768), dict(bias=True)), ("ffn2 bias", (M, 768, 3072), dict(bias=True)),
                          ("o-proj bias", (M, 768, 768), dict(bias=True)), ("qkv bias", (M, 2304, 768), dict(bias=True)),
                          ("txt ffn1 bias", (10240, 3072, 768), dict(bias=True)), ("txt o-proj", (10240, 768, 768), dict(bias=True)),
                          ("dgrad N3072 K768", (M, 3072, 768), dict(b_mn=True)), ("dgrad N768 K3072 +res", (M, 768, 3072), dict(b_mn=True, res=True)),
                          ("wgrad 768x3072", (768, 3072, M), dict(a_mn=True, b_mn=True, f32=True, split=4)),
                          ("wgrad 2304x768", (2304, 768, M), dict(a_mn=True, b_mn=True, f32=True, split=5))):
    run(nm + " [1cta]", m, n, k, pair=1, **kw)
    run(nm + " [pair]", m, n, k, pair=2, **kw)
print("== epilogue ablation, K=768, N=768 ==")
run("plain store", M, 768, 768)
run("plain store bn128", M, 768, 768, bn=128)
run("bias", M, 768, 768, bias=True)
run("bias+res", M, 768, 768, bias=True, res=True)
run("bias+res+drop", M, 768, 768, bias=True, res=True, drop=True)
run("f32 out", M, 768, 768, f32=True)
print("== N=2304 / 3072, K=768 ==")
run("qkv bias", M, 2304, 768, bias=True)
run("ffn1 plain", M, 3072, 768)
run("ffn1 bias+gelu", M, 3072, 768, bias=True, gelu=True)
run("ffn1 bias+gelu+pre", M, 3072, 768, bias=True, gelu=True, pre=True)
print("== K=3072 ==")
run("ffn2 plain", M, 768, 3072)
run("ffn2 bias+res+drop", M, 768, 3072, bias=True, res=True, drop=True)
print("== dgrad (B MN-major) ==")
run("dgrad N=3072 K=768 plain", M, 3072, 768, b_mn=True)
run("dgrad N=3072 K=768 gelu_grad", M, 3072, 768, b_mn=True, ggrad=True)
run("dgrad N=768 K=3072 +res", M, 768, 3072, b_mn=True, res=True)
run("dgrad N=768 K=2304 +res", M, 768, 2304, b_mn=True, res=True)
print("== wgrad (both MN-major, split-K, f32 reduce-add) ==")
run("wgrad 768x3072 K=46080 split4", 768, 3072, M, a_mn=True, b_mn=True, f32=True, split=4)
run("wgrad 3072x768 K=46080 split4", 3072, 768, M, a_mn=True, b_mn=True, f32=True, split=4)
run("wgrad 768x768 K=46080 split16", 768, 768, M, a_mn=True, b_mn=True, f32=True, split=16)
run("wgrad 2304x768 K=46080 split5", 2304, 768, M, a_mn=True, b_mn=True, f32=True, split=5)
print("== big square ==")
run("8192^3", 8192, 8192, 8192)
