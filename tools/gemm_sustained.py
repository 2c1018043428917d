"""Is the in-step GEMM rate a power/clock effect?  Same GEMM (a) isolated with idle gaps and an L2
flush, (b) back to back for ~1.5 s on rotating cold buffers, with NVML clock/power samples."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from mvp_pytorch_b200 import _lib

BF16 = torch.bfloat16
dev = "cuda"
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def sample(stop, out):
    while not stop[0]:
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.02)


def run(name, M, N, K, b_mn=False, pair=0):
    NB = 4
    As = [torch.randn(M, K, device=dev).to(BF16) for _ in range(NB)]
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).to(BF16)
    Ds = [torch.zeros(M, N, device=dev, dtype=BF16) for _ in range(NB)]
    bias = torch.randn(N, device=dev).to(BF16)

    def go(i):
        _lib.gemm(As[i % NB], B, Ds[i % NB], M, N, K, lda=K, ldb=N if b_mn else K, ldd=N, b_mn=b_mn, bias=bias, cta_pair=pair)
    ts = []
    for i in range(6):
        flush.zero_(); torch.cuda.synchronize(); time.sleep(0.05)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); go(i); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    iso = sorted(ts[1:])[len(ts[1:]) // 2]
    n = int(1500.0 / iso)
    stop, smp = [False], []
    th = threading.Thread(target=sample, args=(stop, smp)); th.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        go(i)
    e.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    sus = s.elapsed_time(e) / n
    clk = sorted(c for c, _ in smp)[len(smp) // 2] if smp else 0
    pw = max(p for _, p in smp) if smp else 0
    f = 2.0 * M * N * K
    print(f"{name:28s} isolated {iso*1e3:7.1f} us {f/iso/1e9:7.1f} TF/s | back-to-back x{n:5d} {sus*1e3:7.1f} us {f/sus/1e9:7.1f} TF/s"
          f"  sm {clk} MHz  power max {pw:.0f} W", flush=True)


M = 46080
run("qkv 1cta", M, 2304, 768, pair=1)
run("qkv pair", M, 2304, 768, pair=2)
run("ffn1 1cta", M, 3072, 768, pair=1)
run("ffn2 pair", M, 768, 3072, pair=2)
run("o-proj 1cta", M, 768, 768, pair=1)
run("dgrad ffn1 pair", M, 3072, 768, b_mn=True, pair=2)
run("txt ffn1 1cta", 10240, 3072, 768, pair=1)
# memory-bound reference: device copy back to back
x = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev); y = torch.empty_like(x)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): y.copy_(x)
s.record()
for _ in range(50): y.copy_(x)
e.record(); torch.cuda.synchronize()
print(f"copy 512 MiB: {2*x.numel()*50/s.elapsed_time(e)/1e9:.2f} TB/s")
