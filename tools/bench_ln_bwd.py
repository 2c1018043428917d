"""LayerNorm backward microbench (CUDA events, 40 launches after 5 warm-ups, buffers > L2 in rotation):
    python tools/bench_ln_bwd.py [path/to/alternative/libmvptr_b200.so]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_pytorch_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
BF16, F32 = torch.bfloat16, torch.float32
dev = "cuda"
H = 768
for M in (46080, 17920, 15360):
    nb = 3  # rotate 3 buffer sets: 3 x 4 x 70 MB > 126 MB L2
    sets = []
    for i in range(nb):
        dy = torch.randn(M, H, device=dev).to(BF16); x = torch.randn(M, H, device=dev).to(BF16)
        sets.append((dy, x, torch.empty_like(x), torch.empty_like(x)))
    mean = torch.zeros(M, device=dev); rstd = torch.ones(M, device=dev)
    g = torch.ones(H, device=dev, dtype=BF16)
    dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev); dbias = torch.zeros(H, device=dev)
    for name, out_p, in_p, want_drop in (("dropout 0.1 (dx + dx_drop + dbias)", 0.1, 0.1, True), ("no dropout (dx + dbias)", 0.0, 0.0, False)):
        def run(i):
            dy, x, dx, dxd = sets[i % nb]
            _lib.call("mvptr_ln_bwd", dy, 0, 0, x, mean, rstd, g, dx, dxd if want_drop else None, dg, db, dbias, M, H,
                      out_p, 3, in_p, 7)
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(40):
            run(i)
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) / 40 * 1e3
        nbytes = M * H * 2 * (4 if want_drop else 3)
        print(f"M={M:6d} {name:36s} {us:7.1f} us  {nbytes / us / 1e6:6.2f} TB/s algorithmic")
