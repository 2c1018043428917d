"""Attention kernels alone (CUDA events, L2-cold inputs rotated): tcgen05 path vs the mma.sync kernels, forward and
backward, at the step's three shapes.  MVPTR_ATTN_TC is read once per process, so the old kernels are timed in a child."""
import os, subprocess, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvp_pytorch_b200 import _lib


def run(tag):
    out = {}
    H, nh = 768, 12
    for B, L in ((512, 90), (256, 70), (256, 40), (2048, 105)):
        M = B * L
        n_rot = 4
        qkv = [torch.randn(M, 3 * H, device="cuda").to(torch.bfloat16) for _ in range(n_rot)]
        dctx = [torch.randn(M, H, device="cuda").to(torch.bfloat16) * 0.01 for _ in range(n_rot)]
        mask = torch.zeros(B, L, device="cuda")
        ctx = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B, nh, L, device="cuda")
        dqkv = torch.empty(M, 3 * H, device="cuda", dtype=torch.bfloat16)
        dbias = torch.zeros(3 * H, device="cuda")
        for name, fn in (("fwd", lambda i: _lib.call("mvptr_attn_fwd", qkv[i], 3 * H, mask, ctx, H, lse, B, L, nh, H, 0.1, 1234)),
                         ("bwd", lambda i: _lib.call("mvptr_attn_bwd", qkv[i], 3 * H, mask, ctx, dctx[i], H, lse, dqkv, dbias, B, L, nh, H, 0.1, 1234))):
            for i in range(3):
                fn(i % n_rot)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            s.record()
            for i in range(reps):
                fn(i % n_rot)
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) / reps * 1e3
            byts = M * H * 2 * (4 if name == "fwd" else 8)
            out[f"{name} B={B} L={L}"] = {"us": round(us, 1), "GB/s": round(byts / us / 1e3, 0),
                                          "TFLOP/s": round((4 if name == "fwd" else 10) * B * nh * L * L * 64 / us / 1e6, 1)}
    print(tag, json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        run("tcgen05")
        env = dict(os.environ, MVPTR_ATTN_TC="0")
        subprocess.run([sys.executable, __file__, "mma.sync"], env=env)
