"""GPU probe for the tcgen05 GEMM: runs each operand-layout / epilogue variant in its own
subprocess (a hung kernel cannot take the others down) and prints error statistics."""
import json
import subprocess
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    "kk_bf16_256": dict(M=300, N=768, K=768, bn=256),
    "kk_bf16_128": dict(M=300, N=768, K=768, bn=128),
    "kk_small": dict(M=128, N=256, K=64, bn=256),
    "kk_f32": dict(M=300, N=768, K=768, bn=256, f32=True),
    "kk_bias_gelu_pre": dict(M=257, N=3072, K=768, bn=256, bias=True, act="gelu", pre=True),
    "kk_bias_res": dict(M=1000, N=768, K=3072, bn=256, bias=True, res=True),
    "kk_big_persistent": dict(M=20000, N=2304, K=768, bn=256, bias=True),
    "kmn_dgrad": dict(M=500, N=768, K=3072, bn=256, b_mn=True),
    "mnk": dict(M=768, N=768, K=1000, bn=256, a_mn=True),
    "mnmn_wgrad_f32_split": dict(M=768, N=3072, K=5000, bn=256, a_mn=True, b_mn=True, f32=True, split=7, acc=True),
    "mnmn_wgrad_bf16": dict(M=3072, N=768, K=777 * 8, bn=128, a_mn=True, b_mn=True),
    "k2054": dict(M=400, N=768, K=2054, bn=256, kpitch=2056, bias=True),
    "n_odd": dict(M=130, N=1002, K=768, bn=256, npitch=1008, bias=True, f32=True),
    "gelu_grad": dict(M=300, N=3072, K=768, bn=256, b_mn=True, ggrad=True),
    "dropout": dict(M=512, N=768, K=768, bn=256, bias=True, res=True, p_drop=0.1),
    "tanh": dict(M=64, N=768, K=768, bn=128, bias=True, act="tanh", f32=True),
}


def run_case(name):
    import torch
    from mvp_pytorch_b200 import _lib
    c = CASES[name]
    torch.manual_seed(0)
    dev = "cuda"
    M, N, K = c["M"], c["N"], c["K"]
    kp = c.get("kpitch", K)
    npitch = c.get("npitch", N)
    a_mn, b_mn = c.get("a_mn", False), c.get("b_mn", False)
    A_log = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B_log = torch.randn(N, K, device=dev).to(torch.bfloat16)
    if a_mn:
        A_st = A_log.t().contiguous(); lda = M
    else:
        A_st = torch.zeros(M, kp, device=dev, dtype=torch.bfloat16); A_st[:, :K] = A_log; lda = kp
    if b_mn:
        B_st = B_log.t().contiguous(); ldb = N
    else:
        B_st = torch.zeros(N, kp, device=dev, dtype=torch.bfloat16); B_st[:, :K] = B_log; ldb = kp
    odt = torch.float32 if c.get("f32") else torch.bfloat16
    D = torch.full((M, npitch), 7.0, device=dev, dtype=odt)
    ref = A_log.float() @ B_log.float().t()
    kw = {}
    bias = pre = res = gg = None
    if c.get("bias"):
        bias = torch.randn(N, device=dev).to(torch.bfloat16); kw["bias"] = bias
        ref = ref + bias.float()
    if c.get("pre"):
        pre = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); kw["pre_act"] = pre; kw["ld_aux"] = N
    pre_ref = ref.clone()
    if c.get("act") == "gelu":
        ref = torch.nn.functional.gelu(ref); kw["act"] = "gelu"
    if c.get("act") == "tanh":
        ref = torch.tanh(ref); kw["act"] = "tanh"
    if c.get("ggrad"):
        gg = torch.randn(M, N, device=dev).to(torch.bfloat16); kw["gelu_grad_of"] = gg; kw["ld_aux"] = N
        x = gg.float().requires_grad_(True)
        torch.nn.functional.gelu(x).sum().backward()
        ref = ref * x.grad
    if c.get("res"):
        res = torch.randn(M, N, device=dev).to(torch.bfloat16); kw["residual"] = res; kw["ld_aux"] = N
    if c.get("acc"):
        D.fill_(1.0)
        ref = ref + 1.0
    p_drop = c.get("p_drop", 0.0)
    _lib.gemm(A_st, B_st, D, M, N, K, lda=lda, ldb=ldb, ldd=npitch, a_mn=a_mn, b_mn=b_mn,
              accumulate=c.get("acc", False), split_k=c.get("split", 1), block_n=c["bn"], p_drop=p_drop, seed=1234,
              cta_pair=int(os.environ.get("MVPTR_CTA_PAIR", "0")), **kw)
    torch.cuda.synchronize()
    out = D[:, :N].float()
    info = {"case": name}
    if p_drop > 0:
        # out = keep ? ref_pre/0.9 : 0, + res
        core = out - res.float()
        dropped = (core == 0)
        info["drop_frac"] = float(dropped.float().mean())
        kept_err = ((core - ref / (1 - p_drop)).abs() / (ref.abs() / (1 - p_drop) + 1))[~dropped].max().item()
        info["kept_err"] = kept_err
        info["ok"] = bool(abs(info["drop_frac"] - p_drop) < 0.01 and kept_err < 0.02)
    else:
        if res is not None:
            ref = ref + res.float()
        err = (out - ref).abs()
        tol = 0.02 * ref.abs() + 0.05 * (K ** 0.5) * 0.02 + 0.02
        info["max_abs_err"] = float(err.max())
        info["ref_absmax"] = float(ref.abs().max())
        info["frac_bad"] = float((err > tol).float().mean())
        info["ok"] = bool((err <= tol).all())
        if not info["ok"]:
            bad = (err > tol).nonzero()
            info["first_bad"] = bad[:5].tolist()
            info["bad_rows_mod128"] = sorted(set((bad[:2000, 0] % 128).tolist()))[:20]
            info["bad_cols_mod64"] = sorted(set((bad[:2000, 1] % 64).tolist()))[:20]
        if pre is not None:
            perr = (pre.float() - pre_ref).abs().max().item()
            info["pre_err"] = perr
            info["ok"] = info["ok"] and perr < 0.02 * pre_ref.abs().max().item() + 0.05
        if npitch != N:
            info["pad_untouched"] = bool((D[:, N:] == 7.0).all())
            info["ok"] = info["ok"] and info["pad_untouched"]
    print("PROBE " + json.dumps(info), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "all":
        run_case(sys.argv[1])
    else:
        n_bad = 0
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120)
                lines = [l for l in r.stdout.splitlines() if l.startswith("PROBE")]
                if lines:
                    print(lines[-1])
                    n_bad += 0 if json.loads(lines[-1][6:])["ok"] else 1
                else:
                    n_bad += 1
                    print("PROBE-FAIL", name, "rc", r.returncode, (r.stdout + r.stderr)[-600:].replace("\n", " | "))
            except subprocess.TimeoutExpired:
                n_bad += 1
                print("PROBE-TIMEOUT", name)
        print("PROBE-SUMMARY bad =", n_bad)
