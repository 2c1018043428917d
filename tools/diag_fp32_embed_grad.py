"""Diagnostic (GPU): where does the fp32 tier's word-embedding gradient differ from the fp32 oracle at the base size?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import mvptr_oracle as O
import mvptr_parity_utils as P
from test_fp32_tier import _run_pretrain32
from test_model_parity import oracle_pretrain

g = torch.load("tests/golden/pretrain_base.pt", weights_only=False)
cfg = O.Cfg()
sd = O.random_state_dict(cfg, "pretrain", seed=g["wseed"])
B, La, Lt, R = g["dims"]
b = O.synthetic_batch(cfg, B, La, Lt, R, seed=g["bseed"], ragged=True, with_labels=True)
model, losses = _run_pretrain32(cfg, sd, b, Lt)
_, g32 = oracle_pretrain(cfg, sd, b, Lt, bf16=False)
params = dict(model.named_parameters())
k = "bert.embeddings.word_embeddings.weight"
got, ref = params[k].grad.float().cpu(), g32[k]
ow = cfg.only_word_size
for name, sl in (("rows [0, only_word)", slice(0, ow)), ("rows [only_word, V)", slice(ow, None)), ("row 0", slice(0, 1)), ("all", slice(None))):
    d = got[sl] - ref[sl]
    print(f"{name}: |ref| {float(ref[sl].norm()):.4e} |got| {float(got[sl].norm()):.4e} rel L2 err {float(d.norm() / (ref[sl].norm() + 1e-30)):.3e}")
touched = torch.unique(torch.cat([b["input_ids_a"].reshape(-1), b["input_ids_b"].reshape(-1)]))
mask = torch.zeros(got.shape[0], dtype=torch.bool); mask[touched] = True
for name, m in (("rows of input tokens", mask), ("other rows", ~mask)):
    d = got[m] - ref[m]
    print(f"{name}: |ref| {float(ref[m].norm()):.4e} rel L2 err {float(d.norm() / (ref[m].norm() + 1e-30)):.3e}")
worst = sorted(((P.rel_l2(params[n].grad, r), n) for n, r in g32.items() if float(r.norm()) > 1e-6), reverse=True)[:8]
print("worst tensors:", worst)
