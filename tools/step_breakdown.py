"""Per-kernel / per-GEMM-shape breakdown of one instrumented pre-training step (library profiler)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mvp_pytorch_b200 import _lib
from mvp_pytorch_b200.modeling_vlbert import BiBertImgForPreTraining
from mvp_pytorch_b200.optimization import AdamW

W = bench.WORK
dev = torch.device("cuda")
model = BiBertImgForPreTraining(bench.make_config(0.1)).to(dev).train()
opt = AdamW.for_model(model, lr=1e-4, weight_decay=0.01, max_grad_norm=10.0)
b = {k: v.to(dev) for k, v in bench.synthetic_batch(0, 256, W["La"], W["Lt"], W["R"], W["n_phrase"], W["vocab"],
                                                    W["only_word"], W["img_dim"], W["mlm_prob"], torch.bfloat16).items()}


def step():
    model.zero_grad()
    out = model(max_tag_length=W["Lt"], **b)
    out[0].backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
_lib.profile_enable(True)
step()
recs = _lib.profile_collect(raw=True)
_lib.profile_enable(False)
agg = collections.OrderedDict()
for name, work, ms in recs:
    key = (name, round(work / 1e6))
    d = agg.setdefault(key, [0, 0.0, work])
    d[0] += 1
    d[1] += ms
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot:.2f} ms over {len(recs)} launches")
for (name, _), (n, ms, work) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    rate = work * n / (ms / 1e3) / 1e12 if ms > 0 else 0
    unit = "TFLOP/s" if name.startswith("gemm") or name.startswith("attn") else "TB/s"
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{n:<3d} {name:18s} work/launch {work/1e9:10.3f} G  -> {rate:8.1f} {unit}")
