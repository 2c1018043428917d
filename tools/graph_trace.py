"""Per-kernel durations INSIDE CUDA-graph replays of the pre-training step (torch.profiler / CUPTI):
busy time vs step time (= launch gaps), per-kernel totals, GEMM durations by launch order."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from mvp_pytorch_b200.modeling_vlbert import BiBertImgForPreTraining
from mvp_pytorch_b200.optimization import AdamW
from mvp_pytorch_b200.graphs import GraphedTrainStep

W = bench.WORK
dev = torch.device("cuda")
model = BiBertImgForPreTraining(bench.make_config(0.1)).to(dev).train()
opt = AdamW.for_model(model, lr=1e-4, weight_decay=0.01, max_grad_norm=10.0)
b = {k: v.to(dev) for k, v in bench.synthetic_batch(0, 256, W["La"], W["Lt"], W["R"], W["n_phrase"], W["vocab"],
                                                    W["only_word"], W["img_dim"], W["mlm_prob"], torch.bfloat16).items()}
step = GraphedTrainStep(model, opt, b, forward_kwargs=dict(max_tag_length=W["Lt"]))
for _ in range(5):
    step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"{len(evs)} device activities over {N} replays; span {(t1-t0)/N/1e3:.3f} ms/step, busy {busy/N/1e3:.3f} ms/step, "
      f"gaps {(t1-t0-busy)/N/1e3:.3f} ms/step")
agg = collections.OrderedDict()
for e in evs:
    n = e.name.split("(")[0][-70:]
    d = agg.setdefault(n, [0, 0.0])
    d[0] += 1; d[1] += (e.time_range.end - e.time_range.start)
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{us/N/1e3:8.3f} ms  x{c//N:<4d} {n}")
# gaps by preceding kernel
gaps = collections.Counter()
for a, bb in zip(evs, evs[1:]):
    g = bb.time_range.start - a.time_range.end
    if g > 0:
        gaps[a.name.split("(")[0][-40:]] += g
print("largest gap totals (us/step) by preceding kernel:")
for n, g in gaps.most_common(8):
    print(f"  {g/N:8.1f}  {n}")
# the largest individual gaps with their neighbours (what the device was waiting for)
big = []
for a, bb in zip(evs, evs[1:]):
    g = bb.time_range.start - a.time_range.end
    if g > 15:
        big.append((g, a.name.split("(")[0][-60:], bb.name.split("(")[0][-60:]))
print(f"{len(big)} gaps > 15 us over {N} replays:")
for g, a, bb in sorted(big, reverse=True)[:12]:
    print(f"  {g:8.1f} us   after [{a}]   before [{bb}]")
