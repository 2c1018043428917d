"""Pure host cost of one training step: run it at batch 4 (GPU work negligible) and time the phases."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mvp_pytorch_b200.modeling_vlbert import BiBertImgForPreTraining
from mvp_pytorch_b200.optimization import AdamW

W = bench.WORK
dev = torch.device("cuda")
model = BiBertImgForPreTraining(bench.make_config(0.1)).to(dev).train()
opt = AdamW.for_model(model, lr=1e-4, weight_decay=0.01, max_grad_norm=10.0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
b = {k: v.to(dev) for k, v in bench.synthetic_batch(0, B, W["La"], W["Lt"], W["R"], W["n_phrase"], W["vocab"],
                                                    W["only_word"], W["img_dim"], W["mlm_prob"], torch.bfloat16).items()}


def step(t=None):
    t0 = time.perf_counter()
    model.zero_grad()
    t1 = time.perf_counter()
    out = model(max_tag_length=W["Lt"], **b)
    t2 = time.perf_counter()
    out[0].backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    if t is not None:
        for i, d in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            t[i] += d * 1e3


for _ in range(5):
    step()
torch.cuda.synchronize()
acc = [0, 0, 0, 0]
n = 20
t0 = time.perf_counter()
for _ in range(n):
    step(acc)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / n
print("B=%d host ms/step: zero_grad %.2f fwd %.2f bwd %.2f opt %.2f | wall %.2f" % tuple([B] + [a / n for a in acc] + [wall]))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:6000])
