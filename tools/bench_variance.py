"""Is the step time stable? Repeat the timed loop with the NVML sampler off / on."""
import sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mvp_pytorch_b200.modeling_vlbert import BiBertImgForPreTraining
from mvp_pytorch_b200.optimization import AdamW
W = bench.WORK
dev = torch.device("cuda")
model = BiBertImgForPreTraining(bench.make_config(0.1)).to(dev).train()
opt = AdamW.for_model(model, lr=1e-4, weight_decay=0.01, max_grad_norm=10.0)
bs = [{k: v.to(dev) for k, v in bench.synthetic_batch(i, 256, W["La"], W["Lt"], W["R"], W["n_phrase"], W["vocab"],
      W["only_word"], W["img_dim"], W["mlm_prob"], torch.bfloat16).items()} for i in range(4)]
def step(i):
    model.zero_grad(); out = model(max_tag_length=W["Lt"], **bs[i % 4]); out[0].backward(); opt.step()
for i in range(8): step(i)
torch.cuda.synchronize()
def timed(n, label):
    st0 = torch.cuda.memory_stats()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = []
    s.record()
    t0 = time.perf_counter()
    for i in range(n):
        t1 = time.perf_counter(); step(i); per.append((time.perf_counter() - t1) * 1e3)
    e.record(); torch.cuda.synchronize()
    st1 = torch.cuda.memory_stats()
    print(f"{label:28s} {s.elapsed_time(e)/n:7.2f} ms/step  host per-step max {max(per):6.1f} min {min(per):6.1f}  "
          f"cudaMalloc +{st1['num_device_alloc']-st0['num_device_alloc']} cudaFree +{st1['num_device_free']-st0['num_device_free']} "
          f"reserved {st1['reserved_bytes.all.current']/2**30:.1f} GiB slow-steps {[i for i,p in enumerate(per) if p > 60]}", flush=True)
timed(15, "no sampler #1"); timed(15, "no sampler #2"); timed(15, "no sampler #3"); timed(15, "no sampler #4")
sm = bench.ClockSampler(0); sm.start(); timed(15, "nvml 50ms #1"); timed(15, "nvml 50ms #2"); print(sm.stop())
gc.disable(); timed(15, "no sampler, gc off #1"); timed(15, "no sampler, gc off #2"); gc.enable()
timed(30, "no sampler 30 steps")
