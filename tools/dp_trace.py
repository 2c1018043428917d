"""Where a data-parallel step loses time (rank 0's CUPTI timeline of CUDA-graph replays, torch.profiler):
step span, compute-kernel busy time, NCCL kernel time, the part of it that is EXPOSED (no compute kernel running),
and per-kernel totals to compare with the 1-GPU trace (tools/graph_trace.py).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/dp_trace.py
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity

import bench
from mvp_pytorch_b200.modeling_vlbert import BiBertImgForPreTraining
from mvp_pytorch_b200.optimization import AdamW
from mvp_pytorch_b200.graphs import GraphedTrainStep

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
W = bench.WORK
torch.manual_seed(1234 + rank)
model = BiBertImgForPreTraining(bench.make_config(0.1)).to(dev).train()
if world > 1:
    dist.broadcast(model.runtime().arena.master, 0)
opt = AdamW.for_model(model, lr=1e-4, weight_decay=0.01, max_grad_norm=10.0)
if world > 1:
    from mvp_pytorch_b200.parallel import enable_overlapped_allreduce
    rd = {"fp32": torch.float32, "bf16": torch.bfloat16}.get(os.environ.get("MVPTR_DP_REDUCE", "tail-bf16"), "tail-bf16")
    enable_overlapped_allreduce(model, reduce_dtype=rd, min_bucket=int(os.environ.get("MVPTR_DP_MIN_BUCKET", str(1 << 16))))
b = {k: v.to(dev) for k, v in bench.synthetic_batch(100 * rank, 256, W["La"], W["Lt"], W["R"], W["n_phrase"], W["vocab"],
                                                    W["only_word"], W["img_dim"], W["mlm_prob"], torch.float32).items()}


def eager():
    model.zero_grad()
    out = model(max_tag_length=W["Lt"], **b)
    out[0].backward()
    if world > 1:
        from mvp_pytorch_b200.parallel import allreduce_gradients
        allreduce_gradients(model)
    opt.step()


for _ in range(3):
    eager()
step = GraphedTrainStep(model, opt, b, forward_kwargs=dict(max_tag_length=W["Lt"]), allreduce=world > 1)
for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    is_nccl = lambda e: "nccl" in e.name.lower()
    comp = [(e.time_range.start, e.time_range.end) for e in evs if not is_nccl(e)]
    comm = [(e.time_range.start, e.time_range.end, e.name) for e in evs if is_nccl(e)]
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)

    def union(iv):
        out = []
        for s, e in sorted(iv):
            if out and s <= out[-1][1]:
                out[-1][1] = max(out[-1][1], e)
            else:
                out.append([s, e])
        return out

    cu = union(comp)
    busy = sum(e - s for s, e in cu)
    nccl_total = sum(e - s for s, e, _ in comm)
    exposed = 0.0
    for s, e, _ in comm:  # NCCL time not covered by any compute kernel
        cov = sum(max(0, min(e, ce) - max(s, cs)) for cs, ce in cu if ce > s and cs < e)
        exposed += (e - s) - cov
    print(f"world {world}: {len(evs)} device activities over {N} replays; span {(t1 - t0) / N / 1e3:.3f} ms/step; compute "
          f"kernels busy (union) {busy / N / 1e3:.3f} ms/step; NCCL kernels {len(comm) // N} per step, {nccl_total / N / 1e3:.3f} "
          f"ms/step of which EXPOSED (no compute kernel running) {exposed / N / 1e3:.3f} ms/step")
    agg = collections.OrderedDict()
    for e in evs:
        n = e.name.split("(")[0][-70:]
        d = agg.setdefault(n, [0, 0.0])
        d[0] += 1
        d[1] += (e.time_range.end - e.time_range.start)
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print(f"{us / N / 1e3:8.3f} ms  x{c // N:<4d} {n}")
    # the last replay: when does backward end, when does the last collective end, when does AdamW start
    last = [e for e in evs if e.time_range.start >= evs[0].time_range.start + (t1 - t0) * (N - 1) / N]
    adam = [e for e in last if "adamw" in e.name]
    nc = [e for e in last if is_nccl(e)]
    if adam and nc:
        a0 = adam[0].time_range.start
        before = [e for e in last if not is_nccl(e) and e.time_range.end <= a0 and "sumsq" not in e.name and "adamw" not in e.name]
        print(f"last replay: last backward kernel ends {(before[-1].time_range.end - last[0].time_range.start) / 1e3:.3f} ms, last "
              f"collective ends {(max(e.time_range.end for e in nc) - last[0].time_range.start) / 1e3:.3f} ms, AdamW starts "
              f"{(a0 - last[0].time_range.start) / 1e3:.3f} ms into the step")
        for e in nc[-6:]:
            print(f"   nccl {e.name[:50]} start {(e.time_range.start - last[0].time_range.start) / 1e3:.3f} dur {(e.time_range.end - e.time_range.start) / 1e3:.3f} ms")
if world > 1:
    torch.cuda.synchronize()
    step.release()
    bench._shutdown(world)
