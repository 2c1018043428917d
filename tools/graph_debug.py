"""Bisect which part of the training step breaks CUDA-graph capture: each stage in its own process."""
import subprocess, sys, traceback
STAGES = ["fwdbwd", "all", "all:same"]
if len(sys.argv) == 1:
    for s in STAGES:
        r = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True)
        print("=====", s, "rc", r.returncode); print(r.stdout[-3000:]); print(r.stderr[-1500:])
    sys.exit(0)
stage, *flags = sys.argv[1].split(":")
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import mvptr_oracle as O
import mvptr_parity_utils as P
from mvp_pytorch_b200 import _lib
from mvp_pytorch_b200.optimization import AdamW
import test_graphs as T
model, opt, batches, Lt = T._setup(0.1)
b = batches[0]
if "noeager" not in flags:
    for _ in range(2):
        T._eager(model, opt, b, Lt)
model.mlm_capacity = (256, 256)
model.mlm_overflow = torch.zeros((), dtype=torch.bool, device="cuda")
opt.enable_graph_mode()
epoch = torch.zeros(1, dtype=torch.int32).pin_memory()
def body():
    if stage in ("epoch", "all"):
        _lib.call("mvptr_set_dropout_epoch", epoch)
    if stage in ("zero", "all"):
        model.zero_grad()
    if stage in ("fwd", "fwdbwd", "all"):
        out = model(max_tag_length=Lt, **b)
        if stage != "fwd":
            out[0].backward()
    if stage in ("opt", "all"):
        opt.step()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    body(); body()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with (torch.cuda.graph(g, stream=side) if "same" in flags else torch.cuda.graph(g)):
        body()
    g.replay(); torch.cuda.synchronize()
    print("OK", stage)
except Exception:
    traceback.print_exc(file=sys.stdout)
    sys.exit(1)
