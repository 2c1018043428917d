#!/bin/bash
# quick perf check: bench twice (variance) without tests
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RUN1 ms/step',d['ms_per_step'],'host',d['host_enqueue_ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'gemm',d['roofline']['gemm_ms_per_step'])
for k in d['roofline']['top_kernels_ms']: print(k)
"
python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RUN2 ms/step',d['ms_per_step'],'host',d['host_enqueue_ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'gemm',d['roofline']['gemm_ms_per_step'], d['clocks'])
"
python - <<'PY'
import time, torch, sys
sys.path.insert(0,'.')
import bench
from mvp_pytorch_b200.modeling_vlbert import BiBertImgForPreTraining
from mvp_pytorch_b200.optimization import AdamW
W=bench.WORK
dev=torch.device('cuda')
model=BiBertImgForPreTraining(bench.make_config(0.1)).to(dev).train()
opt=AdamW.for_model(model, lr=1e-4, weight_decay=0.01, max_grad_norm=10.0)
b={k:v.to(dev) for k,v in bench.synthetic_batch(0,256,W['La'],W['Lt'],W['R'],W['n_phrase'],W['vocab'],W['only_word'],W['img_dim'],W['mlm_prob'],torch.bfloat16).items()}
def step(prof=None):
    t=[time.perf_counter()]
    model.zero_grad(); t.append(time.perf_counter())
    out=model(max_tag_length=W['Lt'],**b); t.append(time.perf_counter())
    out[0].backward(); t.append(time.perf_counter())
    opt.step(); t.append(time.perf_counter())
    return [ (t[i+1]-t[i])*1e3 for i in range(4)]
for i in range(4): step()
torch.cuda.synchronize()
acc=[0,0,0,0]
t0=time.perf_counter()
for i in range(10):
    r=step()
    acc=[a+x for a,x in zip(acc,r)]
th=time.perf_counter()-t0
torch.cuda.synchronize()
tt=time.perf_counter()-t0
print('HOST per step ms: zero_grad %.2f fwd %.2f bwd %.2f opt %.2f | host total %.2f wall %.2f'%tuple([a/10 for a in acc]+[th*100, tt*100]))
PY
