#!/bin/bash
# kernel tests, then ncu --set full of the attention and LayerNorm-backward kernels inside one steady-state step
set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels.py -q -m gpu --timeout 900 2>&1 | tail -15
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'attn_|ln_bwd' -c 12 -o gpurun_out/attn_ln python bench.py --profile-step --no-cpu > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log
python tools/step_breakdown.py 2>&1 | tail -50 > gpurun_out/step_breakdown.txt; cat gpurun_out/step_breakdown.txt
