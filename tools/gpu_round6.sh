#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/gemm_k768_probe.txt
for d in 0 1 2; do MVPTR_GEMM_DEBUG=$d python tools/gemm_k768_probe.py 2>&1 | tee -a gpurun_out/gemm_k768_probe.txt; done
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -15 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --quick --steps 20 --warmup 3 2>> gpurun_out/bench.err | tee gpurun_out/bench_quick.json
