#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/gemm_direct_probe.txt
for d in 0 1 0 1; do echo "DIRECT_STORE=$d" | tee -a gpurun_out/gemm_direct_probe.txt; MVPTR_GEMM_DIRECT_STORE=$d python tools/gemm_k768_probe.py 2>&1 | tee -a gpurun_out/gemm_direct_probe.txt; done
MVPTR_GEMM_DIRECT_STORE=1 python -m pytest tests/test_kernels.py tests/test_model_parity.py -q -m gpu --timeout 900 2>&1 | tail -8 | tee gpurun_out/t_direct.log
