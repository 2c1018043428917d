#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/gemm_k768_probe2.txt
for d in 0 4 8 2; do MVPTR_GEMM_DEBUG=$d python tools/gemm_k768_probe.py 2>&1 | tee -a gpurun_out/gemm_k768_probe2.txt; done
