#!/bin/bash
# A/B of the 128-row TMA store boxes of the generic GEMM epilogue (MVPTR_GEMM_WIDE_STORE=1 default / 0)
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -12 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
rm -f gpurun_out/wide_store_ab.txt
for v in 1 0; do echo "WIDE_STORE=$v" | tee -a gpurun_out/wide_store_ab.txt; MVPTR_GEMM_WIDE_STORE=$v python tools/gemm_k768_probe.py 2>&1 | tee -a gpurun_out/wide_store_ab.txt; done
for v in 1 0 1 0; do echo "WIDE_STORE=$v" | tee -a gpurun_out/wide_store_ab.txt; MVPTR_GEMM_WIDE_STORE=$v python bench.py --quick --steps 20 --warmup 3 2>> gpurun_out/bench.err | tee -a gpurun_out/wide_store_ab.txt; done
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
