#!/bin/bash
# GPU visit: full GPU test suite, then the bench with and without the CUDA-graph step.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --steps 10 --warmup 3 --no-graph --no-cpu > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; tail -3 gpurun_out/bench_nograph.err; cat gpurun_out/bench_nograph.json
