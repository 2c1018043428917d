#!/bin/bash
# GPU visit: tests, bench, A/B of the gelu'-factor scheme, in-graph kernel trace.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
rm -f gpurun_out/bench_ab.json
for v in 1 0 1 0; do
  echo "MVPTR_GELU_GRAD_FACTOR=$v" | tee -a gpurun_out/bench_ab.json
  MVPTR_GELU_GRAD_FACTOR=$v python bench.py --quick --steps 20 --warmup 3 2>> gpurun_out/bench.err | tee -a gpurun_out/bench_ab.json
done
python tools/graph_trace.py > gpurun_out/graph_trace.txt 2>&1; head -34 gpurun_out/graph_trace.txt
