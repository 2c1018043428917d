#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -8 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
