#!/bin/bash
# One GPU-box visit: tests, bench, ncu launch list of one steady-state step, full ncu capture of the GEMM.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; cat gpurun_out/launches_summary.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s 20 -c 4 -o gpurun_out/gemm_full python bench.py --profile-step --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
fi
