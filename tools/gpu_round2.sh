#!/bin/bash
# One GPU-box visit: tests, bench, dropout A/B, in-graph kernel trace, ncu launch list with DRAM bytes.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --quick --steps 10 --warmup 3 --p-drop 0.0 > gpurun_out/bench_p0.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_p0.json
python tools/graph_trace.py > gpurun_out/graph_trace.txt 2>&1; head -45 gpurun_out/graph_trace.txt
if [ "$1" != "noncu" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/gemm_traffic.json > gpurun_out/launches_summary.txt; cat gpurun_out/launches_summary.txt
fi
