#!/bin/bash
# GPU visit: smoke, in-graph trace with gap detail, full-size config-3 retrieval, ncu launch list (+DRAM bytes),
# ncu --set full of the cross-modal encoder's forward GEMMs.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python tools/graph_trace.py > gpurun_out/graph_trace.txt 2>&1; tail -22 gpurun_out/graph_trace.txt
python tools/retrieval_c3.py > gpurun_out/retrieval_c3.json 2> gpurun_out/retrieval_c3.err; tail -2 gpurun_out/retrieval_c3.err; cat gpurun_out/retrieval_c3.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/gemm_traffic.json > gpurun_out/launches_summary.txt; head -20 gpurun_out/launches_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s 60 -c 8 -f -o gpurun_out/gemm_mul python bench.py --profile-step --no-cpu > gpurun_out/ncu_gemm_mul.log 2>&1; tail -2 gpurun_out/ncu_gemm_mul.log
