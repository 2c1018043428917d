#!/bin/bash
# usage: gpu_ncu_kernel.sh <kernel-regex> <out-name> [count] [skip]: ncu --set full of matching kernels inside one steady-state step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$1" -s ${4:-0} -c ${3:-1} -f -o gpurun_out/$2 python bench.py --profile-step --no-cpu > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log
