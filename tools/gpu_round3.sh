#!/bin/bash
# GPU visit: tests, bench, end-to-end loop A/B, ncu --set full of the row kernels + attention at the cross-modal shape.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
for mode in "" noh2d nod2h; do
  python bench.py --quick --steps 10 --warmup 3 --e2e-debug "$mode" 2>> gpurun_out/bench.err | tee -a gpurun_out/bench_e2e_ab.json
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd|ln_bwd|attn_fwd|attn_bwd" -s 4 -c 4 -f -o gpurun_out/rowk2 python tools/rowkernels_one.py > gpurun_out/ncu_rowk2.log 2>&1; tail -2 gpurun_out/ncu_rowk2.log
