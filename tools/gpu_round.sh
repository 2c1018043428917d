#!/bin/bash
# One GPU-box visit (gpurun -- 'bash tools/gpu_round.sh'): GPU tests, smoke, the bench line.  Profiles: tools/gpu_profiles.sh.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -25 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
