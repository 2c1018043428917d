#!/bin/bash
# 8-GPU visit: A/B the gradient all-reduce variants with the quick resident-step timer, then the full bench line with the best one.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
Q="bench.py --gpus 8 --steps 15 --warmup 4 --quick"
mkdir -p gpurun_out
best=""; best_ms=1000000
i=0
for v in "tail-bf16:65536" "bf16:8388608" "tail-bf16:1073741824"; do
  i=$((i+1))
  rd=${v%%:*}; mb=${v##*:}
  MVPTR_DP_REDUCE=$rd MVPTR_DP_MIN_BUCKET=$mb timeout 200 $TR --master-port $((29520+i)) $Q > gpurun_out/n8f_${rd}_${mb}.json 2>/dev/null
  ms=$(python -c "import json,sys; print(json.loads([l for l in open('gpurun_out/n8f_${rd}_${mb}.json') if l.startswith('{')][-1])['ms_per_step'])" 2>/dev/null || echo 1000000)
  echo "$v $ms" >> gpurun_out/n8f_choice.txt
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$best_ms') else 1)"; then best=$v; best_ms=$ms; fi
done
echo "best $best $best_ms" >> gpurun_out/n8f_choice.txt
MVPTR_DP_REDUCE=${best%%:*} MVPTR_DP_MIN_BUCKET=${best##*:} timeout 400 $TR --master-port 29530 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
timeout 200 python bench.py --steps 15 --warmup 4 --quick > gpurun_out/r2_bench_n1_on_n8box_final.json 2>/dev/null
