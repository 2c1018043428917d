TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
Q="bench.py --gpus 2 --steps 15 --warmup 4 --quick"
i=0
for v in "tail-bf16:65536" "tail-bf16:1073741824" "fp32:1073741824" "tail-bf16:33554432" "tail-bf16:65536"; do
  i=$((i+1)); rd=${v%%:*}; mb=${v##*:}
  MVPTR_DP_REDUCE=$rd MVPTR_DP_MIN_BUCKET=$mb timeout 200 $TR --master-port $((29550+i)) $Q > gpurun_out/n2ab_$i.json 2>/dev/null
  echo "$v $(cat gpurun_out/n2ab_$i.json)" >> gpurun_out/n2ab.txt
done
timeout 200 python bench.py --steps 15 --warmup 4 --quick >> gpurun_out/n2ab.txt 2>/dev/null
