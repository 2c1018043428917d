TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
Q="bench.py --gpus 8 --steps 15 --warmup 4 --quick"
timeout 200 $TR --master-port 29511 $Q > gpurun_out/n8q_default.json 2> gpurun_out/n8q_default.err
timeout 200 $TR --master-port 29512 tools/dp_trace.py > gpurun_out/r2_dp_trace_n8.txt 2>/dev/null
NCCL_DEBUG=WARN NCCL_MAX_CTAS=8 timeout 200 $TR --master-port 29513 $Q > gpurun_out/n8q_maxctas8.json 2>/dev/null
NCCL_DEBUG=WARN NCCL_MAX_CTAS=16 timeout 200 $TR --master-port 29514 $Q > gpurun_out/n8q_maxctas16.json 2>/dev/null
NCCL_DEBUG=WARN NCCL_ALGO=Ring timeout 200 $TR --master-port 29515 $Q > gpurun_out/n8q_ring.json 2>/dev/null
NCCL_DEBUG=WARN NCCL_NVLS_ENABLE=0 timeout 200 $TR --master-port 29516 $Q > gpurun_out/n8q_nonvls.json 2>/dev/null
NCCL_DEBUG=WARN MVPTR_DP_MIN_BUCKET=8388608 timeout 200 $TR --master-port 29517 $Q > gpurun_out/n8q_bucket8m.json 2>/dev/null
