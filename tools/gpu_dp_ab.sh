#!/bin/bash
# data-parallel A/B of the GEMM SM budget / NCCL CTA cap: usage gpu_dp_ab.sh N "gemm_ctas:nccl_ctas" ...
N=$1; shift
for cfg in "$@"; do
  g=${cfg%%:*}; c=${cfg##*:}
  envs="MVPTR_GEMM_MAX_CTAS=$g"
  [ "$c" != "0" ] && envs="$envs NCCL_MAX_CTAS=$c"
  echo "== gemm_ctas=$g nccl_max_ctas=$c"
  env $envs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 12 --warmup 3 --quick 2>/dev/null | grep quick
done
