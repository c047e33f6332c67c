#!/bin/bash
# NCCL CTA budget A/B at N GPUs (quick form): bash tools/nccl_ctas_ab8.sh 8 "default 16 8"
N=${1:-8}
for v in ${2:-default 16 8}; do
  if [ "$v" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  NCCL_DEBUG=${NCCL_DEBUG_AB:-WARN} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus $N --steps 30 --warmup 5 --no-kernel-timing --no-cpu-baseline --no-configs --no-quant-bench > gpurun_out/nccl_ab${N}_$v.json 2> gpurun_out/nccl_ab${N}_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/nccl_ab${N}_$v.json").read().strip().splitlines()[-1])
print("N=$N NCCL_MAX_CTAS=$v", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
P
done
