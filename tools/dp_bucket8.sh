#!/bin/bash
# bucket-size A/B at N GPUs (quick: 30 steps, no side benchmarks): bash tools/dp_bucket8.sh 8 "6 12"
N=${1:-8}
for v in ${2:-6 12}; do
  MV_DP_BUCKET_BLOCKS=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus $N --steps 30 --warmup 5 --no-kernel-timing --no-cpu-baseline --no-configs --no-quant-bench > gpurun_out/dp_bkt_${N}gpu_$v.json 2> gpurun_out/dp_bkt_${N}gpu_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/dp_bkt_${N}gpu_$v.json").read().strip().splitlines()[-1])
print("MV_DP_BUCKET_BLOCKS=$v", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
P
done
