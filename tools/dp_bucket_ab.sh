#!/bin/bash
# A/B of the data-parallel bucket size (encoder blocks per all-reduce) at the GPUs given: bash tools/dp_bucket_ab.sh 2
N=${1:-2}
MV_DP_BUCKET_BLOCKS=3 python -m pytest tests/test_gpu_dp.py -x -q 2>&1 | tail -1
for v in 1 3 6; do
  MV_DP_BUCKET_BLOCKS=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus $N --steps 30 --warmup 5 --no-kernel-timing > gpurun_out/dp_bucket_${N}gpu_$v.json 2> gpurun_out/dp_bucket_${N}gpu_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/dp_bucket_${N}gpu_$v.json").read().strip().splitlines()[-1])
print("MV_DP_BUCKET_BLOCKS=$v", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
P
done
