#!/bin/bash
# N = 1, 2, 4, 8 back to back on one 8-GPU box (the driver's scaling run, quick form: 30 steps, no side benchmarks)
for N in 1 2 4 8; do
  if [ $N = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 296$((RANDOM%90+10))"; fi
  $L bench.py --gpus $N --steps 30 --warmup 5 --no-kernel-timing --no-cpu-baseline --no-configs --no-quant-bench > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2_scale_$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "MHz", d["clocks"]["sm_mhz"], "exposed_comm_ms", d.get("exposed_comm_ms"))
P
done
