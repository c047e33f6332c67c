for v in default 8 4 2; do
  if [ "$v" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus 2 --steps 30 --warmup 5 --no-kernel-timing > gpurun_out/nccl_ab_$v.json 2> gpurun_out/nccl_ab_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/nccl_ab_$v.json").read().strip().splitlines()[-1])
print("NCCL_MAX_CTAS=$v", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
P
done
