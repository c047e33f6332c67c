#!/bin/bash
# A/B of the SM limit for the persistent kernels launched behind a gradient all-reduce: bash tools/dp_smlimit_ab.sh N
# (MV_DP_SM_LIMIT=148 switches it off).  Each line: img/s, ms per step, end-to-end img/s, SM MHz.
N=${1:-2}
[ -z "$MV_AB_SKIP_TEST" ] && python -m pytest tests/test_gpu_dp.py -x -q 2>&1 | tail -1
for cfg in ${MV_AB_CONFIGS:-148:0 132:4 116:4 116:8 100:4 148:0}; do
  set -- ${cfg/:/ }
  MV_DP_SM_LIMIT=$1 MV_DP_LIMIT_LAUNCHES=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) bench.py --gpus $N --steps 30 --warmup 5 --no-kernel-timing --no-cpu-baseline --no-configs --no-quant-bench > gpurun_out/dp_sml_${N}gpu_$1_$2.json 2> gpurun_out/dp_sml_${N}gpu_$1_$2.err
  python - <<P
import json
d=json.loads(open("gpurun_out/dp_sml_${N}gpu_$1_$2.json").read().strip().splitlines()[-1])
print("MV_DP_SM_LIMIT=$1 launches=$2", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
P
done
python bench.py --steps 30 --warmup 5 --no-kernel-timing --no-cpu-baseline --no-configs --no-quant-bench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
