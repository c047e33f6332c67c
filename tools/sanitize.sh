#!/bin/bash
# compute-sanitizer pass over the CUDA path (the reference has no race / memory checker of its own; this is
# ours).  memcheck over the quantiser tests and a subset of the kernel tests, racecheck + synccheck over the elementwise / LayerNorm /
# loss kernels (the tcgen05 / TMA kernels synchronise through mbarriers and the async proxy, which racecheck does not model).
# NOT RUN in this repository's build rounds: the GPU pool they ran on refuses compute-sanitizer ("closed on this pool": earlier runs
# left GPUs needing a reset), so there is no log of it under profiles/ — the script is for a maintainer's own B200.
# usage (GPU box): bash tools/sanitize.sh [out_dir]       -> <out_dir>/sanitize_*.log, one summary line per tool
O=${1:-gpurun_out}
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
run() {   # name, tool, pytest -k expression, files...
  local name=$1 tool=$2 expr=$3; shift 3
  timeout 900 $CS --tool $tool --error-exitcode 99 --launch-timeout 0 python -m pytest "$@" -m gpu -x -q -k "$expr" > $O/sanitize_$name.log 2>&1
  local rc=$?
  echo "$name ($tool): rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitize_$name.log | tail -1)  [$(grep -E 'passed|failed' $O/sanitize_$name.log | tail -1)]"
}
run quant_mem memcheck "not stochastic_is_unbiased" tests/test_gpu_quant.py
run kernels_mem memcheck "layernorm or upsample or patchify or gradient_quantiser_sites or gemm_rejects or (attention and 257)" tests/test_gpu_kernels.py
run elementwise_race racecheck "layernorm or upsample or patchify" tests/test_gpu_kernels.py
run elementwise_sync synccheck "layernorm or upsample or patchify" tests/test_gpu_kernels.py
