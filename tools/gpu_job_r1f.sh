#!/bin/bash
# Round-1f evidence job (one B200): final bench, ncu launch list, ncu --set full of the hot kernels.
O=gpurun_out
python bench.py > $O/bench28.json 2> $O/bench28.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r1f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-timing > $O/r1f_ncu_list.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:attn_fwd_sn -s 2 -c 1 -o $O/r1f_attn_fwd_sn python tools/prof_attn_one.py > $O/r1f_p4.log 2>&1; echo "p4 rc=$?"
$NCU -k regex:attn_bwd_sn -s 2 -c 1 -o $O/r1f_attn_bwd_sn python tools/prof_attn_one.py bwd > $O/r1f_p5.log 2>&1; echo "p5 rc=$?"
$NCU -k regex:ln_bwd -s 2 -c 1 -o $O/r1f_ln_bwd python tools/prof_elementwise_one.py ln_bwd > $O/r1f_p7.log 2>&1; echo "p7 rc=$?"
$NCU -k regex:gemm2 -s 439 -c 4 -o $O/r1f_gemm_fwd_inmodel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-timing > $O/r1f_a.log 2>&1; echo "gemm fwd rc=$?"
$NCU -k regex:gemm2 -s 487 -c 8 -o $O/r1f_gemm_bwd_inmodel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-timing > $O/r1f_b.log 2>&1; echo "gemm bwd rc=$?"
python tools/probe_quant.py 2>&1 | tail -5 > $O/r1f_quant.log
