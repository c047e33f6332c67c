#!/bin/bash
# Round-1e evidence job (one B200): bench, ncu launch list, ncu --set full of the attention kernels after the store / issuer work.
O=gpurun_out
python bench.py > $O/bench19.json 2> $O/bench19.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r1e_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-timing > $O/r1e_ncu_list.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:attn_fwd_sn -s 2 -c 1 -o $O/r1e_attn_fwd_sn python tools/prof_attn_one.py > $O/r1e_p4.log 2>&1; echo "p4 rc=$?"
$NCU -k regex:attn_bwd_sn -s 2 -c 1 -o $O/r1e_attn_bwd_sn python tools/prof_attn_one.py bwd > $O/r1e_p5.log 2>&1; echo "p5 rc=$?"
$NCU -k regex:ln_fwd -s 2 -c 1 -o $O/r1e_ln_fwd python tools/prof_elementwise_one.py ln_fwd > $O/r1e_p6.log 2>&1; echo "p6 rc=$?"
$NCU -k regex:ln_bwd -s 2 -c 1 -o $O/r1e_ln_bwd python tools/prof_elementwise_one.py ln_bwd > $O/r1e_p7.log 2>&1; echo "p7 rc=$?"
