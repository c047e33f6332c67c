#!/bin/bash
# Round-1g evidence job (one B200): final bench (with the CPU baseline), the reference arm, ncu launch list,
# ncu --set full of the kernels that changed this session.
O=gpurun_out
python bench.py > $O/bench38.json 2> $O/bench38.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench38_ref.json 2> $O/bench38_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r1g_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-timing > $O/r1g_ncu_list.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:attn_bwd_sn -s 2 -c 1 -o $O/r1g_attn_bwd_sn python tools/prof_attn_one.py bwd > $O/r1g_p5.log 2>&1; echo "p5 rc=$?"
$NCU -k regex:gemm2 -s 487 -c 8 -o $O/r1g_gemm_bwd_inmodel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-timing > $O/r1g_b.log 2>&1; echo "gemm bwd rc=$?"
$NCU -k regex:patchify -s 2 -c 1 -o $O/r1g_patchify python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-timing > $O/r1g_c.log 2>&1; echo "patchify rc=$?"
python tools/bench_configs.py detection > $O/det_device2.json 2> $O/det_device2.err; tail -1 $O/det_device2.json
python tools/bench_configs.py segmentation > $O/seg2.json 2> $O/seg2.err; tail -1 $O/seg2.json
