#!/bin/bash
# round 2, call 10: record run of the current build -- full default bench line (with other workloads + reference arm), launch list, ncu --set full of the top kernels
set -x
O=gpurun_out/r2c10; mkdir -p $O
timeout 900 python bench.py --steps 100 --warmup 10 --kernel-times > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-others > $O/l_default.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'resblock1d|mlp_chain|setconv_tc|setconv_sorted|linear_bwd_fused64|linear_ws|thin_' --launch-skip 120 -c 40 -o $O/ncu_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-others > $O/ncu_top.log 2>&1
timeout 300 python profiles/microbench/trace_resblock.py > $O/trace_resblock_fwd.txt 2>&1
ls -la $O
