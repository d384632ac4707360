#!/bin/bash
# round 2, final record run of the build: default bench line (other workloads; reference arm beside it), launch list,
# ncu --set full of the top kernels of ConvCNP / AttnCNP / GridConvCNP, parity margins at the benched shapes
# (the whole GPU suite + smoke of the same build: profiles/scripts/r2_tests.sh -> 240 passed, 2 skipped [2-GPU tests])
set -x
O=gpurun_out/r2final; mkdir -p $O
timeout 900 python bench.py --steps 100 --warmup 10 --kernel-times > $O/bench_default.json 2> $O/bench_default.err; cut -c1-250 $O/bench_default.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-200 $O/bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-others > $O/l_default.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'resblock1d|mlp_chain|setconv_tc|setconv_sorted|linear_bwd_fused64|linear_ws|thin_' --launch-skip 120 -c 24 -o $O/ncu_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-others > $O/ncu_top.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'xattn' --launch-skip 4 -c 4 -o $O/ncu_attn python bench.py --workload attncnp_b64_c512_t512 --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu_attn.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'dwconv2d' --launch-skip 6 -c 6 -o $O/ncu_dw2 python bench.py --workload gridconvcnp_b128_32x32 --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu_dw2.log 2>&1
timeout 300 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -s > $O/t_baseline_shapes.log 2>&1; echo "rc=$?" >> $O/t_baseline_shapes.log
rm -f $O/l_default.log $O/ncu_top.log $O/ncu_attn.log $O/ncu_dw2.log
du -sh $O gpurun_out
