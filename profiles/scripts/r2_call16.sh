#!/bin/bash
set -x
O=gpurun_out/r2c16; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -k "setconv" -q -m gpu -x > $O/t_sc.log 2>&1; echo "rc=$?" >> $O/t_sc.log
tail -4 $O/t_sc.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -q -m gpu -k "convcnp or convlnp" > $O/t_models.log 2>&1; echo "rc=$?" >> $O/t_models.log
tail -3 $O/t_models.log
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err
head -8 $O/bench_default.err
timeout 400 ncu --set full --clock-control none -k regex:'setconv_tc_fwd' -c 2 -o $O/ncu_scf python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu_scf.log 2>&1
