#!/bin/bash
# round 2, GPU call 6: fused ResConvBlock backward parity + bench + ncu
set -x
O=gpurun_out/r2c6; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -k "resblock" -q -m gpu -x > $O/t_rb.log 2>&1; echo "rc=$?" >> $O/t_rb.log
tail -12 $O/t_rb.log
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_tc.py -q -m gpu -k "convcnp or convlnp or graph" > $O/t_models.log 2>&1; echo "rc=$?" >> $O/t_models.log
tail -4 $O/t_models.log
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err
tail -3 $O/bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'resblock' -c 6 -o $O/ncu_rb python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu_rb.log 2>&1
ls -la $O
