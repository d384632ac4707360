#!/bin/bash
# round 2, GPU call 2: new SetConv kernels (sorted few-channel, tcgen05 backward) -- parity first, then bench + ncu
set -x
O=gpurun_out/r2c2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -k "setconv" -q -m gpu > $O/t_setconv.log 2>&1; echo "rc=$?" >> $O/t_setconv.log
tail -5 $O/t_setconv.log
timeout 1200 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_tc.py -q -m gpu -s > $O/t_models.log 2>&1; echo "rc=$?" >> $O/t_models.log
tail -5 $O/t_models.log
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'setconv' -c 10 -o $O/ncu_setconv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu_setconv.log 2>&1
ls -la $O
