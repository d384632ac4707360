#!/bin/bash
set -x
O=gpurun_out/r2c9; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -k "resblock" -q -m gpu -x > $O/t_rb.log 2>&1; echo "rc=$?" >> $O/t_rb.log
tail -4 $O/t_rb.log
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err
tail -14 $O/bench_default.err
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -q -m gpu -k "convcnp or convlnp" > $O/t_models.log 2>&1; echo "rc=$?" >> $O/t_models.log
tail -3 $O/t_models.log
