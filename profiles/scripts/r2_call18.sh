#!/bin/bash
set -x
O=gpurun_out/r2c18; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -k "resblock" -q -m gpu -x > $O/t_rb.log 2>&1; echo "rc=$?" >> $O/t_rb.log
tail -3 $O/t_rb.log
timeout 600 python bench.py --steps 60 --warmup 5 --kernel-times --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err
head -3 $O/bench_default.err
