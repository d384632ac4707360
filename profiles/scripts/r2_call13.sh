#!/bin/bash
set -x
O=gpurun_out/r2c13; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -k "resblock or chain" -q -m gpu -x > $O/t_rb.log 2>&1; echo "rc=$?" >> $O/t_rb.log
tail -3 $O/t_rb.log
timeout 300 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_resblock_bwd.txt 2>&1
head -2 $O/trace_resblock_bwd.txt
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err
head -8 $O/bench_default.err
