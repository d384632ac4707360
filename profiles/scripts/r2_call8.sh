#!/bin/bash
set -x
O=gpurun_out/r2c8; mkdir -p $O
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
tail -3 $O/bench_default.err
timeout 300 python profiles/microbench/trace_resblock.py > $O/trace_resblock_fwd.txt 2>&1
timeout 300 python profiles/microbench/trace_timeline.py > $O/trace_chain_bwd.txt 2>&1
head -3 $O/trace_chain_bwd.txt; head -2 $O/trace_resblock_fwd.txt
timeout 1500 python -m pytest tests -q -m gpu -x > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
tail -5 $O/t_all.log
ls -la $O
