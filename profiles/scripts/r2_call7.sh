#!/bin/bash
set -x
O=gpurun_out/r2c7; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py -k "chain" -q -m gpu -x > $O/t_chain.log 2>&1; echo "rc=$?" >> $O/t_chain.log
tail -5 $O/t_chain.log
timeout 300 python profiles/microbench/trace_resblock.py > $O/trace_resblock_fwd.txt 2>&1
timeout 300 python profiles/microbench/trace_timeline.py > $O/trace_chain_bwd.txt 2>&1
head -3 $O/trace_chain_bwd.txt
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err
ls -la $O
