#!/bin/bash
set -x
O=gpurun_out/r2c11; mkdir -p $O
timeout 300 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_resblock_bwd.txt 2>&1
head -3 $O/trace_resblock_bwd.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -k "resblock" -q -m gpu -x > $O/t_rb.log 2>&1; echo "rc=$?" >> $O/t_rb.log
tail -3 $O/t_rb.log
