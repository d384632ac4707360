#!/bin/bash
# A/B: bwd -- weight-gradient MMAs held back until the epilogue has read its first dO window (NPF_RB_BWD_LDF=1)
set -x
O=gpurun_out/r2c25; mkdir -p $O
B="python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-others --kernel-times"
NPF_RB_BWD_LDF=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "resblock1d_fused" > $O/t_ldf.log 2>&1; echo "rc=$?" >> $O/t_ldf.log; tail -n 2 $O/t_ldf.log
timeout 300 $B > $O/b_ldf0.json 2> $O/b_ldf0.err; cut -c1-160 $O/b_ldf0.json
NPF_RB_BWD_LDF=1 timeout 300 $B > $O/b_ldf1.json 2> $O/b_ldf1.err; cut -c1-160 $O/b_ldf1.json
timeout 300 $B > $O/b_ldf0b.json 2> $O/b_ldf0b.err; cut -c1-160 $O/b_ldf0b.json
NPF_RB_BWD_LDF=1 timeout 300 $B > $O/b_ldf1b.json 2> $O/b_ldf1b.err; cut -c1-160 $O/b_ldf1b.json
NPF_RB_BWD_LDF=1 timeout 200 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_bwd_ldf1.txt 2>&1
