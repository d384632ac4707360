#!/bin/bash
# A/B: resblock bwd epilogue with 12 warps (3 row groups x 16 rows, one pass) vs 8, with / without Wpw^T in TMEM + deeper raw rings
set -x
O=gpurun_out/r2c22; mkdir -p $O
B="python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-others --kernel-times"
for cfg in "8 0" "8 1" "12 0" "12 1"; do
  set -- $cfg
  NPF_RB_BWD_EPI=$1 NPF_RB_BWD_TW=$2 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "resblock1d_fused" > $O/t_e$1_tw$2.log 2>&1; echo "rc=$?" >> $O/t_e$1_tw$2.log; tail -n 2 $O/t_e$1_tw$2.log
  NPF_RB_BWD_EPI=$1 NPF_RB_BWD_TW=$2 timeout 300 $B > $O/b_e$1_tw$2.json 2> $O/b_e$1_tw$2.err; cut -c1-160 $O/b_e$1_tw$2.json
done
NPF_RB_BWD_EPI=12 NPF_RB_BWD_TW=1 timeout 200 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_bwd_e12_tw1.txt 2>&1
NPF_RB_BWD_EPI=12 NPF_RB_BWD_TW=0 timeout 200 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_bwd_e12_tw0.txt 2>&1
