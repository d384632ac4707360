#!/bin/bash
set -x
O=gpurun_out/r2c21; mkdir -p $O
timeout 200 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_bwd_base.txt 2>&1
NPF_RB_BWD_TW=1 timeout 200 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_bwd_TW.txt 2>&1
head -3 $O/trace_bwd_base.txt $O/trace_bwd_TW.txt
