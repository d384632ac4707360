#!/bin/bash
# A/B: resblock bwd with Wpw^T in TMEM + 3 raw X / 2 raw dY buffers (NPF_RB_BWD_TW=1); regenerated circular fixtures; GP per-task hyper-parameters
set -x
O=gpurun_out/r2c20; mkdir -p $O
B="python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-others --kernel-times"
NPF_RB_BWD_TW=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "resblock1d_fused" > $O/t_rbbwd_TW.log 2>&1; echo "rc=$?" >> $O/t_rbbwd_TW.log; tail -n 3 $O/t_rbbwd_TW.log
timeout 300 $B > $O/b_base.json 2> $O/b_base.err; cut -c1-200 $O/b_base.json
NPF_RB_BWD_TW=1 timeout 300 $B > $O/b_TW.json 2> $O/b_TW.err; cut -c1-200 $O/b_TW.json
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_gp.py -q -m gpu -k "extrap or xl_pre or hyp" > $O/t_new.log 2>&1; echo "rc=$?" >> $O/t_new.log; tail -n 5 $O/t_new.log
NPF_RB_BWD_TW=1 timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -k "convcnp" > $O/t_base_TW.log 2>&1; echo "rc=$?" >> $O/t_base_TW.log; tail -n 3 $O/t_base_TW.log
