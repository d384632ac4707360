#!/bin/bash
# round 2, GPU call 4: re-measure the SetConv backward (deep V prefetch) and the chain backward (L2 prefetch), train_models on device
set -x
O=gpurun_out/r2c4; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py tests/test_gpu_train.py -k "setconv or chain or mlp or train" -q -m gpu > $O/t_ops.log 2>&1; echo "rc=$?" >> $O/t_ops.log
tail -6 $O/t_ops.log
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'setconv_tc_bwd|mlp_chain_bwd' -c 4 -o $O/ncu_new python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu_new.log 2>&1
ls -la $O
