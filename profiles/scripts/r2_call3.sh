#!/bin/bash
# round 2, GPU call 3: mlp_chain_bwd (new) parity, SetConv backward with the pipelined epilogue, bench
set -x
O=gpurun_out/r2c3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py -k "chain" -q -m gpu -x > $O/t_chain.log 2>&1; echo "rc=$?" >> $O/t_chain.log
tail -8 $O/t_chain.log
timeout 900 python -m pytest tests/test_gpu_ops.py -k "setconv or mlp" -q -m gpu > $O/t_ops.log 2>&1; echo "rc=$?" >> $O/t_ops.log
tail -4 $O/t_ops.log
timeout 1200 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_tc.py tests/test_gpu_optim.py -q -m gpu > $O/t_models.log 2>&1; echo "rc=$?" >> $O/t_models.log
tail -6 $O/t_models.log
timeout 600 python bench.py --steps 50 --warmup 5 --kernel-times --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'setconv_tc|mlp_chain_bwd' -c 8 -o $O/ncu_new python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu_new.log 2>&1
ls -la $O
