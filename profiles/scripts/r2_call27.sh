#!/bin/bash
# A/B: attention backward, 64-query blocks with the transposed score tile (two CTAs per SM) vs the 128 x 128 arrangement
set -x
O=gpurun_out/r2c27; mkdir -p $O
B="python bench.py --workload attncnp_b64_c512_t512 --steps 30 --warmup 5 --no-cpu-baseline --no-others --kernel-times"
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -q -m gpu -k "attn" -x > $O/t_q64.log 2>&1; echo "rc=$?" >> $O/t_q64.log; tail -n 12 $O/t_q64.log
NPF_XATTN_BWD_Q64=0 timeout 300 $B > $O/b_q128.json 2> $O/b_q128.err; cut -c1-160 $O/b_q128.json
timeout 300 $B > $O/b_q64.json 2> $O/b_q64.err; cut -c1-160 $O/b_q64.json
timeout 400 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -q -m gpu -k "attn" > $O/t_q64_models.log 2>&1; echo "rc=$?" >> $O/t_q64_models.log; tail -n 5 $O/t_q64_models.log
