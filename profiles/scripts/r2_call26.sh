#!/bin/bash
# A/B: attention forward with 64-key chunks (4 CTAs/SM) vs 128-key chunks (2 CTAs/SM)
set -x
O=gpurun_out/r2c26; mkdir -p $O
B="python bench.py --workload attncnp_b64_c512_t512 --steps 30 --warmup 5 --no-cpu-baseline --no-others --kernel-times"
NPF_XATTN_KC=64 timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -q -m gpu -k "attn" > $O/t_kc64.log 2>&1; echo "rc=$?" >> $O/t_kc64.log; tail -n 2 $O/t_kc64.log
NPF_XATTN_KC=128 timeout 300 $B > $O/b_kc128.json 2> $O/b_kc128.err; cut -c1-160 $O/b_kc128.json
NPF_XATTN_KC=64 timeout 300 $B > $O/b_kc64.json 2> $O/b_kc64.err; cut -c1-160 $O/b_kc64.json
NPF_XATTN_KC=64 timeout 400 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -q -m gpu -k "attn" > $O/t_kc64_models.log 2>&1; echo "rc=$?" >> $O/t_kc64_models.log; tail -n 2 $O/t_kc64_models.log
