#!/bin/bash
# A/B: 2-D depthwise conv (forward + data gradient), warp = channel pair (NPF_DWCONV2D_V2=1, default) vs the 4-channel x 8-column x 2-row kernel
set -x
O=gpurun_out/r2c35; mkdir -p $O
B="python bench.py --workload gridconvcnp_b128_32x32 --steps 20 --warmup 5 --no-cpu-baseline --no-others --kernel-times"
timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "dwconv" > $O/t_v2.log 2>&1; echo "rc=$?" >> $O/t_v2.log; tail -n 12 $O/t_v2.log
NPF_DWCONV2D_V2=0 timeout 300 $B > $O/b_v1.json 2> $O/b_v1.err; cut -c1-160 $O/b_v1.json
timeout 300 $B > $O/b_v2.json 2> $O/b_v2.err; cut -c1-160 $O/b_v2.json
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -q -m gpu -k "gridconv" > $O/t_v2_models.log 2>&1; echo "rc=$?" >> $O/t_v2_models.log; tail -n 5 $O/t_v2_models.log
