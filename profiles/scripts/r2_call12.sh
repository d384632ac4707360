#!/bin/bash
set -x
O=gpurun_out/r2c12; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -k "dwconv or resblock" -q -m gpu > $O/t_dw.log 2>&1; echo "rc=$?" >> $O/t_dw.log
tail -3 $O/t_dw.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -q -m gpu -k "grid or convcnp_notebook or convlnp" > $O/t_models.log 2>&1; echo "rc=$?" >> $O/t_models.log
tail -3 $O/t_models.log
for wl in gridconvcnp_b128_32x32 gridconvlnp_b64_32x32_nz16; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 --kernel-times --no-cpu-baseline > $O/bench_$wl.json 2> $O/bench_$wl.err
  head -6 $O/bench_$wl.err
done
