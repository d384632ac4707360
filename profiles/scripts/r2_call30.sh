#!/bin/bash
set -x
O=gpurun_out/r2c32; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dwconv2d' --launch-skip 6 -c 4 -o $O/ncu_dw2 python bench.py --workload gridconvcnp_b128_32x32 --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-others > $O/ncu.log 2>&1
tail -5 $O/ncu.log
ls -la $O
