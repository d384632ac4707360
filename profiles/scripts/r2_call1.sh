#!/bin/bash
# round 2, GPU call 1: baseline numbers of every BASELINE workload with the round-1 kernels + pretrained eval + attention / dwconv2d ncu
set -x
O=gpurun_out/r2c1; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -q -s -m gpu > $O/baseline_shapes.log 2>&1
timeout 600 python examples/eval_pretrained.py 10200 > $O/eval_pretrained.jsonl 2> $O/eval_pretrained.err
for wl in cnp_b16_c32_t64 attncnp_b64_c512_t512 attncnp_b256_c512_t512 gridconvcnp_b128_32x32 gridconvlnp_b64_32x32_nz16; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 --kernel-times > $O/bench_$wl.json 2> $O/bench_$wl.err
done
timeout 400 python bench.py --steps 50 --warmup 5 --kernel-times > $O/bench_default.json 2> $O/bench_default.err
# ncu: attention kernels (fwd+bwd) at B=64 C=T=512 and the 2-D depthwise kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'xattn' -c 8 -o $O/ncu_xattn python bench.py --workload attncnp_b64_c512_t512 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $O/ncu_xattn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dwconv2d|gridconv' -c 12 -o $O/ncu_dw2d python bench.py --workload gridconvcnp_b128_32x32 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $O/ncu_dw2d.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_attn.csv python bench.py --workload attncnp_b64_c512_t512 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $O/l_attn.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_grid.csv python bench.py --workload gridconvcnp_b128_32x32 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $O/l_grid.log 2>&1
ls -la $O
