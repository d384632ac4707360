#!/bin/bash
# round 2, record run of the final build (the two earlier attempts, r2_final.sh, produced > 64 MiB of ncu reports and were not copied back):
# default bench line with the other workloads + launch list.  GPU suite of the same build: r2_tests.sh (240 passed, 2 skipped).
set -x
O=gpurun_out/r2final; mkdir -p $O
timeout 600 python bench.py --steps 100 --warmup 10 --kernel-times > $O/bench_default.json 2> $O/bench_default.err; cut -c1-250 $O/bench_default.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-others > /dev/null 2>&1
du -sh $O
