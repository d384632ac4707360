#!/bin/bash
set -x
O=gpurun_out/r2c14; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
tail -4 $O/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-250 $O/bench_default.json
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-200 $O/bench_reference.json
