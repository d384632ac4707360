#!/bin/bash
set -x
O=gpurun_out/r2c15; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_graph.py -q -m gpu -x > $O/t_graph.log 2>&1; echo "rc=$?" >> $O/t_graph.log
tail -4 $O/t_graph.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-others > $O/bench_default.json 2> $O/bench_default.err; python -c "import json;d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'])"
