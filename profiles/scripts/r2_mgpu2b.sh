#!/bin/bash
set -x
O=gpurun_out/r2m2b; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 60 --warmup 5 --no-cpu-baseline --no-others > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.loads(open('$O/bench_$name.json').read().strip().splitlines()[-1]);print('$name',d['ms_per_step'],d['value'])"; }
run outside_avg NPF_GRAPH_ALLREDUCE=0 NPF_ALLREDUCE_AVG=1
run ingraph_avg NPF_GRAPH_ALLREDUCE=1 NPF_ALLREDUCE_AVG=1
run outside_sum NPF_GRAPH_ALLREDUCE=0 NPF_ALLREDUCE_AVG=0
run ingraph_sum NPF_GRAPH_ALLREDUCE=1 NPF_ALLREDUCE_AVG=0
timeout 200 python -m pytest tests/test_gpu_syncbn_nccl.py -q -m gpu -x > $O/t_syncbn.log 2>&1; echo "rc=$?" >> $O/t_syncbn.log
tail -5 $O/t_syncbn.log
