#!/bin/bash
set -x
O=gpurun_out/r2m8c; mkdir -p $O
run() { name=$1; n=$2; shift; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 60 --warmup 5 --no-cpu-baseline --no-others > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.loads(open('$O/bench_$name.json').read().strip().splitlines()[-1]);print('$name',d['ms_per_step'],d['value'])"; }
run n8_two 8 NPF_P2P_TWO_SHOT=1
run n8_nccl 8 NPF_P2P_ALLREDUCE=0
run n4_two 4 NPF_P2P_TWO_SHOT=1
run n4_nccl 4 NPF_P2P_ALLREDUCE=0
