#!/bin/bash
set -x
O=gpurun_out/r2m2c; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_p2p_allreduce.py -q -m gpu -x > $O/t_p2p.log 2>&1; echo "rc=$?" >> $O/t_p2p.log
tail -15 $O/t_p2p.log
run() { name=$1; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 60 --warmup 5 --no-cpu-baseline --no-others > $O/bench_$name.json 2> $O/bench_$name.err; tail -2 $O/bench_$name.err; python -c "import json;d=json.loads(open('$O/bench_$name.json').read().strip().splitlines()[-1]);print('$name',d['ms_per_step'],d['value'])"; }
run p2p NPF_P2P_ALLREDUCE=1
run nccl NPF_P2P_ALLREDUCE=0
run p2p_ingraph NPF_P2P_ALLREDUCE=1 NPF_GRAPH_ALLREDUCE=1
timeout 200 python -m pytest tests/test_gpu_syncbn_nccl.py -q -m gpu -x > $O/t_syncbn.log 2>&1; echo "rc=$?" >> $O/t_syncbn.log
tail -3 $O/t_syncbn.log
