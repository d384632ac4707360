#!/bin/bash
set -x
O=gpurun_out/r2m2; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_syncbn_nccl.py -q -m gpu -x > $O/t_syncbn.log 2>&1; echo "rc=$?" >> $O/t_syncbn.log
tail -5 $O/t_syncbn.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-others > $O/bench_n2.json 2> $O/bench_n2.err
tail -2 $O/bench_n2.err; cat $O/bench_n2.json | cut -c1-400
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline --no-others > $O/bench_n1.json 2> $O/bench_n1.err
cat $O/bench_n1.json | cut -c1-300
