#!/bin/bash
set -x
O=gpurun_out/r2m8; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_n8.json 2> $O/bench_n8.err
tail -3 $O/bench_n8.err; cut -c1-300 $O/bench_n8.json
