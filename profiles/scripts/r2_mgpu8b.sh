#!/bin/bash
set -x
O=gpurun_out/r2m8b; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 60 --warmup 5 --no-cpu-baseline --no-others > $O/bench_n8_p2p.json 2> $O/bench_n8_p2p.err
tail -2 $O/bench_n8_p2p.err; cut -c1-200 $O/bench_n8_p2p.json
NPF_P2P_ALLREDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 60 --warmup 5 --no-cpu-baseline --no-others > $O/bench_n8_nccl.json 2> $O/bench_n8_nccl.err
cut -c1-200 $O/bench_n8_nccl.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 60 --warmup 5 --no-cpu-baseline --no-others > $O/bench_n4_p2p.json 2> $O/bench_n4_p2p.err
cut -c1-200 $O/bench_n4_p2p.json
