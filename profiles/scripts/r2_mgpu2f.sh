#!/bin/bash
set -x
O=gpurun_out/r2m2f; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_p2p_allreduce.py tests/test_gpu_syncbn_nccl.py -q -m gpu -x > $O/t_mgpu.log 2>&1; echo "rc=$?" >> $O/t_mgpu.log
tail -4 $O/t_mgpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err; python -c "import json;d=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]);print('n2',d['ms_per_step'],d['value'],{k:round(v['value']) for k,v in d['other_workloads'].items()})"
