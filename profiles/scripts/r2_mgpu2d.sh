#!/bin/bash
set -x
O=gpurun_out/r2m2d; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_p2p_allreduce.py -q -m gpu -x > $O/t_p2p.log 2>&1; echo "rc=$?" >> $O/t_p2p.log
tail -6 $O/t_p2p.log
