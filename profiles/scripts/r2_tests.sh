#!/bin/bash
set -x
O=gpurun_out/r2tests; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log; tail -n 40 $O/t_all.log
