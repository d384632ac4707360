#!/bin/bash
O=gpurun_out/r2tests2; mkdir -p $O
timeout 110 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log; tail -n 30 $O/t_all.log
