#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (small shapes: the sanitizer slows a kernel 10-50x)
set -x
O=gpurun_out/r2san; mkdir -p $O
export NPF_MBAR_TRAP=0
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "resblock1d_fused and (1-7 or 2-128 or 1-129 or 3-100)" > $O/san_resblock.log 2>&1; echo "rc=$?" >> $O/san_resblock.log
tail -5 $O/san_resblock.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "chain_bwd_entry and (2-64 or 4-100 or 5-257 or 4-1024)" > $O/san_chain.log 2>&1; echo "rc=$?" >> $O/san_chain.log
tail -5 $O/san_chain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "setconv and (2-384-128 or 2-384-50 or 3-128-384 or 5-500-100 or 3-17-29)" > $O/san_setconv.log 2>&1; echo "rc=$?" >> $O/san_setconv.log
tail -5 $O/san_setconv.log
