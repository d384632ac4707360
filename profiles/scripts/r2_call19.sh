#!/bin/bash
# A/B call: new fixtures (circular padding, XL), the aligned-run dW flush, resblock fwd product orientation / TMEM-resident A operand,
# thin_in128_bwd grid size.  Every variant is a separate process (the switches are read once per process).
set -x
O=gpurun_out/r2c19; mkdir -p $O
B="python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-others --kernel-times"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q -m gpu -k "extrap or xl_pre" > $O/t_newfix.log 2>&1; echo "rc=$?" >> $O/t_newfix.log; tail -5 $O/t_newfix.log
for m in 1 2; do
  NPF_RB_FWD_T=$m timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "resblock1d_fused" > $O/t_rbfwd_T$m.log 2>&1; echo "rc=$?" >> $O/t_rbfwd_T$m.log; tail -3 $O/t_rbfwd_T$m.log
done
timeout 300 $B > $O/b_base.json 2> $O/b_base.err; cut -c1-200 $O/b_base.json
for m in 1 2; do
  NPF_RB_FWD_T=$m timeout 300 $B > $O/b_T$m.json 2> $O/b_T$m.err; cut -c1-200 $O/b_T$m.json
done
for c in 4 8; do
  NPF_THIN_IN_BWD_CTAS=$c timeout 300 $B > $O/b_thin$c.json 2> $O/b_thin$c.err; cut -c1-200 $O/b_thin$c.json
done
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -q -m gpu -k "linear or setconv or chain" > $O/t_lin.log 2>&1; echo "rc=$?" >> $O/t_lin.log; tail -3 $O/t_lin.log
