#!/bin/bash
# A/B: fwd tiles of 96 rows (7 balanced rounds) vs 128 (6 rounds, 768 rows on the critical CTA); bwd with interior-only O recompute + trimmed epilogue
set -x
O=gpurun_out/r2c24; mkdir -p $O
B="python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-others --kernel-times"
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "resblock1d_fused" > $O/t_tr96.log 2>&1; echo "rc=$?" >> $O/t_tr96.log; tail -n 2 $O/t_tr96.log
NPF_RB_FWD_TR=128 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "resblock1d_fused" > $O/t_tr128.log 2>&1; echo "rc=$?" >> $O/t_tr128.log; tail -n 2 $O/t_tr128.log
NPF_RB_FWD_TR=128 timeout 300 $B > $O/b_tr128.json 2> $O/b_tr128.err; cut -c1-160 $O/b_tr128.json
timeout 300 $B > $O/b_tr96.json 2> $O/b_tr96.err; cut -c1-160 $O/b_tr96.json
NPF_RB_FWD_TR=128 timeout 300 $B > $O/b_tr128b.json 2> $O/b_tr128b.err; cut -c1-160 $O/b_tr128b.json
timeout 300 $B > $O/b_tr96b.json 2> $O/b_tr96b.err; cut -c1-160 $O/b_tr96b.json
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -k "convcnp" > $O/t_base.log 2>&1; echo "rc=$?" >> $O/t_base.log; tail -n 2 $O/t_base.log
timeout 200 python profiles/microbench/trace_resblock.py 256 384 1 > $O/trace_bwd.txt 2>&1
timeout 200 python profiles/microbench/trace_resblock.py 256 384 0 > $O/trace_fwd.txt 2>&1
