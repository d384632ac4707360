#!/bin/bash
# A/B: mbar_wait retry loop out of line (-DNPF_MBAR_OUTLINE=1, prebuilt into profiles/tmp_alt) vs inline
set -x
O=gpurun_out/r2c23; mkdir -p $O
B="python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-others --kernel-times"
L=neural-process-family_b200/npf_b200/lib/libnpf_b200.so
timeout 300 $B > $O/b_inline.json 2> $O/b_inline.err; cut -c1-160 $O/b_inline.json
cp $L /tmp/lib_inline.so; cp profiles/tmp_alt/libnpf_b200.so $L
timeout 300 $B > $O/b_outline.json 2> $O/b_outline.err; cut -c1-160 $O/b_outline.json
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -q -m gpu -k "resblock or chain or linear or setconv" > $O/t_outline.log 2>&1; echo "rc=$?" >> $O/t_outline.log; tail -n 3 $O/t_outline.log
cp /tmp/lib_inline.so $L
timeout 300 $B > $O/b_inline2.json 2> $O/b_inline2.err; cut -c1-160 $O/b_inline2.json
cp profiles/tmp_alt/libnpf_b200.so $L
timeout 300 $B > $O/b_outline2.json 2> $O/b_outline2.err; cut -c1-160 $O/b_outline2.json
