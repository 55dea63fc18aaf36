#!/bin/bash
# A/B timing of alternative builds of the library (build/ab/*.so) on the C2 batch: kernel times without a profiler.
mkdir -p gpurun_out
for f in npore_b200/libnpore_b200.so build/ab/*.so; do
  echo "== $f"
  NPORE_B200_LIB=$PWD/$f timeout 300 python tools/prof_run.py 3000 4 2>&1 | tail -1 | python -c "
import sys,ast
d=ast.literal_eval(sys.stdin.read().strip())
print({k:d[k] for k in ('ms_forward','ms_annotate','ms_kernels_total','fwd_warps_per_sm')})"
done 2>&1 | tee gpurun_out/ab_${1:-x}.txt
