#!/bin/bash
# A/B: RR slice lengths on the shipped build, then alternative builds in build/ab
mkdir -p gpurun_out
for sl in 256 512 1024 2048 4096 100000; do
  echo "== slice $sl"; NPORE_RR_SLICE=$sl timeout 300 python tools/prof_run.py 3000 4 2>&1 | tail -1 | python -c "
import sys,ast
d=ast.literal_eval(sys.stdin.read().strip())
print({k:d[k] for k in ('ms_forward','ms_kernels_total')})"
done 2>&1 | tee gpurun_out/ab_slice.txt
bash tools/gpu_ab.sh ${1:-y}
