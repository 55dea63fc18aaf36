#!/bin/bash
# A/B of the file pipeline on one box: the tree as built vs a build with extra nvcc flags.  usage: bash tools/gpu_ab3.sh "<flags>"
mkdir -p gpurun_out
timeout 200 python tools/e2e_quick.py 3000 15 2>&1 | tail -1
( cd npore_b200/csrc && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=false $1 -shared -o /tmp/libnpore_ab.so api.cu bamio.cpp -lcudart -lz > /tmp/ab_build.log 2>&1; echo build rc=$? )
NPORE_B200_LIB=/tmp/libnpore_ab.so timeout 200 python tools/e2e_quick.py 3000 15 2>&1 | tail -1
timeout 200 python tools/e2e_quick.py 3000 15 2>&1 | tail -1
