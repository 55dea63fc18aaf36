#!/bin/bash
# A/B of the file pipeline on one box: the tree as built vs the same tree with one header of csrc/ replaced.
# usage: bash tools/gpu_ab4.sh <replacement file> <header name in csrc>
mkdir -p /tmp/ab/npore_b200 && cp -r npore_b200/csrc /tmp/ab/npore_b200/ && cp -r include /tmp/ab/ && cp "$1" /tmp/ab/npore_b200/csrc/"$2"
( cd /tmp/ab/npore_b200/csrc && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=false -shared -o /tmp/libnpore_ab.so api.cu bamio.cpp -lcudart -lz > /tmp/ab_build.log 2>&1; echo build rc=$? )
for k in 1 2; do
echo "as built:"; timeout 200 python tools/e2e_quick.py 3000 15 2>&1 | tail -1
echo "with $1:"; NPORE_B200_LIB=/tmp/libnpore_ab.so timeout 200 python tools/e2e_quick.py 3000 15 2>&1 | tail -1
done
