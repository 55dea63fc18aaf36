#!/bin/bash
# One GPU-box visit: quick live fuzz, the GPU test-suite, a bench line, the ncu launch list and one --set full capture of the
# forward kernel.  usage (under gpurun): bash tools/gpu_check.sh <tag> [fuzz_groups]
TAG=${1:-r02}; GROUPS_N=${2:-25}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python tools/gpu_fuzz_live.py $GROUPS_N 4242 2 1 > gpurun_out/${TAG}_fuzz.txt 2>&1; echo "fuzz rc=$?" | tee -a gpurun_out/${TAG}_fuzz.txt
tail -3 gpurun_out/${TAG}_fuzz.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-file-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_run.py 3000 2 > gpurun_out/${TAG}_prof_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_fwd python tools/prof_run.py 3000 2 > gpurun_out/${TAG}_ncu_fwd.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -8
