#!/bin/bash
# GPU test-suite + smoke + default bench line.  usage (under gpurun): bash tools/gpu_tests.sh <tag>
TAG=${1:-t}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
