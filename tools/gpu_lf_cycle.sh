#!/bin/bash
# usage (under gpurun): bash tools/gpu_lf_cycle.sh <tag> [bench_lf args]
# locate/extract parity tests -> LF workloads (locate / extractUntilBoundary / extract) with oracle spot checks -> ncu of the LF kernels
set -x
TAG=${1:-lf}; shift
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "locate or extract or smoke" 2>&1 | tail -5
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
python tools/bench_lf.py "$@" > gpurun_out/${TAG}_lf.json 2> gpurun_out/${TAG}_lf.log
tail -5 gpurun_out/${TAG}_lf.log
cat gpurun_out/${TAG}_lf.json
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_locate" -c 3 -f -o gpurun_out/${TAG}_k_lf \
    python tools/bench_lf.py --steps 1 --warmup 0 --check 0 --n-pat 200000 --n-eub 200000 --n-ext 200000 > /dev/null 2> gpurun_out/${TAG}_ncu.log
tail -2 gpurun_out/${TAG}_ncu.log
