#!/bin/bash
# usage (under gpurun): bash tools/gpu_lf_cycle.sh <tag> [bench_lf args]
# gather micro-benchmark -> LF workloads (locate / extractUntilBoundary / extract) with oracle spot checks -> ncu of k_walk
set -x
TAG=${1:-lf}; shift
cd /root/repo
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_peak tools/gather_peak.cu && /tmp/gather_peak > gpurun_out/${TAG}_gather_peak.jsonl
cat gpurun_out/${TAG}_gather_peak.jsonl
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
python tools/bench_lf.py "$@" > gpurun_out/${TAG}_lf.json 2> gpurun_out/${TAG}_lf.log
tail -5 gpurun_out/${TAG}_lf.log
cat gpurun_out/${TAG}_lf.json
ncu --set full --clock-control none --import-source on -k regex:k_walk -c 3 -f -o gpurun_out/${TAG}_k_walk \
    python tools/bench_lf.py --steps 1 --warmup 0 --check 0 --n-pat 200000 --n-eub 200000 --n-ext 200000 > /dev/null 2> gpurun_out/${TAG}_ncu.log
tail -2 gpurun_out/${TAG}_ncu.log
