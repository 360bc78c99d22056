#!/bin/bash
# usage (under gpurun): bash tools/gpu_cycle.sh <tag> [pytest-expr]
# parity tests -> index build (cached per box) -> bench line -> ncu full capture of k_count
set -x
TAG=${1:-cycle}
KEXPR=${2:-"count"}
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$KEXPR" 2>&1 | tail -5
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log
cat gpurun_out/${TAG}_bench.json
ncu --set full --clock-control none --import-source on -k regex:k_count -s 3 -c 1 -f -o gpurun_out/${TAG}_k_count \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu.log
tail -2 gpurun_out/${TAG}_ncu.log
