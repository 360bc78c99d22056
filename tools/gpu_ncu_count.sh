#!/bin/bash
# usage (under gpurun): bash tools/gpu_ncu_count.sh <tag>   -> ncu --set full of one k_count launch (cfg-2 workload)
TAG=${1:-ncu}
cd /root/repo; mkdir -p gpurun_out
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
ncu --set full --clock-control none --import-source on -k regex:k_count -s 3 -c 1 -f -o gpurun_out/${TAG}_k_count \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lf > /dev/null 2> gpurun_out/${TAG}_ncu.log
tail -2 gpurun_out/${TAG}_ncu.log
