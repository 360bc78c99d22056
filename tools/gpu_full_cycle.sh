#!/bin/bash
# usage (under gpurun): bash tools/gpu_full_cycle.sh <tag>
# all GPU parity tests -> index build (cached per box) -> bench line (+cpu baseline, LF legs) -> reference arm -> ncu launch list
# -> ncu full captures of k_count and of the LF kernels
set -x
TAG=${1:-cycle}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
nproc
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
( time python bench.py --build-only ) 2> gpurun_out/${TAG}_build.log
tail -4 gpurun_out/${TAG}_build.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log
cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.log
cat gpurun_out/${TAG}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lf-steps 1 > /dev/null 2> gpurun_out/${TAG}_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:k_count -s 3 -c 1 -f -o gpurun_out/${TAG}_k_count \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lf > /dev/null 2> gpurun_out/${TAG}_ncu2.log
tail -2 gpurun_out/${TAG}_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:"k_locate|k_extract" -c 6 -f -o gpurun_out/${TAG}_k_lf \
    python tools/bench_lf.py --steps 1 --warmup 0 --check 0 --n-pat 200000 --n-eub 200000 --n-ext 200000 > /dev/null 2> gpurun_out/${TAG}_ncu3.log
tail -2 gpurun_out/${TAG}_ncu3.log
ls -la gpurun_out
