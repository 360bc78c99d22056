#!/bin/bash
# round 2, cycle B: GPU parity suite, bench line with the flat kernel (v6) and a v5 comparison, ncu launch list + --set full captures
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
python bench.py --build-only 2> gpurun_out/${TAG}_build.log; tail -2 gpurun_out/${TAG}_build.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.log
tail -3 gpurun_out/${TAG}_bench_n1.log; cut -c1-700 gpurun_out/${TAG}_bench_n1.json
FMGPU_COUNT_KERNEL=5 timeout 600 python bench.py --steps 10 --warmup 3 --no-lf --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_v5.json 2> gpurun_out/${TAG}_bench_n1_v5.log
cut -c1-300 gpurun_out/${TAG}_bench_n1_v5.json
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_launches.log
ncu --set full --clock-control none --import-source on -k regex:k_count_flat -s 3 -c 1 -f -o gpurun_out/${TAG}_k_count \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lf > /dev/null 2> gpurun_out/${TAG}_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -f -o gpurun_out/${TAG}_k_locate \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:k_extract -s 1 -c 1 -f -o gpurun_out/${TAG}_k_extract \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu3.log
ls -la gpurun_out/${TAG}_*.ncu-rep
fi
