set -x
cd /root/repo
mkdir -p gpurun_out
python bench.py --build-only 2> gpurun_out/build.log
tail -3 gpurun_out/build.log
# launch list (every launch with its device time)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_launches_count.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu1.log
tail -5 gpurun_out/ncu1.log
# full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:k_count -s 3 -c 1 -f -o gpurun_out/r01_k_count \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu2.log
tail -5 gpurun_out/ncu2.log
ls -la gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.log
cat gpurun_out/bench_r01_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.log
cat gpurun_out/bench_r01_ref.json
