set -x
cd /root/repo
mkdir -p gpurun_out
nproc; free -g | head -2
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.log
tail -30 gpurun_out/bench_first.log
cat gpurun_out/bench_first.json
