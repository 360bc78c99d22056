#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [shard_chars]
set -x
TAG=${1:-multi}; N=${2:-2}; SHARD=${3:-1073741824}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.log
tail -3 gpurun_out/${TAG}_bench_n$N.log
cat gpurun_out/${TAG}_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_sharded.py --shard-chars $SHARD \
    > gpurun_out/${TAG}_sharded_n$N.json 2> gpurun_out/${TAG}_sharded_n$N.log
tail -8 gpurun_out/${TAG}_sharded_n$N.log
cat gpurun_out/${TAG}_sharded_n$N.json
