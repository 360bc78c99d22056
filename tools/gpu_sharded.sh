#!/bin/bash
# round 2, sharded index (BASELINE.json configs[4]): bash tools/gpu_sharded.sh <tag> <N> <shard_chars> <n_pat>   (under gpurun --gpus N)
TAG=${1:-r2s}; N=${2:-2}; S=${3:-16777216}; NP=${4:-20000}
mkdir -p gpurun_out
free -g | head -2; nproc
if [ "$5" == "tests" ]; then
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_records.py -q 2>&1 | tail -8
fi
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --mode sharded --gpus $N \
    --steps 3 --shard-chars $S --sharded-n-pat $NP > gpurun_out/${TAG}_sharded_n${N}.json 2> gpurun_out/${TAG}_sharded_n${N}.log
grep -a "sharded r0\|Error\|error\|Traceback\|assert" gpurun_out/${TAG}_sharded_n${N}.log | tail -8
cut -c1-1800 gpurun_out/${TAG}_sharded_n${N}.json
