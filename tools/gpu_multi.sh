#!/bin/bash
# round 2, multi-GPU: bash tools/gpu_multi.sh <tag> <N>   (under gpurun --gpus N)
TAG=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -6 > gpurun_out/${TAG}_n${N}_pytest.txt; tail -3 gpurun_out/${TAG}_n${N}_pytest.txt
python bench.py --build-only 2> gpurun_out/${TAG}_n${N}_build.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.log
grep -a "strong\|Error\|error" gpurun_out/${TAG}_bench_n${N}.log | tail -5
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print("value %.3g e2e %.3g strong %s" % (d["value"], d["e2e"]["value"], json.dumps(d.get("strong"))[:900]))
PY
