#!/bin/bash
# round 2, cycle C (2 GPUs): full GPU parity suite (incl. replicated index + fused records), N=2 bench with strong-scaling legs
TAG=${1:-r2c}; N=${2:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.log
grep -a "Error\|error\|Traceback" gpurun_out/${TAG}_bench_n${N}.log | tail -5
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print("value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]))
print("strong", json.dumps(d.get("strong"))[:1200])
print("records", json.dumps(d.get("locate_records"))[:1200])
PY
