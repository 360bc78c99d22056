#!/bin/bash
# usage (under gpurun): bash tools/gpu_trace.sh "<env assignments>" ...   -> count parity tests, then per env: e2e rates + one traced call
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_utf8_patterns.py -m gpu -x -q -k "count" 2>&1 | tail -2
python bench.py --build-only 2> gpurun_out/trace_build.log
for envs in "$@"; do
  echo "== env [$envs]"
  env $envs python tools/pipe_trace.py 2>&1 | grep -v "^\[bench\]"
done
