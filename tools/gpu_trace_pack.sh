#!/bin/bash
# per-chunk timelines of the host-pointer count call under a few settings of the packed transport
mkdir -p gpurun_out
python bench.py --build-only 2>/dev/null
i=0
for envs in ${TRACE_ENVS:-"A=1" "FMGPU_PACK_HELP=0" "CUDA_DEVICE_MAX_CONNECTIONS=32" "FMGPU_HOST_PACK=0"}; do
  i=$((i+1))
  echo "== $envs"
  env $envs FMGPU_PIPE_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sr-sweep --no-lf 2> gpurun_out/trace_$i.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   e2e %.1f M/s  utf8 %.1f M/s' % (d['e2e']['value']/1e6, d['e2e_utf8']['value']/1e6))"
  grep -n "packed transport" gpurun_out/trace_$i.log | sed -n 4p
  L=$(grep -n "packed transport" gpurun_out/trace_$i.log | sed -n 4p | cut -d: -f1)
  sed -n "$((L-8)),$((L-1))p;$((L+1)),$((L+8))p" gpurun_out/trace_$i.log | cut -c15-
done
