#!/bin/bash
# usage (under gpurun): bash tools/gpu_e2e_try.sh <tag> "<nvcc flags>|<env assignments>" ...   (the last build stays)
TAG=$1; shift
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_utf8_patterns.py -m gpu -x -q 2>&1 | tail -3
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
i=0
for spec in "$@"; do
  i=$((i+1))
  flags="${spec%%|*}"; envs="${spec#*|}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC,-O3,-pthread -shared -Xptxas -v \
       -I include $flags -o index4j_b200/libfmgpu.so index4j_b200/csrc/fmgpu.cu -lcudart 2> gpurun_out/${TAG}_$i.nvcc.log
  echo "== $i: flags [$flags] env [$envs]"
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-lf 2> gpurun_out/${TAG}_$i.log | tee gpurun_out/${TAG}_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('   value %.1f M/s  kernel %.3f ms  e2e %.1f M/s  e2e_utf8 %.1f M/s' % (d['value']/1e6, r['kernel_ms'], d['e2e']['value']/1e6, (d.get('e2e_utf8') or {'value':0})['value']/1e6))"
done
