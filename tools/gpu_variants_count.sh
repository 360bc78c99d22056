#!/bin/bash
# k_count variants: "ENV=... | nvcc flags" per argument; count leg only
TAG=${1:-vc}; shift
mkdir -p gpurun_out
python bench.py --build-only 2> gpurun_out/${TAG}_build.log; tail -1 gpurun_out/${TAG}_build.log
i=0
for spec in "$@"; do
  i=$((i+1))
  envs="${spec%%|*}"; flags="${spec#*|}"
  if [ "$flags" != "$prev_flags" ] || [ $i -eq 1 ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC,-O3,-pthread -shared -Xptxas -v \
       -I include $flags -o index4j_b200/libfmgpu.so index4j_b200/csrc/fmgpu.cu -lcudart 2> gpurun_out/${TAG}_variant_$i.nvcc.log
  prev_flags="$flags"
  fi
  [ -f gpurun_out/${TAG}_variant_$i.nvcc.log ] && grep -A2 "k_countILb0" gpurun_out/${TAG}_variant_$i.nvcc.log | grep -E "Used|spill" | tr '\n' ' ' | sed -e 's/ptxas info *://g' -e 's/bytes//g'
  echo "== variant $i: env[$envs] flags[$flags]"
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sr-sweep --no-lf 2> gpurun_out/${TAG}_variant_$i.log | tee gpurun_out/${TAG}_variant_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('   count %.1f M/s  kernel %.3f ms  frac %.3f  e2e %.1f M/s  utf8 %.1f M/s' % (d['value']/1e6, r['kernel_ms'], r['frac'], d['e2e']['value']/1e6, d['e2e_utf8']['value']/1e6))"
done
