#!/bin/bash
# usage (under gpurun): bash tools/variants_lf.sh "<nvcc -D flags variant 1>" ...   -> LF workload line per variant
cd /root/repo
mkdir -p gpurun_out
python bench.py --build-only 2> gpurun_out/variants_build.log
i=0
for flags in "$@"; do
  i=$((i+1))
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC,-O3,-pthread -shared -Xptxas -v \
       -I include $flags -o index4j_b200/libfmgpu.so index4j_b200/csrc/fmgpu.cu -lcudart 2> gpurun_out/variant_$i.nvcc.log
  grep -A2 "k_locate\|k_extract" gpurun_out/variant_$i.nvcc.log | grep -E "Used" | tr '\n' ' '
  echo "== variant $i: $flags"
  python tools/bench_lf.py --check 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('   locate ms %.2f (%.2f G hits/s)  eub ms %.3f  extract ms %.3f' % (d['locate']['ms_per_step'], d['locate']['hits_per_s']/1e9, d['eub']['ms_per_step'], d['extract']['ms_per_step']))"
done
