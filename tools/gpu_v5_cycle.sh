#!/bin/bash
# usage (under gpurun): bash tools/gpu_v5_cycle.sh <tag> "<variant flags>"...   -> all GPU parity tests, a count bench line per
# variant (the LAST variant stays built and is the one the full bench line is taken with)
TAG=$1; shift
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
i=0
for flags in "$@"; do
  i=$((i+1))
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC,-O3,-pthread -shared -Xptxas -v \
       -I include $flags -o index4j_b200/libfmgpu.so index4j_b200/csrc/fmgpu.cu -lcudart 2> gpurun_out/${TAG}_variant_$i.nvcc.log
  grep -A2 "k_countILb0" gpurun_out/${TAG}_variant_$i.nvcc.log | grep -E "Used|spill" | tr '\n' ' '
  echo "== variant $i: $flags"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-lf 2> gpurun_out/${TAG}_variant_$i.log | tee gpurun_out/${TAG}_variant_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('   value %.1f M/s  step %.3f ms  kernel %.3f ms  frac %.3f  e2e %.1f M/s' % (d['value']/1e6, d['ms_per_step'], r['kernel_ms'], r['frac'], d['e2e']['value']/1e6))"
done
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log
tail -c 2500 gpurun_out/${TAG}_bench.json
