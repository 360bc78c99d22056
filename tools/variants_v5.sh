#!/bin/bash
# usage (under gpurun): bash tools/variants_v5.sh "<nvcc -D flags variant 1>" ...  -> GPU parity tests with the in-tree build, then a
# count bench line per variant (the last variant stays built: put the default last)
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --build-only 2> gpurun_out/variants_build.log
i=0
for flags in "$@"; do
  i=$((i+1))
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC,-O3,-pthread -shared -Xptxas -v \
       -I include $flags -o index4j_b200/libfmgpu.so index4j_b200/csrc/fmgpu.cu -lcudart 2> gpurun_out/variant_$i.nvcc.log
  grep -A2 "k_countE" gpurun_out/variant_$i.nvcc.log | grep -E "Used|spill" | tr '\n' ' '
  echo "== variant $i: $flags"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-lf 2> gpurun_out/variant_$i.log | tee gpurun_out/variant_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('   value %.1f M/s  step %.3f ms  kernel %.3f ms  frac %.3f  e2e %.1f M/s' % (d['value']/1e6, d['ms_per_step'], r['kernel_ms'], r['frac'], d['e2e']['value']/1e6))"
done
