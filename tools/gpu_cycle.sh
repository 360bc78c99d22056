#!/bin/bash
# round 2, cycle G: 8-byte cells + dense locate samples.  GPU parity suite, CTA-shape variants of k_count / k_locate<dense>
# (each: count + locate + extractUntilBoundary legs), full bench line, ncu launch list + captures
TAG=${1:-r2g}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
python bench.py --build-only 2> gpurun_out/${TAG}_build.log; tail -2 gpurun_out/${TAG}_build.log
i=0
for flags in "$@"; do
  i=$((i+1))
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC,-O3,-pthread -shared -Xptxas -v \
       -I include $flags -o index4j_b200/libfmgpu.so index4j_b200/csrc/fmgpu.cu -lcudart 2> gpurun_out/${TAG}_variant_$i.nvcc.log
  for k in k_countILb0 k_locateILb0ELb1; do grep -A2 "$k" gpurun_out/${TAG}_variant_$i.nvcc.log | grep -E "Used|spill" | tr '\n' ' ' | sed -e 's/ptxas info *://g' -e 's/bytes//g'; done
  echo "== variant $i: $flags"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sr-sweep 2> gpurun_out/${TAG}_variant_$i.log | tee gpurun_out/${TAG}_variant_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; l=d['locate']; e=d['extract_until_boundary']
print('   count %.1f M/s  kernel %.3f ms  frac %.3f  e2e %.1f M/s  utf8 %.1f M/s | locate %.2f G hits/s kernel %.2f ms (own samples %.2f G/s) | eub %.1f M rec/s' % (d['value']/1e6, r['kernel_ms'], r['frac'], d['e2e']['value']/1e6, d['e2e_utf8']['value']/1e6, l['value']/1e9, l['roofline']['kernel_ms'], l.get('own_samples_rank0',{}).get('hits_per_s',0)/1e9, e['value']/1e6))"
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.log
tail -3 gpurun_out/${TAG}_bench_n1.log; cut -c1-600 gpurun_out/${TAG}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sr-sweep > /dev/null 2> gpurun_out/${TAG}_launches.log
ncu --set full --clock-control none --import-source on -k regex:k_count -s 3 -c 1 -f -o gpurun_out/${TAG}_k_count \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lf > /dev/null 2> gpurun_out/${TAG}_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -f -o gpurun_out/${TAG}_k_locate \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sr-sweep > /dev/null 2> gpurun_out/${TAG}_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:k_extract -s 1 -c 1 -f -o gpurun_out/${TAG}_k_extract \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sr-sweep > /dev/null 2> gpurun_out/${TAG}_ncu3.log
ls -la gpurun_out/${TAG}_*.ncu-rep
