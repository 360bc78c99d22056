cd /root/repo
mkdir -p gpurun_out
python bench.py --build-only 2> gpurun_out/variants_build.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count or locate" 2>&1 | tail -2
for c in 2000000 500000 333334 250000 125000; do
  echo "== FMGPU_PIPE_CHUNK=$c"
  FMGPU_PIPE_CHUNK=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/pipe_$c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('   value %.1f M/s  step %.3f ms  kernel %.3f ms  frac %.3f  e2e %.1f M/s' % (d['value']/1e6, d['ms_per_step'], r['kernel_ms'], r['frac'], d['e2e']['value']/1e6))"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r01_launches_count_v3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_l.log
grep -E "k_" gpurun_out/r01_launches_count_v3.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200- | head -12
