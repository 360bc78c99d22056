cd /root/repo
python bench.py --build-only 2>/dev/null
for g in 32 64 128; do
echo "== L2 fetch $g"
FMGPU_L2_FETCH=$g python tools/bench_lf.py --check 0 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('locate ms %.2f  eub ms %.3f  extract ms %.3f' % (d['locate']['ms_per_step'], d['eub']['ms_per_step'], d['extract']['ms_per_step']))"
FMGPU_L2_FETCH=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('count: value %.1f M/s kernel %.3f ms' % (d['value']/1e6, r['kernel_ms']))"
done
