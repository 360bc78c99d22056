cd /root/repo; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
bash tools/variants.sh "-DCOUNT_MIN_CTAS=4" "-DCOUNT_MIN_CTAS=5" "-DCOUNT_MIN_CTAS=6"
python tools/bench_lf.py --check 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('locate ms %.2f (%.2f G hits/s)  eub ms %.3f  extract ms %.3f  hbm %.0f MB' % (d['locate']['ms_per_step'], d['locate']['hits_per_s']/1e9, d['eub']['ms_per_step'], d['extract']['ms_per_step'], d['index_hbm_bytes']/1e6))"
