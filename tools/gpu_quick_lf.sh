#!/bin/bash
# usage (under gpurun): bash tools/gpu_quick_lf.sh <tag>   -> LF parity tests + LF workload line
TAG=${1:-qlf}
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "locate or extract" 2>&1 | tail -3
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
python tools/bench_lf.py --check 200 > gpurun_out/${TAG}_lf.json 2> gpurun_out/${TAG}_lf.log
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_lf.json'))
print('locate ms %.2f (%.2f G hits/s, %.1f G lf/s)  eub ms %.3f  extract ms %.3f  hbm %.0f MB' % (d['locate']['ms_per_step'], d['locate']['hits_per_s']/1e9, d['locate']['lf_steps_per_s']/1e9, d['eub']['ms_per_step'], d['extract']['ms_per_step'], d['index_hbm_bytes']/1e6))"
