#!/bin/bash
# usage (under gpurun): bash tools/gpu_quick.sh <tag> [pytest -k expr]   -> parity tests, count bench line (no cpu baseline, no LF legs)
TAG=${1:-quick}; KEXPR=${2:-"count or locate"}
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$KEXPR" 2>&1 | tail -3
python bench.py --build-only 2> gpurun_out/${TAG}_build.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-lf > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); r=d['roofline']
print('value %.1f M/s  step %.3f ms  kernel %.3f ms  frac %.3f  e2e %.1f M/s  hbm %.1f MB' % (d['value']/1e6, d['ms_per_step'], r['kernel_ms'], r['frac'], d['e2e']['value']/1e6, d['index']['hbm_bytes']/1e6))"
