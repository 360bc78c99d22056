#!/bin/bash
# BASELINE.json configs[2]: locate with max 1000 hits/pattern, sampleRate sweep 16/32/64 (LF-walk length vs index size)
cd /root/repo; mkdir -p gpurun_out
for sr in 16 32 64; do
  python tools/bench_lf.py --sample-rate $sr --check 200 > gpurun_out/r01_lf_sr$sr.json 2> gpurun_out/r01_lf_sr$sr.log
  tail -2 gpurun_out/r01_lf_sr$sr.log
  cat gpurun_out/r01_lf_sr$sr.json
done
