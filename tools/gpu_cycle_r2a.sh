#!/bin/bash
# round 2, cycle A: GPU parity suite, the CPU reference arm (host-built index, native oracle build), the 1-GPU bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_reference.json 2> gpurun_out/r2a_reference.log
tail -3 gpurun_out/r2a_reference.log; cat gpurun_out/r2a_reference.json | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.log
tail -5 gpurun_out/r2a_bench_n1.log; cat gpurun_out/r2a_bench_n1.json | cut -c1-1500
