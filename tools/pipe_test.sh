#!/bin/bash
# usage (under gpurun): bash tools/pipe_test.sh   -> e2e count throughput vs chunk size / number of compute streams, and the PCIe copy rate
cd /root/repo
mkdir -p gpurun_out
python bench.py --build-only 2> gpurun_out/variants_build.log
python - <<'PY'
import torch, time
a = torch.empty(76_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty_like(a, device="cuda")
for _ in range(3): d.copy_(a, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(a, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("H2D 76 MB pinned: %.3f ms = %.1f GB/s" % (dt * 1e3, 76e6 / dt / 1e9))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count" 2>&1 | tail -2
for s in 1 2 3; do
for c in 500000 250000 125000; do
  echo "== FMGPU_PIPE_STREAMS=$s FMGPU_PIPE_CHUNK=$c"
  FMGPU_PIPE_STREAMS=$s FMGPU_PIPE_CHUNK=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-lf 2> gpurun_out/pipe_$c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('   value %.1f M/s  step %.3f ms  kernel %.3f ms  frac %.3f  e2e %.1f M/s' % (d['value']/1e6, d['ms_per_step'], r['kernel_ms'], r['frac'], d['e2e']['value']/1e6))"
done
done
