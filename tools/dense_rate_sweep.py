"""locate (max 1000 hits per pattern, 250 k patterns of the bench batch) against the device-side sample rate of the dense SA
samples (fmgpu_opts.locate_sample_rate): hits/s and HBM per rate.  python tools/dense_rate_sweep.py  (one GPU)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from index4j_b200 import FmIndex, workloads  # noqa: E402

n_text, n_pat = 1 << 30, 250_000
holder = {}
blob = bench.get_index_blob(n_text, 32, holder)
chars, off = bench.get_patterns(n_text, 1_000_000, 4, 64, 42, holder)
dev = torch.device("cuda", 0)
d_off = torch.from_numpy(off[: n_pat + 1].view(np.int64)).to(dev)
d_chars = torch.from_numpy(chars[: int(off[n_pat])].view(np.int16)).to(dev)
out = {}
for rate in (-1, 16, 8, 4, 2):
    ix = FmIndex.read(blob, device=0, locate_sample_rate=rate)
    leg, _, _ = workloads.locate_workload_nostats(ix, d_chars, d_off, 1000, 3, 1)
    out[str(ix.locate_sample_rate)] = {"hits_per_s": leg["hits"] / (leg["ms_per_step"] / 1e3), "ms_per_step": leg["ms_per_step"], "hits": leg["hits"],
                                       "dense_sample_bytes": ix.dense_sample_bytes(), "index_hbm_bytes": ix.device_bytes()}
    ix.close()
    torch.cuda.empty_cache()
print(json.dumps({"workload": "locate, max 1000 hits per pattern, first 250000 patterns of the bench batch, FmIndex(sampleRate 32) over 2^30 chars; "
                              "key = rate of the samples the walks end at (32 = the index's own)", "by_rate": out}))
