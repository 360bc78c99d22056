#!/usr/bin/env python
"""Timeline of one host-pointer count call (char[] and UTF-8) over the bench workload: FMGPU_PIPE_TRACE output + wall time."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from index4j_b200 import FmIndex  # noqa: E402

n_text, n_pat = 1 << 30, 1_000_000
holder = {}
blob = bench.get_index_blob(n_text, 32, holder)
chars, off = bench.get_patterns(n_text, n_pat, 4, 64, 42, holder)
ix = FmIndex.read(blob, device=0)
h_chars = torch.from_numpy(chars.view(np.int16)).pin_memory().numpy().view(np.uint16)
h_bytes = torch.from_numpy(chars.astype(np.uint8)).pin_memory().numpy()
h_off = torch.from_numpy(off.view(np.int64)).pin_memory().numpy().view(np.uint64)
h_counts = torch.empty(n_pat, dtype=torch.int32).pin_memory().numpy()
h_status = torch.empty(n_pat, dtype=torch.int32).pin_memory().numpy()
for name, fn in (("char[]", lambda: ix.count_batch_into(h_chars, h_off, h_counts, h_status)),
                 ("utf8", lambda: ix.count_batch_utf8_into(h_bytes, h_off, h_counts, h_status))):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    dt = (time.perf_counter() - t0) / 10
    print("%s: %.3f ms per call, %.1f M patterns/s" % (name, dt * 1e3, n_pat / dt / 1e6), flush=True)
    os.environ["FMGPU_PIPE_TRACE"] = "1"
    fn()
    del os.environ["FMGPU_PIPE_TRACE"]
    sys.stderr.flush()
