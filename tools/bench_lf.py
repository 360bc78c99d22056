#!/usr/bin/env python
"""LF-walk workloads of BASELINE.json (configs[2] and [3]) on one GPU, device-resident inputs:

  locate  : n_pat patterns (substrings, len 4-64), max_hits per pattern  -> located hits/s
  eub     : extractUntilBoundary('\\n', dst 512) of n_eub located hits (spread over the batch) -> records/s, chars/s
  extract : n_ext random ranges of 32 chars (the reference's JMH extract workload)  -> chars/s

  python tools/bench_lf.py [--sample-rate 32] [--n-pat 1000000] [--max-hits 1000] [--steps 3]

One JSON line on stdout.  A sample of the results is checked against the CPU oracle (--check N).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-text", type=int, default=1 << 30)
    ap.add_argument("--n-pat", type=int, default=1_000_000)
    ap.add_argument("--min-len", type=int, default=4)
    ap.add_argument("--max-len", type=int, default=64)
    ap.add_argument("--max-hits", type=int, default=1000)
    ap.add_argument("--sample-rate", type=int, default=32)
    ap.add_argument("--n-eub", type=int, default=1_000_000)
    ap.add_argument("--n-ext", type=int, default=1_000_000)
    ap.add_argument("--dst-len", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", type=int, default=300, help="patterns / records checked against the CPU oracle")
    ap.add_argument("--skip", default="", help="comma list of workloads to skip: eub,extract")
    args = ap.parse_args()
    skip = set(x for x in args.skip.split(",") if x)

    import torch
    from index4j_b200 import FmIndex, workloads
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    holder = {}
    blob = bench.get_index_blob(args.n_text, args.sample_rate, holder)
    chars, off = bench.get_patterns(args.n_text, args.n_pat, args.min_len, args.max_len, 42, holder)
    holder.clear()
    ix = FmIndex.read(blob, device=0)
    n_pat = off.size - 1
    out = {"sample_rate": args.sample_rate, "n_text": args.n_text, "index_hbm_bytes": ix.device_bytes(), "serialized_bytes": len(blob)}
    d_chars = torch.from_numpy(chars.view(np.int16)).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    oracle = None
    if args.check:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        oracle = pyoracle.OracleFmIndex(blob)
    threads = os.cpu_count() or 1

    out["locate"], d_hit_off, d_pos = workloads.locate_workload(ix, d_chars, d_off, args.max_hits, args.steps, args.warmup)
    total = out["locate"]["hits"]
    bench.log("locate: %d patterns -> %d hits (max %d per pattern), %.1f ms" % (n_pat, total, args.max_hits, out["locate"]["ms_per_step"]))
    if oracle is not None:
        k = min(args.check, n_pat)
        hit_off = d_hit_off[: k + 1].cpu().numpy().astype(np.int64)
        pos = d_pos[: int(hit_off[-1])].cpu().numpy()
        w_n, w_pos, _ = oracle.locate_batch(chars[: int(off[k])], off[: k + 1], args.max_hits, max(args.max_hits, 1), threads=threads)
        assert np.array_equal(np.diff(hit_off), w_n), "locate hit counts differ from the oracle"
        for i in range(k):
            assert np.array_equal(pos[hit_off[i]: hit_off[i + 1]], w_pos[i, : w_n[i]]), "located positions differ from the oracle (pattern %d)" % i
        out["locate"]["oracle_checked_patterns"] = k

    if "eub" not in skip and total:
        n_eub = min(args.n_eub, total)
        sel = torch.linspace(0, total - 1, n_eub, device=dev, dtype=torch.float64).to(torch.int64)  # hits spread over the whole batch
        d_from = d_pos[sel].contiguous()
        out["eub"], d_arena, d_len, d_st = workloads.eub_workload(ix, d_from, args.dst_len, args.steps, args.warmup)
        if oracle is not None:
            k = min(args.check, n_eub)
            frm = d_from[:k].cpu().numpy().astype(np.int32)
            w_arena, w_ln, w_st = oracle.extract_until_boundary_batch(frm, 10, args.dst_len, 0, threads=threads)
            arena = d_arena[:k].cpu().numpy().view(np.uint16)
            ln, stt = d_len[:k].cpu().numpy(), d_st[:k].cpu().numpy()
            assert np.array_equal(stt, w_st) and np.array_equal(ln[w_st == 0], w_ln[w_st == 0]), "extractUntilBoundary differs from the oracle"
            for i in range(k):
                if w_st[i] == 0:
                    assert np.array_equal(arena[i, : w_ln[i]], w_arena[i, : w_ln[i]]), "extracted record %d differs from the oracle" % i
            out["eub"]["oracle_checked_records"] = k

    if "extract" not in skip:
        out["extract"], start, stop, d_arena = workloads.extract_workload(ix, args.n_text, args.n_ext, 32, args.steps, args.warmup)
        if oracle is not None:
            k = min(args.check, args.n_ext)
            arena = d_arena[: k * 32].cpu().numpy().view(np.uint16).reshape(k, 32)
            for i in range(k):
                want = oracle.extract(int(start[i]), int(stop[i]), 32)
                assert np.array_equal(arena[i], np.asarray(want, dtype=np.uint16)[:32]), "extract %d differs from the oracle" % i
            out["extract"]["oracle_checked_ranges"] = k

    print(json.dumps(out), flush=True)
    ix.close()


if __name__ == "__main__":
    t0 = time.time()
    main()
    bench.log("bench_lf done in %.1fs" % (time.time() - t0))
