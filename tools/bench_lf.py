#!/usr/bin/env python
"""LF-walk workloads of BASELINE.json (configs[2] and [3]) on one GPU, device-resident inputs:

  locate  : n_pat patterns (substrings, len 4-64), max_hits per pattern  -> located hits/s
  eub     : extractUntilBoundary('\\n', dst 512) of the first n_eub located hits -> records/s, chars/s
  extract : n_ext random ranges of 32 chars (the reference's JMH extract workload)  -> chars/s

  python tools/bench_lf.py [--sample-rate 32] [--n-pat 1000000] [--max-hits 1000] [--steps 3]

One JSON line on stdout.  Parity of a sample of the results is checked against the CPU oracle.
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
    ap.add_argument("--skip", default="", help="comma list of workloads to skip: locate,eub,extract")
    args = ap.parse_args()
    skip = set(x for x in args.skip.split(",") if x)

    import torch
    from index4j_b200 import FmIndex
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    holder = {}
    blob = bench.get_index_blob(args.n_text, args.sample_rate, holder)
    chars, off = bench.get_patterns(args.n_text, args.n_pat, args.min_len, args.max_len, 42, holder)
    holder.clear()
    ix = FmIndex.read(blob, device=0)
    n_pat = off.size - 1
    out = {"sample_rate": args.sample_rate, "n_text": args.n_text, "index_hbm_bytes": ix.device_bytes()}

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    d_chars = torch.from_numpy(chars.view(np.int16)).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    d_n_hits = torch.empty(n_pat, dtype=torch.int32, device=dev)
    d_hit_off = torch.empty(n_pat + 1, dtype=torch.int64, device=dev)
    d_status = torch.empty(n_pat, dtype=torch.int32, device=dev)
    total = ix.locate_batch_device(d_chars, d_off, args.max_hits, d_n_hits, d_hit_off, None, d_status)
    d_pos = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    bench.log("locate: %d patterns -> %d hits (max %d per pattern)" % (n_pat, total, args.max_hits))

    oracle = None
    if args.check:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        oracle = pyoracle.OracleFmIndex(blob)

    if "locate" not in skip:
        ms = timed(lambda: ix.locate_batch_device(d_chars, d_off, args.max_hits, d_n_hits, d_hit_off, d_pos, d_status), args.steps, args.warmup)
        st = ix.last_stats()
        out["locate"] = {"patterns": n_pat, "max_hits": args.max_hits, "hits": total, "ms_per_step": ms, "hits_per_s": total / (ms / 1e3),
                         "lf_steps": st["lf_steps"], "lf_steps_per_s": st["lf_steps"] / (ms / 1e3), "lf_levels": st["lf_levels"],
                         "sampled_tests": st["sampled_tests"], "generic_ranks": st["ranks"], "generic_rank_levels": st["rank_levels"],
                         "launches": st["launches"]}
        # algorithmic 32-byte records: per LF step 1 sampled group + 1 block descriptor + per level (1 sector, +1 node record below
        # the root); per generic rank 1 cell + levels; per hit 1 SA record
        recs = st["sampled_tests"] + st["lf_steps"] + 2 * st["lf_levels"] - st["lf_steps"] + st["ranks"] + st["rank_levels"] + total
        out["locate"]["alg_bytes"] = 32.0 * recs
        out["locate"]["alg_gb_per_s"] = 32.0 * recs / (ms / 1e3) / 1e9
        if oracle is not None:
            k = min(args.check, n_pat)
            hit_off = d_hit_off[: k + 1].cpu().numpy().astype(np.int64)
            pos = d_pos[: int(hit_off[-1])].cpu().numpy()
            w_n, w_pos, _ = oracle.locate_batch(chars[: int(off[k])], off[: k + 1], args.max_hits, max(args.max_hits, 1), threads=os.cpu_count() or 1)
            assert np.array_equal(np.diff(hit_off), w_n), "locate hit counts differ from the oracle"
            for i in range(k):
                assert np.array_equal(pos[hit_off[i]: hit_off[i + 1]], w_pos[i, : w_n[i]]), "located positions differ from the oracle (pattern %d)" % i
            out["locate"]["oracle_checked_patterns"] = k

    if "eub" not in skip and total:
        n_eub = min(args.n_eub, total)
        if "locate" in skip:
            ix.locate_batch_device(d_chars, d_off, args.max_hits, d_n_hits, d_hit_off, d_pos, d_status)
        # hits spread over the whole batch (every total/n_eub-th hit), as cfg 4 takes "1M hits"
        sel = torch.linspace(0, total - 1, n_eub, device=dev, dtype=torch.float64).to(torch.int64)
        d_from = d_pos[sel].contiguous()
        d_arena = torch.empty((n_eub, args.dst_len), dtype=torch.int16, device=dev)
        d_len = torch.empty(n_eub, dtype=torch.int32, device=dev)
        d_st = torch.empty(n_eub, dtype=torch.int32, device=dev)
        ms = timed(lambda: ix.extract_until_boundary_batch_device(d_from, "\n", args.dst_len, 0, d_arena, d_len, d_st), args.steps, args.warmup)
        st = ix.last_stats()
        ln = d_len.cpu().numpy()
        stt = d_st.cpu().numpy()
        okc = int(ln[stt == 0].sum())
        out["eub"] = {"records": n_eub, "dst_len": args.dst_len, "ms_per_step": ms, "records_per_s": n_eub / (ms / 1e3),
                      "chars": okc, "chars_per_s": okc / (ms / 1e3), "status_nonzero": int((stt != 0).sum()),
                      "lf_steps": st["lf_steps"], "lf_steps_per_s": st["lf_steps"] / (ms / 1e3), "lf_levels": st["lf_levels"],
                      "generic_ranks": st["ranks"], "launches": st["launches"]}
        if oracle is not None:
            k = min(args.check, n_eub)
            frm = d_from[:k].cpu().numpy().astype(np.int32)
            w_arena, w_ln, w_st = oracle.extract_until_boundary_batch(frm, 10, args.dst_len, 0, threads=os.cpu_count() or 1)
            arena = d_arena[:k].cpu().numpy().view(np.uint16)
            assert np.array_equal(stt[:k], w_st) and np.array_equal(ln[:k][w_st == 0], w_ln[w_st == 0]), "extractUntilBoundary differs from the oracle"
            for i in range(k):
                if w_st[i] == 0:
                    assert np.array_equal(arena[i, : w_ln[i]], w_arena[i, : w_ln[i]]), "extracted record %d differs from the oracle" % i
            out["eub"]["oracle_checked_records"] = k

    if "extract" not in skip:
        rng = np.random.default_rng(7)
        n_ext = args.n_ext
        start = rng.integers(0, args.n_text - 64, n_ext).astype(np.int32)
        stop = (start + 32).astype(np.int32)
        aoff = (np.arange(n_ext + 1, dtype=np.int64) * 32)
        d_start, d_stop = torch.from_numpy(start).to(dev), torch.from_numpy(stop).to(dev)
        d_aoff = torch.from_numpy(aoff).to(dev)
        d_arena = torch.empty(n_ext * 32, dtype=torch.int16, device=dev)
        d_len = torch.empty(n_ext, dtype=torch.int32, device=dev)
        d_st = torch.empty(n_ext, dtype=torch.int32, device=dev)
        ms = timed(lambda: ix.extract_batch_device(d_start, d_stop, d_arena, d_aoff, d_len, d_st), args.steps, args.warmup)
        st = ix.last_stats()
        out["extract"] = {"ranges": n_ext, "chars_each": 32, "ms_per_step": ms, "ranges_per_s": n_ext / (ms / 1e3),
                          "chars_per_s": 32.0 * n_ext / (ms / 1e3), "lf_steps": st["lf_steps"], "lf_steps_per_s": st["lf_steps"] / (ms / 1e3)}
        if oracle is not None:
            k = min(args.check, n_ext)
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
