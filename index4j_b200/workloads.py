"""Device-resident LF-walk workloads of BASELINE.json configs[2] / [3], shared by bench.py and tools/bench_lf.py.

Every function times `steps` passes with CUDA events on the current stream (after `warmup` passes) and returns
a dict; inputs and outputs stay in HBM.  Nothing here touches the oracle: parity spot checks live in the callers.
"""
from __future__ import annotations

import numpy as np
import torch


def _timed(fn, steps: int, warmup: int) -> float:
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


def _kernel_ms(ix, kind: int, steps: int) -> float:
    """mean device time of the last `steps` launches of one kernel kind (events recorded by the library on the launching stream)"""
    return float(np.mean([ix.kernel_ms(kind, i) for i in range(min(steps, 64))]))


def _counted(ix, fn) -> dict:
    """Work counters of one extra, untimed pass with the instrumented kernels (the timed passes run the production ones)."""
    ix.set_stats(True)
    try:
        fn()
        torch.cuda.synchronize()
        return ix.last_stats()
    finally:
        ix.set_stats(False)


def locate_workload(ix, d_chars, d_off, max_hits: int, steps: int, warmup: int):
    """FmIndex.locate over the whole batch (count kernels + hit scan + LF walks).  -> (stats dict, d_hit_off, d_pos)"""
    dev = d_chars.device
    n_pat = d_off.numel() - 1
    d_n_hits = torch.empty(n_pat, dtype=torch.int32, device=dev)
    d_hit_off = torch.empty(n_pat + 1, dtype=torch.int64, device=dev)
    d_status = torch.empty(n_pat, dtype=torch.int32, device=dev)
    total = ix.locate_batch_device(d_chars, d_off, max_hits, d_n_hits, d_hit_off, None, d_status)  # sizing pass
    d_pos = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    fn = lambda: ix.locate_batch_device(d_chars, d_off, max_hits, d_n_hits, d_hit_off, d_pos, d_status)  # noqa: E731
    ms = _timed(fn, steps, warmup)
    k_ms = _kernel_ms(ix, 1, steps)
    st = _counted(ix, fn)
    # algorithmic 32-byte records: per sampled-row test 1 group record; per LF step 1 block descriptor; per TWO wavelet levels
    # 1 level record + 1 node record; per generic rank 1 cell; per hit 1 SA record (+ 4 bytes read and 4 written per hit row)
    recs = st["sampled_tests"] + st["lf_steps"] + 2 * st["level_records"] + st["ranks"] + total
    out = {"patterns": n_pat, "max_hits": max_hits, "hits": int(total), "ms_per_step": ms, "hits_per_s": total / (ms / 1e3),
           "kernel_ms": k_ms, "kernel_alg_bytes": 32.0 * recs + 8.0 * total,
           "records": {"sampled_row_groups": st["sampled_tests"], "block_descriptors": st["lf_steps"], "level_records": st["level_records"],
                       "node_records": st["level_records"], "cells": st["ranks"], "sa_records": int(total)},
           "lf_steps": st["lf_steps"], "lf_steps_per_s": st["lf_steps"] / (ms / 1e3), "lf_levels": st["lf_levels"],
           "sampled_tests": st["sampled_tests"], "generic_ranks": st["ranks"], "launches": st["launches"],
           "alg_bytes": 32.0 * recs, "alg_gb_per_s": 32.0 * recs / (ms / 1e3) / 1e9}
    return out, d_hit_off, d_pos


def locate_workload_nostats(ix, d_chars, d_off, max_hits: int, steps: int, warmup: int):
    """locate_workload without the instrumented pass / kernel timing (the sampleRate sweep of bench.py)."""
    dev = d_chars.device
    n_pat = d_off.numel() - 1
    d_n_hits = torch.empty(n_pat, dtype=torch.int32, device=dev)
    d_hit_off = torch.empty(n_pat + 1, dtype=torch.int64, device=dev)
    total = ix.locate_batch_device(d_chars, d_off, max_hits, d_n_hits, d_hit_off, None, None)
    d_pos = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    fn = lambda: ix.locate_batch_device(d_chars, d_off, max_hits, d_n_hits, d_hit_off, d_pos, None)  # noqa: E731
    ms = _timed(fn, steps, warmup)
    out = {"patterns": n_pat, "hits": int(total), "ms_per_step": ms, "lf_steps_est": float(total) * (ix.locate_sample_rate - 1) / 2.0}
    return out, d_hit_off, d_pos


def eub_workload(ix, d_from, dst_len: int, steps: int, warmup: int, boundary="\n", mode: int = 0):
    """FmIndex.extractUntilBoundary for every position of d_from into an n x dst_len device arena."""
    dev = d_from.device
    n = d_from.numel()
    d_arena = torch.empty((n, dst_len), dtype=torch.int16, device=dev)
    d_len = torch.empty(n, dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    fn = lambda: ix.extract_until_boundary_batch_device(d_from, boundary, dst_len, mode, d_arena, d_len, d_st)  # noqa: E731
    ms = _timed(fn, steps, warmup)
    k_ms = _kernel_ms(ix, 2, steps)
    st = _counted(ix, fn)
    ok_chars = int(d_len[d_st == 0].sum().item())
    # algorithmic bytes of k_extract: per LF step 1 block descriptor, per TWO wavelet levels 1 level + 1 node record, per generic
    # rank 1 cell, per sample interval walked 1 inverse-SA record (~ LF steps / sampleRate, + 1 per record), 2 bytes per char written
    isa = st["lf_steps"] // max(ix.sample_rate, 1) + n
    recs = st["lf_steps"] + 2 * st["level_records"] + st["ranks"] + isa
    out = {"records": n, "dst_len": dst_len, "ms_per_step": ms, "records_per_s": n / (ms / 1e3), "chars": ok_chars,
           "kernel_ms": k_ms, "kernel_alg_bytes": 32.0 * recs + 2.0 * ok_chars + 12.0 * n,
           "records_read": {"block_descriptors": st["lf_steps"], "level_records": st["level_records"], "node_records": st["level_records"],
                            "cells": st["ranks"], "isa_records": isa},
           "chars_per_s": ok_chars / (ms / 1e3), "status_nonzero": int((d_st != 0).sum().item()), "lf_steps": st["lf_steps"],
           "lf_steps_per_s": st["lf_steps"] / (ms / 1e3), "lf_levels": st["lf_levels"], "generic_ranks": st["ranks"],
           "launches": st["launches"]}
    return out, d_arena, d_len, d_st


def records_workload(ix, d_from, dst_len: int, steps: int, warmup: int, boundary="\n"):
    """Fused locate -> extractUntilBoundary (fmgpu_extract_records_batch_device): per hit its record, every distinct record read once."""
    dev = d_from.device
    n = d_from.numel()
    d_idx = torch.empty(n, dtype=torch.int32, device=dev)
    d_len = torch.empty(n, dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    d_arena = torch.empty((n, dst_len), dtype=torch.int16, device=dev)
    box = {}

    def fn():
        box["n_rec"] = ix.extract_records_batch_device(d_from, boundary, dst_len, d_idx, d_len, d_st, d_arena)

    ms = _timed(fn, steps, warmup)
    out = {"hits": n, "distinct_records": box["n_rec"], "hits_per_record": n / max(box["n_rec"], 1), "dst_len": dst_len, "ms_per_step": ms,
           "hits_per_s": n / (ms / 1e3), "records_per_s": box["n_rec"] / (ms / 1e3)}
    return out, d_idx, d_len, d_st, d_arena


def extract_workload(ix, n_text: int, n_ext: int, chars_each: int, steps: int, warmup: int, seed: int = 7):
    """FmIndex.extract of n_ext random ranges of chars_each chars (the reference's JMH extract workload shape)."""
    dev = torch.device("cuda", ix.device)
    rng = np.random.default_rng(seed)
    start = rng.integers(0, n_text - 2 * chars_each, n_ext).astype(np.int32)
    stop = (start + chars_each).astype(np.int32)
    aoff = np.arange(n_ext + 1, dtype=np.int64) * chars_each
    d_start, d_stop, d_aoff = torch.from_numpy(start).to(dev), torch.from_numpy(stop).to(dev), torch.from_numpy(aoff).to(dev)
    d_arena = torch.empty(n_ext * chars_each, dtype=torch.int16, device=dev)
    d_len = torch.empty(n_ext, dtype=torch.int32, device=dev)
    d_st = torch.empty(n_ext, dtype=torch.int32, device=dev)
    fn = lambda: ix.extract_batch_device(d_start, d_stop, d_arena, d_aoff, d_len, d_st)  # noqa: E731
    ms = _timed(fn, steps, warmup)
    st = _counted(ix, fn)
    out = {"ranges": n_ext, "chars_each": chars_each, "ms_per_step": ms, "ranges_per_s": n_ext / (ms / 1e3),
           "chars_per_s": float(chars_each) * n_ext / (ms / 1e3), "lf_steps": st["lf_steps"], "lf_steps_per_s": st["lf_steps"] / (ms / 1e3)}
    return out, start, stop, d_arena
