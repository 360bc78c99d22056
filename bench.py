#!/usr/bin/env python
"""bench.py — backward-search (count) throughput of the FM-index query engine.

Workload (BASELINE.json configs[1]): FmIndex sampleRate=32 over 1 GiB (2^30 chars) of synthetic
log-like text; 1,000,000 count queries, substrings of the text of length 4-64 (the reference's JMH
workload shape, indices/src/jmh/java/com/dynatrace/fm/FmIndexThroughputState.java:76-83).
A step = one pass of FmIndex.count over the whole 1 M-pattern batch.  With N GPUs the index is
replicated and every rank runs its own 1 M-pattern batch (weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA engine
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (C++ oracle port,
                                                                 # all host threads; no JVM exists in this image)
One JSON line on stdout (rank 0).  See DESIGN.md §5 for how every field is measured.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CACHE = os.path.join(ROOT, ".index_cache")
ALT_CACHE = "/tmp/index_cache"  # where indexes built in the development container are kept (outside the repo snapshot)

METRIC = "backward-search patterns/sec (count)"
UNIT = "patterns/s"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else that writes to fd 1 (e.g. NCCL's version banner) is sent to
    stderr, and emit() writes to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def cache_path(name: str) -> str:
    for d in (CACHE, ALT_CACHE):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return p
    os.makedirs(CACHE, exist_ok=True)
    return os.path.join(CACHE, name)


def get_text(n: int) -> np.ndarray:
    from index4j_b200.builder import gen_log_text
    return gen_log_text(n)


def get_index_blob(n: int, sr: int, text_holder: dict, use_gpu_sa: bool = True) -> bytes:
    """Serialized FmIndex (reference layout) of the synthetic text; built once per box and cached."""
    p = cache_path("log_n%d_sr%d.fmi" % (n, sr))
    if os.path.exists(p):
        with open(p, "rb") as fh:
            return fh.read()
    from index4j_b200.builder import build_index
    t0 = time.time()
    text = text_holder.get("text")
    if text is None:
        text = text_holder["text"] = get_text(n)
    blob = None
    if use_gpu_sa:
        import torch
        if torch.cuda.is_available():
            # suffix array, BWT and sampled structures on the GPU (index4j_b200/gpu_sa.py, csrc/kernels_build.cuh); the host encodes
            # the wavelet structure / RRR vector and serializes
            from index4j_b200.gpu_sa import build_index_gpu
            blob = build_index_gpu(text, sr, True, framed=False, verbose=True)
            torch.cuda.empty_cache()
    if blob is None:
        blob = build_index(text, sr, True, framed=False, verbose=True)
    tmp = p + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as fh:
        fh.write(blob)
    os.replace(tmp, p)
    log("index built in %.1fs (%d bytes) -> %s" % (time.time() - t0, len(blob), p))
    return blob


def get_patterns(n: int, n_pat: int, lo: int, hi: int, seed: int, text_holder: dict):
    p = cache_path("pat_n%d_%d_%d_%d_s%d.npz" % (n, n_pat, lo, hi, seed))
    if os.path.exists(p):
        z = np.load(p)
        return z["chars"], z["off"]
    from index4j_b200.builder import gen_patterns
    text = text_holder.get("text")
    if text is None:
        text = text_holder["text"] = get_text(n)
    chars, off = gen_patterns(text, n_pat, lo, hi, seed)
    tmp = p + ".tmp%d.npz" % os.getpid()
    np.savez(tmp, chars=chars, off=off)
    os.replace(tmp, p)
    return chars, off


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ["clocks.sm", "clocks.max.sm", "clocks_event_reasons.hw_slowdown", "clocks_event_reasons.hw_thermal_slowdown",
              "clocks_event_reasons.sw_thermal_slowdown", "clocks_event_reasons.sw_power_cap"]

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + ",".join(self.FIELDS),
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as fh:
                d = json.load(fh)
            v = d.get("hbm_gbs")
            if v:
                return float(v), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_oracle(blob: bytes):
    """The CPU arm's engine: the C++ restatement of the reference's Java loops, compiled ON this box with -march=native and
    without its work counters (oracle/Makefile `native`)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    build = pyoracle.use_native()
    return pyoracle.OracleFmIndex(blob), build


def cpu_count_throughput(o, chars, off, n_sample: int, threads: int, repeats: int = 1):
    """FmIndex.count of the first n_sample patterns on `threads` host threads (one shared immutable index, contiguous slices of
    the batch per thread — one @ThreadSafe FmIndex shared by a Java thread pool).  -> (patterns/s, seconds, counts)"""
    n_sample = min(n_sample, off.size - 1)
    sub_off = off[: n_sample + 1]
    sub_chars = chars[: int(sub_off[-1])]
    best = None
    counts = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        counts, _ = o.count_batch(sub_chars, sub_off, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_sample / best, best, counts


def cpu_thread_scaling(o, chars, off, n_sample: int, max_threads: int) -> dict:
    """patterns/s of the CPU arm at 1 / 8 / 16 / all host threads on a small sample (how the baseline scales on this box)."""
    out = {}
    for th in sorted({1, 8, 16, max_threads}):
        if th > max_threads:
            continue
        n = min(n_sample if th > 1 else max(n_sample // 8, 1000), off.size - 1)
        v, _, _ = cpu_count_throughput(o, chars, off, n, th)
        out[str(th)] = v
    return out


def cpu_lf_arms(o, chars, off, args, threads: int, frm) -> dict:
    """CPU arms of the locate and extractUntilBoundary legs on bounded samples (same engine, all host threads)."""
    k = min(args.cpu_locate_sample, off.size - 1)
    t0 = time.perf_counter()
    w_n, w_pos, _ = o.locate_batch(chars[: int(off[k])], off[: k + 1], args.max_hits, max(args.max_hits, 1), threads=threads)
    dt_loc = time.perf_counter() - t0
    m = min(args.cpu_eub_sample, frm.size)
    t0 = time.perf_counter()
    w_arena, w_ln, w_st = o.extract_until_boundary_batch(frm[:m], 10, args.dst_len, 0, threads=threads)
    dt_eub = time.perf_counter() - t0
    return {"locate": {"value": float(w_n.sum()) / dt_loc, "unit": "hits/s", "cores": threads, "kind": "port",
                       "sample": "locate(max %d) of the first %d patterns: %d hits in %.2fs" % (args.max_hits, k, int(w_n.sum()), dt_loc)},
            "extract_until_boundary": {"value": m / dt_eub, "unit": "records/s", "cores": threads, "kind": "port",
                                       "sample": "the first %d of the located records, dst %d: %.2fs" % (m, args.dst_len, dt_eub)},
            "_check": (k, w_n, w_pos, m, w_arena, w_ln, w_st)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores — the C++ restatement of
    its Java loops (no JVM exists in this image), all host threads.  Nothing of the GPU engine runs here: the index is built by
    the host producer."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    holder = {}
    blob = get_index_blob(args.n_text, args.sample_rate, holder, use_gpu_sa=False)
    chars, off = get_patterns(args.n_text, args.n_pat, args.min_len, args.max_len, 42, holder)
    holder.clear()
    threads = os.cpu_count() or 1
    o, build = load_oracle(blob)
    n_sample = min(args.ref_sample, off.size - 1)
    sub_off = off[: n_sample + 1]
    sub_chars = chars[: int(sub_off[-1])]
    for _ in range(args.warmup):
        o.count_batch(sub_chars, sub_off, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.count_batch(sub_chars, sub_off, threads=threads)
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    scaling = cpu_thread_scaling(o, chars, off, 100_000, threads)
    sample = "first %d of the %d patterns per step (C++ restatement of the Java loops, %s; no JVM in this image)" % (n_sample, args.n_pat, build)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "build": build,
                         "patterns_per_s_by_threads": scaling},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def run_sharded(args):
    """--mode sharded (BASELINE.json configs[4]): the text is split into one FmIndex shard per GPU with pattern-length overlap;
    count = per-shard counts summed over NCCL, locate = owned hits of every shard exchanged over NCCL and merged in rank order
    (index4j_b200/sharded.py, csrc/kernels_shard.cuh).  Launch under torchrun like the default mode."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_sharded
    ns = argparse.Namespace(shard_chars=args.shard_chars, n_pat=args.sharded_n_pat, min_len=8, max_len=args.max_len, max_hits=args.sharded_max_hits,
                            sample_rate=args.sample_rate, steps=max(args.steps, 1), verify=args.verify)
    res = bench_sharded.run(ns)
    if res is None:
        return
    out = {"metric": "sharded index: " + METRIC + " and located hits/sec", "value": res["count"]["patterns_per_s"], "unit": UNIT,
           "n_gpus": res["n_gpus"], "steps": ns.steps, "warmup": 1, "ms_per_step": res["count"]["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": {"workload": res["workload"], "index": "one shard per GPU", "text_chars": res["text_chars"]},
           "locate": res["locate"], "sharded": res}
    emit(out)


def workload_config(args):
    return {"workload": "count: %d patterns len %d-%d (substrings of the text) per GPU per step over FmIndex(sampleRate=%d) of %d chars of synthetic log text"
                        % (args.n_pat, args.min_len, args.max_len, args.sample_rate, args.n_text),
            "index": "replicated per GPU", "batch_per_gpu": args.n_pat,
            "l2": "index (>0.7 GB) + inputs larger than the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-text", type=int, default=1 << 30)
    ap.add_argument("--n-pat", type=int, default=1_000_000)
    ap.add_argument("--min-len", type=int, default=4)
    ap.add_argument("--max-len", type=int, default=64)
    ap.add_argument("--sample-rate", type=int, default=32)
    ap.add_argument("--ref-sample", type=int, default=1_000_000, help="patterns per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="patterns of the cpu_baseline leg")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--cpu-locate-sample", type=int, default=20_000, help="patterns of the CPU arm of the locate leg")
    ap.add_argument("--cpu-eub-sample", type=int, default=50_000, help="records of the CPU arm of the extractUntilBoundary leg")
    ap.add_argument("--no-sr-sweep", action="store_true", help="skip the sampleRate 16 / 64 locate legs (BASELINE.json configs[2])")
    ap.add_argument("--sweep-patterns", type=int, default=250_000, help="patterns of the sampleRate-sweep locate legs")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling legs of an N > 1 run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lf", action="store_true", help="skip the locate / extractUntilBoundary legs (extra keys of the JSON line)")
    ap.add_argument("--max-hits", type=int, default=1000)
    ap.add_argument("--lf-steps", type=int, default=2)
    ap.add_argument("--n-eub", type=int, default=1_000_000)
    ap.add_argument("--dst-len", type=int, default=512)
    ap.add_argument("--mode", default="replicated", choices=["replicated", "sharded"],
                    help="sharded = BASELINE.json configs[4]: one FmIndex shard of --shard-chars chars per GPU (text beyond 2^31 chars), "
                         "counts all-reduced, located positions exchanged over NCCL")
    ap.add_argument("--shard-chars", type=int, default=1 << 30)
    ap.add_argument("--sharded-n-pat", type=int, default=200_000)
    ap.add_argument("--sharded-max-hits", type=int, default=100)
    ap.add_argument("--verify", type=int, default=20000, help="sharded mode: located hits read back with extract")
    ap.add_argument("--build-only", action="store_true", help="build + cache the index and the pattern batches, then exit")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "sharded":
        return run_sharded(args)

    import torch
    import torch.distributed as dist
    from index4j_b200 import FmIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side waiting (an NCCL barrier is a kernel that spins on the GPU)

    def barrier():
        if world > 1:
            dist.barrier()

    holder = {}
    sweep_rates = [] if (args.no_sr_sweep or args.no_lf) else [sr for sr in (16, 64) if sr != args.sample_rate]
    if rank == 0:
        get_index_blob(args.n_text, args.sample_rate, holder)  # build + cache once per box
        for sr in sweep_rates:
            get_index_blob(args.n_text, sr, holder)
        for r in range(world):
            get_patterns(args.n_text, args.n_pat, args.min_len, args.max_len, 42 + r, holder)
    barrier()
    if args.build_only:
        return
    blob = get_index_blob(args.n_text, args.sample_rate, holder)
    chars, off = get_patterns(args.n_text, args.n_pat, args.min_len, args.max_len, 42 + rank, holder)
    holder.clear()
    t0 = time.time()
    ix = FmIndex.read(blob, device=local)
    log("rank %d: index on device in %.1fs, %.1f MB HBM, layout %s" % (rank, time.time() - t0, ix.device_bytes() / 1e6, ix.layout_bytes()))
    n_pat = off.size - 1

    # resident inputs / outputs for the kernel-level number
    d_chars = torch.from_numpy(chars.view(np.int16)).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    d_counts = torch.empty(n_pat, dtype=torch.int32, device=dev)
    d_status = torch.empty(n_pat, dtype=torch.int32, device=dev)
    ix.set_timing(True)

    def step():
        ix.count_batch_device(d_chars, d_off, d_counts, d_status)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    kernel_ms = [ix.search_kernel_ms(i) for i in range(min(args.steps, 64))]
    launches_per_step = ix.last_stats()["launches"]
    # the timed steps run the production kernel; its work counters (deterministic per index + batch) come from one
    # extra untimed pass of the instrumented variant
    ix.set_stats(True)
    step()
    torch.cuda.synchronize()
    stats = ix.last_stats()
    stats["launches"] = launches_per_step
    ix.set_stats(False)
    clocks = sampler.stop() if rank == 0 else None
    # the reference's "non-indexed" JMH workload (jmh/.../FmIndexThroughputState.java:106-112): random a-z strings of the same
    # lengths — the early-exit path (the range empties after a few steps); reported separately, rank 0's batch only
    rng_ni = np.random.default_rng(4242 + rank)
    ni_len = rng_ni.integers(args.min_len, args.max_len + 1, n_pat)
    ni_off = np.zeros(n_pat + 1, dtype=np.int64)
    ni_off[1:] = np.cumsum(ni_len)
    ni_chars = rng_ni.integers(ord("a"), ord("z") + 1, int(ni_off[-1])).astype(np.int16)
    d_ni_chars, d_ni_off = torch.from_numpy(ni_chars).to(dev), torch.from_numpy(ni_off).to(dev)
    d_ni_counts = torch.empty(n_pat, dtype=torch.int32, device=dev)
    for _ in range(2):
        ix.count_batch_device(d_ni_chars, d_ni_off, d_ni_counts, None)
    torch.cuda.synchronize()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record()
    for _ in range(args.steps):
        ix.count_batch_device(d_ni_chars, d_ni_off, d_ni_counts, None)
    n1.record()
    torch.cuda.synchronize()
    non_indexed = {"workload": "count of %d random a-z strings len %d-%d (the reference's non-indexed JMH batch: early exit)" % (n_pat, args.min_len, args.max_len),
                   "patterns_per_s_rank0": n_pat * args.steps / (n0.elapsed_time(n1) / 1e3), "ms_per_step": n0.elapsed_time(n1) / args.steps,
                   "patterns_found": int((d_ni_counts > 0).sum().item())}
    del d_ni_chars, d_ni_off, d_ni_counts
    # end to end through the C ABI with pinned HOST buffers (H2D + kernels + D2H inside the timed region)
    p_chars = torch.from_numpy(chars.view(np.int16)).pin_memory()
    p_off = torch.from_numpy(off.view(np.int64)).pin_memory()
    p_counts = torch.empty(n_pat, dtype=torch.int32).pin_memory()
    p_status = torch.empty(n_pat, dtype=torch.int32).pin_memory()
    h_chars, h_off = p_chars.numpy().view(np.uint16), p_off.numpy().view(np.uint64)
    h_counts, h_status = p_counts.numpy(), p_status.numpy()
    for _ in range(2):
        ix.count_batch_into(h_chars, h_off, h_counts, h_status)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.count_batch_into(h_chars, h_off, h_counts, h_status)
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(h_counts, d_counts.cpu().numpy())
    # the same call with the packed transport switched off (the caller's pinned arrays cross PCIe as they are)
    from index4j_b200.fm_index import native
    e2e_direct_s = None
    pack_threads = int(native().fmgpu_host_pack_threads())
    if pack_threads > 0 and os.environ.get("FMGPU_HOST_PACK", "1") != "0":
        os.environ["FMGPU_HOST_PACK"] = "0"
        try:
            ix.count_batch_into(h_chars, h_off, h_counts, h_status)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ix.count_batch_into(h_chars, h_off, h_counts, h_status)
            e2e_direct_s = time.perf_counter() - t0
        finally:
            del os.environ["FMGPU_HOST_PACK"]

    # the same batch as UTF-8 byte patterns through fmgpu_count_batch_utf8 (the reference's convertBytePatternToCharPattern +
    # count, FmIndex.java:239-298): 1 byte per char crosses PCIe, decoding happens in the device pre-pass
    utf8 = None
    if int(chars.max(initial=0)) < 128:
        p_bytes = torch.from_numpy(chars.astype(np.uint8)).pin_memory()
        h_bytes = p_bytes.numpy()
        p_counts8 = torch.empty(n_pat, dtype=torch.int32).pin_memory()
        h_counts8 = p_counts8.numpy()
        for _ in range(2):
            ix.count_batch_utf8_into(h_bytes, h_off, h_counts8, h_status)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ix.count_batch_utf8_into(h_bytes, h_off, h_counts8, h_status)
        utf8 = {"s": time.perf_counter() - t0, "h2d": int(h_bytes.nbytes + off.nbytes)}
        assert np.array_equal(h_counts8, h_counts)

    # second half of the metric: located hits/s (BASELINE.json configs[2]) and extractUntilBoundary of located hits
    # (configs[3]); same patterns, device-resident, timed with CUDA events; max over ranks, hits summed over ranks
    lf = None
    if not args.no_lf:
        from index4j_b200 import workloads
        loc, d_hit_off, d_pos = workloads.locate_workload(ix, d_chars, d_off, args.max_hits, args.lf_steps, 1)
        total_hits = loc["hits"]
        # the same call with the walks ending at the serialized index's OWN samples (the device-side dense samples switched off):
        # identical positions, (sampleRate - 1) / 2 LF steps per hit instead of (dense rate - 1) / 2
        loc_own = None
        if ix.dense_sample_bytes():
            ix.set_locate_dense(False)
            loc_own, _, d_pos_own = workloads.locate_workload(ix, d_chars, d_off, args.max_hits, args.lf_steps, 1)
            ix.set_locate_dense(True)
            assert torch.equal(d_pos_own, d_pos), "dense samples changed the located positions"
            del d_pos_own
        n_eub = min(args.n_eub, total_hits)
        sel = torch.linspace(0, max(total_hits - 1, 0), max(n_eub, 1), device=dev, dtype=torch.float64).to(torch.int64)
        d_from = d_pos[sel].contiguous()
        eub, d_arena, d_len, d_st = workloads.eub_workload(ix, d_from, args.dst_len, args.lf_steps, 1)
        # the same hits through the fused locate -> extractUntilBoundary path (distinct records read once), and a clustered
        # selection (the FIRST n_eub hits of the batch: all hits of the first patterns) where hits share records more often
        rec, r_idx, r_len, r_st, r_arena = workloads.records_workload(ix, d_from, args.dst_len, args.lf_steps, 1)
        assert torch.equal(r_st, d_st) and torch.equal(r_len[d_st == 0], d_len[d_st == 0]), "fused record path differs from extractUntilBoundary"
        samp = torch.arange(0, n_eub, max(n_eub // 20000, 1), device=dev)
        samp = samp[(d_st[samp] == 0) & (d_len[samp] > 0)]
        cols = torch.arange(args.dst_len, device=dev)[None, :]
        m_ok = cols < d_len[samp][:, None]
        assert torch.equal(r_arena[r_idx[samp].long()][m_ok], d_arena[samp][m_ok]), "fused record contents differ from extractUntilBoundary"
        del r_arena, r_idx, r_len, r_st
        rec_first, *_ = workloads.records_workload(ix, d_pos[:n_eub].contiguous(), args.dst_len, args.lf_steps, 1)
        del _
        torch.cuda.empty_cache()
        # end to end through the C ABI: host patterns in, host positions out (one call: ranges, hit scan, LF walks, D2H)
        p_nh = torch.empty(n_pat, dtype=torch.int32).pin_memory().numpy()
        p_ho = torch.empty(n_pat + 1, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
        p_pos = torch.empty(max(total_hits, 1), dtype=torch.int32).pin_memory().numpy()
        ix.locate_batch_into(h_chars, h_off, args.max_hits, p_nh, p_ho, p_pos, h_status)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.lf_steps):
            ix.locate_batch_into(h_chars, h_off, args.max_hits, p_nh, p_ho, p_pos, h_status)
        loc_e2e_s = (time.perf_counter() - t0) / args.lf_steps
        assert int(p_ho[-1]) == total_hits and np.array_equal(p_pos[:4096], d_pos[:4096].cpu().numpy())
        lf = {"rec": rec, "rec_first": rec_first, "loc": loc, "loc_own": loc_own, "eub": eub, "loc_e2e_ms": loc_e2e_s * 1e3, "d_from": d_from, "d_arena": d_arena, "d_len": d_len, "d_st": d_st,
              "d_hit_off": d_hit_off, "d_pos": d_pos, "h2d": int(chars.nbytes + off.nbytes), "d2h": int(p_nh.nbytes + p_ho.nbytes + p_pos.nbytes + h_status.nbytes)}

    # BASELINE.json configs[2]: locate (max 1000 hits per pattern) over sampleRate 16 / 32 / 64 — LF-walk length against index
    # size; every rank runs its own patterns over its replica (query-sharded), hits summed over ranks, time = max over ranks
    sweep = {}
    if lf and sweep_rates:
        ns = min(args.sweep_patterns, n_pat)
        sw_off = d_off[: ns + 1].contiguous()
        sw_chars = d_chars[: int(off[ns])].contiguous()
        ix.set_timing(False)
        def sweep_leg(ix_x, blob_len):
            # own samples (what the sweep is about: LF-walk length against index size), then with the device-side dense samples
            ix_x.set_locate_dense(False)
            own, _, _ = workloads.locate_workload_nostats(ix_x, sw_chars, sw_off, args.max_hits, args.lf_steps, 1)
            ix_x.set_locate_dense(True)
            d = dict(own, hbm_bytes=ix_x.device_bytes() - ix_x.dense_sample_bytes(), serialized_bytes=blob_len, dense_ms_per_step=0.0)
            if ix_x.dense_sample_bytes():
                den, _, _ = workloads.locate_workload_nostats(ix_x, sw_chars, sw_off, args.max_hits, args.lf_steps, 1)
                d.update(dense_ms_per_step=den["ms_per_step"], dense_rate=ix_x.locate_sample_rate, dense_bytes=ix_x.dense_sample_bytes())
            return d
        sweep[args.sample_rate] = sweep_leg(ix, len(blob))
        for sr in sweep_rates:
            blob_sr = get_index_blob(args.n_text, sr, holder)
            ix_sr = FmIndex.read(blob_sr, device=local)
            sweep[sr] = sweep_leg(ix_sr, len(blob_sr))
            ix_sr.close()
            del blob_sr
            torch.cuda.empty_cache()
        ix.set_timing(True)
        sw_t = torch.tensor([[sweep[sr]["ms_per_step"], sweep[sr]["dense_ms_per_step"]] for sr in sorted(sweep)], dtype=torch.float64, device=dev)
        sw_h = torch.tensor([[float(sweep[sr]["hits"]), float(sweep[sr]["lf_steps_est"])] for sr in sorted(sweep)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sw_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(sw_h, op=dist.ReduceOp.SUM)
        for k, sr in enumerate(sorted(sweep)):
            sweep[sr]["ms_per_step_max_over_ranks"] = float(sw_t[k, 0])
            sweep[sr]["hits_all_ranks"] = float(sw_h[k, 0])
            sweep[sr]["hits_per_s"] = float(sw_h[k, 0]) / (float(sw_t[k, 0]) / 1e3)
            sweep[sr]["dense_hits_per_s"] = float(sw_h[k, 0]) / (float(sw_t[k, 1]) / 1e3) if float(sw_t[k, 1]) > 0 else None

    # strong scaling (BASELINE.json configs[1] as written: the ONE 1 M-pattern batch sharded over the GPUs).  (a) device-resident:
    # rank r searches slice r of rank 0's batch, time = max over ranks; (b) one process: rank 0 alone drives all N GPUs through
    # ONE fmgpu_index replicated by the library (fmgpu_opts.devices) with host buffers — what a single JVM would call.
    strong = None
    if world > 1 and not args.no_strong:
        chars0, off0 = get_patterns(args.n_text, args.n_pat, args.min_len, args.max_len, 42, holder)
        lo, hi = n_pat * rank // world, n_pat * (rank + 1) // world
        s_off = off0[lo: hi + 1] - off0[lo]
        s_chars = chars0[int(off0[lo]): int(off0[hi])]
        ds_chars = torch.from_numpy(s_chars.view(np.int16)).to(dev)
        ds_off = torch.from_numpy(s_off.view(np.int64)).to(dev)
        ds_counts = torch.empty(hi - lo, dtype=torch.int32, device=dev)
        for _ in range(args.warmup):
            ix.count_batch_device(ds_chars, ds_off, ds_counts, None)
        torch.cuda.synchronize()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            ix.count_batch_device(ds_chars, ds_off, ds_counts, None)
        s1.record()
        torch.cuda.synchronize()
        barrier()
        strong = {"dev_ms": s0.elapsed_time(s1), "one_ms": 0.0, "one_utf8_ms": 0.0}
        dist.barrier(group=cpu_group)  # the other ranks now wait on the HOST until rank 0 is done: their GPUs stay idle
        if rank == 0:
            t_load = time.time()
            mix = FmIndex.read(blob, devices=list(range(world)))
            log("strong: one handle on devices %s in %.1fs" % (mix.devices, time.time() - t_load))
            from index4j_b200.fm_index import pinned_empty
            m_chars = pinned_empty(chars0.size, np.uint16)
            m_chars[:] = chars0
            m_off = pinned_empty(off0.size, np.uint64)
            m_off[:] = off0
            m_counts = pinned_empty(n_pat, np.int32)
            m_status = pinned_empty(n_pat, np.int32)
            for _ in range(3):
                mix.count_batch_into(m_chars, m_off, m_counts, m_status)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                mix.count_batch_into(m_chars, m_off, m_counts, m_status)
            strong["one_ms"] = (time.perf_counter() - t0) * 1e3
            if rank == 0 and world > 1:
                ref_counts = np.empty(n_pat, dtype=np.int32)
                ix.count_batch_into(np.ascontiguousarray(chars0), np.ascontiguousarray(off0), ref_counts, None)
                assert np.array_equal(ref_counts, m_counts), "multi-device counts differ from the single-device counts"
            if int(chars0.max(initial=0)) < 128:
                m_bytes = pinned_empty(chars0.size, np.uint8)
                m_bytes[:] = chars0.astype(np.uint8)
                for _ in range(3):
                    mix.count_batch_utf8_into(m_bytes, m_off, m_counts, m_status)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    mix.count_batch_utf8_into(m_bytes, m_off, m_counts, m_status)
                strong["one_utf8_ms"] = (time.perf_counter() - t0) * 1e3
            strong["h2d"] = int(chars0.nbytes + off0.nbytes)
            mix.close()
        dist.barrier(group=cpu_group)

    times = torch.tensor([ms_total, e2e_s * 1e3, statistics.mean(kernel_ms), lf["loc"]["ms_per_step"] if lf else 0.0,
                          lf["eub"]["ms_per_step"] if lf else 0.0, lf["loc_e2e_ms"] if lf else 0.0, utf8["s"] * 1e3 if utf8 else 0.0,
                          lf["loc"]["kernel_ms"] if lf else 0.0, lf["eub"]["kernel_ms"] if lf else 0.0, strong["dev_ms"] if strong else 0.0],
                         dtype=torch.float64, device=dev)
    sums = torch.tensor([lf["loc"]["hits"] if lf else 0, lf["eub"]["records"] if lf else 0, lf["eub"]["chars"] if lf else 0,
                         lf["loc"]["lf_steps"] if lf else 0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, kern_ms, loc_ms, eub_ms, loc_e2e_ms, utf8_ms, loc_k_ms, eub_k_ms, strong_dev_ms = [float(x) for x in times.cpu()]
    all_hits, all_records, all_chars, all_lf_steps = [float(x) for x in sums.cpu()]

    if rank == 0:
        value = world * n_pat * args.steps / (ms_total / 1e3)
        e2e_value = world * n_pat * args.steps / (e2e_ms / 1e3)
        peak, peak_src = hbm_peak()
        # algorithmic bytes of one k_count launch: every rank query touches one 32-byte sector for its (block, symbol) cell and at
        # most one 32-byte occurrence record; plus the pattern chars and descriptors it streams.
        alg_bytes = 32.0 * (stats["ranks"] + stats["level_records"]) + 2.0 * chars.size + 16.0 * n_pat + 8.0 * n_pat
        achieved = alg_bytes / (kern_ms / 1e3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(chars.size + 4 * (off.size + 8)) if e2e_direct_s is not None else int(chars.nbytes + off.nbytes),
                    "caller_input_bytes_per_step": int(chars.nbytes + off.nbytes),
                    "d2h_bytes_per_step": int(h_counts.nbytes + h_status.nbytes),
                    "transport": ("packed: %d host threads narrow the Latin-1 char[] to bytes + uint32 offsets into pinned staging (%d bytes cross PCIe per step)"
                                  % (pack_threads, int(chars.size + 4 * (off.size + 8))) if e2e_direct_s is not None
                                  else "direct: the caller's pinned arrays cross PCIe as they are"),
                    "direct_rank0": ({"value": n_pat * args.steps / e2e_direct_s, "unit": UNIT, "what": "the same call with FMGPU_HOST_PACK=0"}
                                     if e2e_direct_s is not None else None)},
            "e2e_utf8": ({"value": world * n_pat * args.steps / (utf8_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": utf8["h2d"],
                          "d2h_bytes_per_step": int(h_counts.nbytes + h_status.nbytes),
                          "call": "fmgpu_count_batch_utf8: the same patterns as UTF-8 bytes (1 byte per char), decoded on the device"}
                         if utf8 else None),
            "gpu_launches": int(stats["launches"]) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "k_count", "kernel_ms": kern_ms, "peak_source": peak_src,
                         "alg_bytes_per_launch": alg_bytes,
                         "alg_bytes_formula": "32 * (ranks + occurrence records) + pattern chars / descriptors: per rank query one 32-byte sector for its "
                                              "8-byte (block, symbol) cell and at most one 32-byte occurrence record (none for absent / run symbols) "
                                              "(DESIGN.md section 5)",
                         "ranks_per_launch": stats["ranks"], "levels_per_launch": stats["rank_levels"],
                         "level_records_per_launch": stats["level_records"],
                         "records_loaded_per_launch": stats["search_records_loaded"],
                         "records_per_rank": (stats["ranks"] + stats["level_records"]) / max(1, stats["ranks"]),
                         "ranks_per_s": stats["ranks"] / (kern_ms / 1e3),
                         "occurrence_records_per_launch": stats["level_records"]},
            "index": {"hbm_bytes": ix.device_bytes(), "layout_bytes": ix.layout_bytes(), "serialized_bytes": len(blob),
                      "start_table_q": ix.start_table_q(), "sample_rate": ix.sample_rate, "locate_sample_rate": ix.locate_sample_rate,
                      "dense_sample_bytes": ix.dense_sample_bytes()},
        }
        out["count_non_indexed"] = non_indexed
        if lf:
            out["locate"] = {"metric": "located hits/sec (max %d hits per pattern)" % args.max_hits, "value": all_hits / (loc_ms / 1e3),
                             "unit": "hits/s", "hits_per_step": all_hits, "ms_per_step": loc_ms, "steps": args.lf_steps,
                             "lf_steps_per_s": all_lf_steps / (loc_ms / 1e3),
                             "e2e": {"value": all_hits / (loc_e2e_ms / 1e3), "unit": "hits/s", "h2d_bytes_per_step": lf["h2d"],
                                     "d2h_bytes_per_step": lf["d2h"]},
                             "alg_gb_per_s_rank0": lf["loc"]["alg_gb_per_s"], "launches_per_step": lf["loc"]["launches"],
                             "walks_end_at": "samples every %d text positions (device-side dense samples, %d bytes of HBM; the serialized index samples every %d)"
                                             % (ix.locate_sample_rate, ix.dense_sample_bytes(), ix.sample_rate) if ix.dense_sample_bytes()
                                             else "the index's own samples (every %d text positions)" % ix.sample_rate}
            if lf["loc_own"]:
                lo = lf["loc_own"]
                out["locate"]["own_samples_rank0"] = {
                    "what": "the same call with the dense samples switched off: walks end at the serialized index's own samples (identical positions)",
                    "hits_per_s": lo["hits_per_s"], "ms_per_step": lo["ms_per_step"], "kernel_ms": lo["kernel_ms"], "lf_steps": lo["lf_steps"],
                    "roofline_frac": lo["kernel_alg_bytes"] / (lo["kernel_ms"] / 1e3) / 1e9 / peak}
            out["extract_until_boundary"] = {"metric": "records/sec (extractUntilBoundary('\\n'), dst %d chars, of located hits)" % args.dst_len,
                                             "value": all_records / (eub_ms / 1e3), "unit": "records/s", "chars_per_s": all_chars / (eub_ms / 1e3),
                                             "ms_per_step": eub_ms, "records_per_step": all_records, "launches_per_step": lf["eub"]["launches"]}
        if lf:
            out["locate_records"] = {
                "metric": "fused locate -> extractUntilBoundary: hits/s with every distinct record read once (fmgpu_extract_records_batch_device, rank 0)",
                "same_hits_as_extract_until_boundary": lf["rec"], "first_hits_of_the_batch": lf["rec_first"],
                "checked": "statuses and lengths of all hits, contents of 20,000 sampled hits equal the extractUntilBoundary leg"}
        # roofline objects of the two LF kernels: algorithmic bytes (records the lanes need, workloads.py) over the live kernel time
        if lf:
            for key, kname, leg, k_ms in (("locate", "k_locate", lf["loc"], loc_k_ms), ("extract_until_boundary", "k_extract<EUB>", lf["eub"], eub_k_ms)):
                ach = leg["kernel_alg_bytes"] / (k_ms / 1e3) / 1e9
                out[key]["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                        "kernel": kname, "kernel_ms": k_ms, "peak_source": peak_src,
                                        "alg_bytes_per_launch_rank0": leg["kernel_alg_bytes"],
                                        "records_rank0": leg.get("records_read") if "records_read" in leg else leg.get("records")}
        # DRAM traffic per launch of each kernel: from the ncu --set full capture of THIS round's kernels (tools/ncu_traffic.py
        # rewrites profiles/kernel_traffic.json from the capture; the file names the commit and the capture's own kernel times)
        traffic_file = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as fh:
                    tr = json.load(fh)
                for key, kname in (("roofline", "k_count"), ("locate", "k_locate"), ("extract_until_boundary", "k_extract")):
                    tgt = out["roofline"] if key == "roofline" else (out.get(key) or {}).get("roofline")
                    if tgt is not None and kname in tr.get("kernels", {}):
                        tgt["traffic"] = tr["kernels"][kname].get("dram_bytes_per_launch")
                        tgt["traffic_source"] = {k: tr.get(k) for k in ("captured_at_commit", "workload", "note")}
                        tgt["dram_over_requested"] = tr["kernels"][kname].get("dram_over_requested")
            except Exception:
                pass
        # the same fraction counting only the records actually loaded (a start / end pair that shares a record loads it once)
        out["roofline"]["dedup"] = {"records_loaded_bytes": 32.0 * stats["search_records_loaded"],
                                    "achieved": 32.0 * stats["search_records_loaded"] / (kern_ms / 1e3) / 1e9,
                                    "frac": 32.0 * stats["search_records_loaded"] / (kern_ms / 1e3) / 1e9 / peak}
        # north-star "fraction of the gather roofline": what the part sustains for dependent, uniformly random 32-byte sector
        # reads (tools/gather_peak.cu -> profiles/r01_gather_peak.jsonl: 38-70 G sectors/s depending on the footprint) next to
        # what k_count moves: record loads it issues, and DRAM sectors (ncu traffic of the committed capture / live kernel time)
        t_s = kern_ms / 1e3
        out["roofline"]["gather"] = {
            "record_loads_per_s": stats["search_records_loaded"] / t_s,
            "dram_sectors_per_s": (out["roofline"]["traffic"] / 32.0 / t_s) if out["roofline"]["traffic"] else None,
            "uniform_random_sector_peak_per_s": [38e9, 70e9],
            "note": "k_count runs above the uniform-random rate because cells hit L2 and neighbouring positions share DRAM rows",
        }
        if sweep:
            out["locate_sample_rate_sweep"] = {
                "workload": "locate, max %d hits per pattern, the first %d patterns of every rank's batch, FmIndex(sampleRate 16 / 32 / 64) over the same text; %d GPU(s), query-sharded"
                            % (args.max_hits, min(args.sweep_patterns, n_pat), world),
                "by_sample_rate": {str(sr): {"hits_per_s": sweep[sr]["hits_per_s"], "ms_per_step": sweep[sr]["ms_per_step_max_over_ranks"],
                                             "hits_per_step": sweep[sr]["hits_all_ranks"], "index_hbm_bytes": sweep[sr]["hbm_bytes"],
                                             "serialized_bytes": sweep[sr]["serialized_bytes"],
                                             "with_dense_samples": ({"hits_per_s": sweep[sr]["dense_hits_per_s"], "rate": sweep[sr].get("dense_rate"),
                                                                     "extra_hbm_bytes": sweep[sr].get("dense_bytes")}
                                                                    if sweep[sr]["dense_hits_per_s"] else None)} for sr in sorted(sweep)},
                "note": "hits_per_s / index_hbm_bytes: walks end at the serialized index's own samples (dense samples off); with_dense_samples: the library's default"}
        if strong:
            out["strong"] = {
                "workload": "the ONE batch of %d patterns sharded over %d GPUs (BASELINE.json configs[1] as written)" % (n_pat, world),
                "device_resident": {"value": n_pat * args.steps / (strong_dev_ms / 1e3), "unit": UNIT, "ms_per_step": strong_dev_ms / args.steps,
                                    "how": "one process per GPU, slice r of the batch resident on GPU r, CUDA events, max over ranks"},
                "e2e_one_process": {"value": n_pat * args.steps / (strong["one_ms"] / 1e3), "unit": UNIT, "ms_per_step": strong["one_ms"] / args.steps,
                                    "h2d_bytes_per_step": strong.get("h2d"), "d2h_bytes_per_step": 8 * n_pat,
                                    "how": "rank 0 alone: one fmgpu_index replicated on all %d GPUs by the library, one fmgpu_count_batch call per step with pinned host buffers" % world},
                "e2e_one_process_utf8": ({"value": n_pat * args.steps / (strong["one_utf8_ms"] / 1e3), "unit": UNIT,
                                          "ms_per_step": strong["one_utf8_ms"] / args.steps} if strong["one_utf8_ms"] else None),
            }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            oracle_ix, build = load_oracle(blob)
            v, dt, cpu_counts = cpu_count_throughput(oracle_ix, chars, off, args.cpu_sample, threads, args.cpu_repeats)
            assert np.array_equal(cpu_counts, h_counts[: cpu_counts.size]), "GPU counts differ from the CPU oracle"
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "build": build,
                                   "patterns_per_s_by_threads": cpu_thread_scaling(oracle_ix, chars, off, 100_000, threads),
                                   "sample": "first %d of the %d patterns, best of %d passes, %.1fs per pass (C++ restatement of the reference's Java loops; no JVM in this image)"
                                             % (cpu_counts.size, n_pat, args.cpu_repeats, dt)}
            if lf:  # CPU arms of the LF legs on bounded samples; their outputs double as the parity check of the GPU legs
                frm = lf["d_from"][: args.cpu_eub_sample].cpu().numpy().astype(np.int32)
                arms = cpu_lf_arms(oracle_ix, chars, off, args, threads, frm)
                k, w_n, w_pos, m, w_arena, w_ln, w_st = arms.pop("_check")
                ho = lf["d_hit_off"][: k + 1].cpu().numpy().astype(np.int64)
                pos = lf["d_pos"][: int(ho[-1])].cpu().numpy()
                assert np.array_equal(np.diff(ho), w_n), "locate hit counts differ from the CPU oracle"
                flat = np.concatenate([w_pos[i, : w_n[i]] for i in range(k)]) if k else np.zeros(0, np.int32)
                assert np.array_equal(pos, flat), "located positions differ from the CPU oracle"
                arena = lf["d_arena"][:m].cpu().numpy().view(np.uint16)
                assert np.array_equal(lf["d_st"][:m].cpu().numpy(), w_st) and np.array_equal(lf["d_len"][:m].cpu().numpy()[w_st == 0], w_ln[w_st == 0])
                cols = np.arange(arena.shape[1])[None, :]
                valid = (cols < w_ln[:, None]) & (w_st == 0)[:, None]
                assert np.array_equal(arena[valid], w_arena[valid]), "extracted records differ from the CPU oracle"
                out["locate"]["oracle_checked_patterns"] = k
                out["locate"]["cpu_baseline"] = arms["locate"]
                out["extract_until_boundary"]["oracle_checked_records"] = m
                out["extract_until_boundary"]["cpu_baseline"] = arms["extract_until_boundary"]
        emit(out)
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
