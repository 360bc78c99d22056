"""Suffix array by prefix doubling on the GPU (torch sort / cumsum / scatter) — index PRODUCTION only.

Not part of the query path and not a port of anything in the reference: the reference builds its
suffix array with jsuffixarrays' DivSufSort on the JVM (indices/src/main/java/com/dynatrace/fm/FmIndex.java:332-341).
The suffix array of a text ending in a unique smallest sentinel is unique, so this produces the
identical array; it exists so that the benchmark can build the 1 GB-text index on the GPU box in
seconds instead of the minutes a sequential host suffix sorter needs.
"""
from __future__ import annotations

import sys

import numpy as np
import torch


def suffix_array(codes: np.ndarray, sigma: int, device=None, verbose: bool = False, keep_on_device: bool = False):
    """``codes``: uint16 alphabet codes of the text with the sentinel (code 0, unique, smallest) as last
    element.  Returns the suffix array as int32 (a numpy array, or the device tensor with ``keep_on_device``)."""
    n = int(codes.size)
    assert n < 2**31
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    bits = max(1, int(sigma - 1).bit_length())
    k0 = max(1, 62 // bits)
    t = torch.from_numpy(codes.astype(np.int16)).to(dev).to(torch.int64)
    # initial key: k0 symbols packed most-significant first (positions past the end count as 0)
    key = torch.zeros(n, dtype=torch.int64, device=dev)
    for j in range(k0):
        key <<= bits
        if j < n:
            key[: n - j] |= t[j:]
    del t
    h = k0
    rank = None
    while True:
        skey, idx = torch.sort(key)
        del key
        flag = torch.ones(n, dtype=torch.int64, device=dev)
        flag[1:] = (skey[1:] != skey[:-1]).to(torch.int64)
        del skey
        r_sorted = torch.cumsum(flag, 0)  # dense ranks 1..m in sorted order
        del flag
        m = int(r_sorted[-1].item())
        rank = torch.empty(n, dtype=torch.int64, device=dev)
        rank[idx] = r_sorted
        del r_sorted
        if verbose:
            print("[gpu_sa] h=%d distinct=%d/%d" % (h, m, n), file=sys.stderr, flush=True)
        if m == n:
            sa = idx.to(torch.int32)
            del idx, rank
            if keep_on_device:
                return sa
            out = sa.cpu().numpy()
            del sa
            return out
        del idx
        key = rank << 32
        if h < n:
            key[: n - h] |= rank[h:]
        del rank
        h *= 2


def build_index_gpu(text, sample_rate: int = 32, enable_extraction: bool = True, framed: bool = True, device=None, verbose: bool = False,
                    threads: int = 0) -> bytes:
    """Serialized ``FmIndex`` with the suffix array, the BWT and the sampled structures produced on the GPU: the suffix array
    never leaves the device (``fmgpu_build_bwt_samples_device``, index4j_b200/csrc/kernels_build.cuh); the host encodes the
    wavelet structure / RRR vector and serializes.  Byte-identical with ``builder.build_index`` (tests/test_build_gpu.py)."""
    import ctypes as C
    import time

    from .builder import as_chars, build_index_from_parts, map_text
    from .fm_index import native

    t0 = time.time()
    t = as_chars(text)
    codes, sigma = map_text(t)
    dev = torch.device(device) if device is not None else torch.device("cuda")
    with torch.cuda.device(dev):
        sa = suffix_array(codes, sigma, device=dev, verbose=verbose, keep_on_device=True)
        t1 = time.time()
        length = int(codes.size)
        d_codes = torch.from_numpy(codes.view(np.int16)).to(dev)
        n_words = (length + 31) // 32
        n_samp = (length - 1) // sample_rate + 1
        d_bwt = torch.empty(length, dtype=torch.int16, device=dev)
        d_mask = torch.empty(n_words, dtype=torch.int32, device=dev)
        d_suf = torch.empty(n_samp, dtype=torch.int32, device=dev)
        d_pos = torch.zeros(length // sample_rate + 2, dtype=torch.int32, device=dev) if enable_extraction else None
        got = C.c_int64(0)
        lib = native()
        lib.fmgpu_build_bwt_samples_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                       C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
        rc = lib.fmgpu_build_bwt_samples_device(d_codes.data_ptr(), sa.data_ptr(), length, sample_rate, d_bwt.data_ptr(), d_mask.data_ptr(),
                                                d_suf.data_ptr(), n_samp, d_pos.data_ptr() if d_pos is not None else None, C.byref(got),
                                                torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(lib.fmgpu_last_error().decode())
        assert got.value == n_samp, (got.value, n_samp)
        del sa, d_codes
        bwt = d_bwt.cpu().numpy().view(np.uint16)
        mask = d_mask.cpu().numpy().view(np.uint32)
        suf = d_suf.cpu().numpy()
        pos = d_pos.cpu().numpy() if d_pos is not None else None
        del d_bwt, d_mask, d_suf, d_pos
        torch.cuda.empty_cache()
    t2 = time.time()
    blob = build_index_from_parts(t, bwt, mask, suf, pos, sample_rate, enable_extraction, framed, threads, verbose)
    if verbose:
        print("[gpu_build] n=%d: suffix array %.1fs, bwt + samples on device (+ D2H) %.1fs, host encode + serialize %.1fs"
              % (t.size, t1 - t0, t2 - t1, time.time() - t2), file=sys.stderr, flush=True)
    return blob
