"""Host-side index producer (ctypes over ``libfmhost.so``).

The reference builds an ``FmIndex`` on the JVM (``new FmIndexBuilder().setSampleRate(..)
.setEnableExtraction(..).build(char[])``, indices/src/main/java/com/dynatrace/fm/FmIndexBuilder.java:21-62)
and ships it as ``Serialization.writeToByteArray(FmIndex::write, index)``.  There is no JVM in this
image, so ``libfmhost`` produces the same serialized layout natively; the GPU engine only ever
sees those bytes.  Nothing here is on the query path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

_lib = None


def _host():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.HOST_LIB):
            _build.build_host()
        lib = C.CDLL(_build.HOST_LIB)
        lib.fmhost_last_error.restype = C.c_char_p
        lib.fmhost_build.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        lib.fmhost_build_with_sa.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        lib.fmhost_build_with_parts.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                                C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        lib.fmhost_free.argtypes = [C.c_void_p]
        lib.fmhost_map_text.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        lib.fmhost_map_text.restype = C.c_int32
        lib.fmhost_suffix_array.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        lib.fmhost_gen_log_text.argtypes = [C.c_void_p, C.c_int64, C.c_uint64]
        lib.fmhost_gen_patterns.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_int32, C.c_int32, C.c_uint64,
                                            C.c_void_p, C.c_void_p]
        lib.fmhost_build_wfbb.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        lib.fmhost_build_rrr.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        _lib = lib
    return _lib


def _take(lib, rc, out, n) -> bytes:
    if rc != 0:
        raise ValueError(lib.fmhost_last_error().decode())
    try:
        return C.string_at(out, n.value)
    finally:
        lib.fmhost_free(out)


def build_wfbb(symbols, sampling_rate: int = 64, framed: bool = False) -> bytes:
    """Serialized stand-alone ``WaveletFixedBlockBoosting`` (``new WaveletFixedBlockBoosting(short[]/char[] text, samplingRate)``)."""
    lib = _host()
    s = as_chars(symbols)
    out, n = C.c_void_p(), C.c_uint64()
    rc = lib.fmhost_build_wfbb(s.ctypes.data, s.size, sampling_rate, int(framed), 2, C.byref(out), C.byref(n))
    return _take(lib, rc, out, n)


def build_rrr(bits, sample_size: int = 32, framed: bool = False) -> bytes:
    """Serialized stand-alone ``RrrVector`` over a 0/1 array (``new RrrVector(BitVector, sampleSize)``)."""
    lib = _host()
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    words = np.zeros((b.size + 63) // 64 + 1, dtype=np.uint64)
    packed = np.packbits(b, bitorder="little")
    words.view(np.uint8)[: packed.size] = packed
    out, n = C.c_void_p(), C.c_uint64()
    rc = lib.fmhost_build_rrr(words.ctypes.data, b.size, sample_size, int(framed), C.byref(out), C.byref(n))
    return _take(lib, rc, out, n)


def as_chars(text) -> np.ndarray:
    """``str`` / ``bytes``-like / array -> contiguous uint16 array of UTF-16 code units (Java ``char[]``)."""
    if isinstance(text, str):
        return np.frombuffer(text.encode("utf-16-le", "surrogatepass"), dtype=np.uint16).copy()
    a = np.ascontiguousarray(text)
    if a.dtype != np.uint16:
        a = a.astype(np.uint16)
    return a


def build_index(text, sample_rate: int = 32, enable_extraction: bool = True, framed: bool = True,
                threads: int = 0, verbose: bool = False, suffix_array: np.ndarray | None = None) -> bytes:
    """Serialized ``FmIndex`` (reference layout) over ``text``.

    ``framed=True`` wraps the primitives in ``ObjectOutputStream`` block-data records like
    ``Serialization.writeToByteArray`` (serialization/Serialization.java:67-78).
    """
    lib = _host()
    t = as_chars(text)
    out = C.c_void_p()
    n = C.c_uint64()
    if suffix_array is None:
        rc = lib.fmhost_build(t.ctypes.data, t.size, sample_rate, int(enable_extraction), int(framed), threads,
                              int(verbose), C.byref(out), C.byref(n))
    else:
        sa = np.ascontiguousarray(suffix_array, dtype=np.int32)
        assert sa.size == t.size + 1
        rc = lib.fmhost_build_with_sa(t.ctypes.data, t.size, sa.ctypes.data, sample_rate, int(enable_extraction),
                                      int(framed), threads, int(verbose), C.byref(out), C.byref(n))
    if rc != 0:
        msg = lib.fmhost_last_error().decode()
        if "more than" in msg:
            raise ValueError(msg)  # IllegalArgumentException, FmIndex.java:423-426
        raise RuntimeError(msg)
    try:
        return C.string_at(out, n.value)
    finally:
        lib.fmhost_free(out)


def build_index_from_parts(text, bwt, mask_words, suffixes, positions, sample_rate: int = 32, enable_extraction: bool = True,
                           framed: bool = True, threads: int = 0, verbose: bool = False) -> bytes:
    """Serialized ``FmIndex`` from the pieces the device stage produced (``fmgpu_build_bwt_samples_device``): the BWT over
    alphabet codes, the sampled-row marks (uint32 words, LSB first), the SA samples in row order and the inverse-SA samples."""
    lib = _host()
    t = as_chars(text)
    bwt = np.ascontiguousarray(bwt, dtype=np.uint16)
    mask_words = np.ascontiguousarray(mask_words, dtype=np.uint32)
    suffixes = np.ascontiguousarray(suffixes, dtype=np.int32)
    assert bwt.size == t.size + 1 and mask_words.size >= (t.size + 1 + 31) // 32
    pos_ptr = None
    if enable_extraction:
        positions = np.ascontiguousarray(positions, dtype=np.int32)
        assert positions.size == (t.size + 1) // sample_rate + 2
        pos_ptr = positions.ctypes.data
    out = C.c_void_p()
    n = C.c_uint64()
    rc = lib.fmhost_build_with_parts(t.ctypes.data, t.size, bwt.ctypes.data, mask_words.ctypes.data, suffixes.ctypes.data, suffixes.size,
                                     pos_ptr, sample_rate, int(enable_extraction), int(framed), threads, int(verbose), C.byref(out), C.byref(n))
    if rc != 0:
        msg = lib.fmhost_last_error().decode()
        if "more than" in msg:
            raise ValueError(msg)
        raise RuntimeError(msg)
    try:
        return C.string_at(out, n.value)
    finally:
        lib.fmhost_free(out)


def map_text(text) -> tuple[np.ndarray, int]:
    """First-appearance code mapping with the sentinel appended (what the suffix sorter sorts)."""
    t = as_chars(text)
    codes = np.empty(t.size + 1, dtype=np.uint16)
    sigma = _host().fmhost_map_text(t.ctypes.data, t.size, codes.ctypes.data)
    return codes, int(sigma)


def suffix_array_host(codes: np.ndarray, sigma: int) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint16)
    sa = np.empty(codes.size, dtype=np.int32)
    _host().fmhost_suffix_array(codes.ctypes.data, codes.size, sigma, sa.ctypes.data)
    return sa


def gen_log_text(n: int, seed: int = 0x1DE40001) -> np.ndarray:
    """``n`` UTF-16 units of deterministic synthetic log-like ASCII text."""
    out = np.empty(n, dtype=np.uint16)
    _host().fmhost_gen_log_text(out.ctypes.data, n, seed)
    return out


def gen_patterns(text: np.ndarray, n_pat: int, min_len: int, max_len: int, seed: int = 42):
    """Substring workload of the reference's JMH state (jmh/.../fm/FmIndexThroughputState.java:76-83).

    Returns ``(chars uint16[total], offsets uint64[n_pat+1])``.
    """
    t = as_chars(text)
    chars = np.empty(n_pat * max_len, dtype=np.uint16)
    off = np.empty(n_pat + 1, dtype=np.uint64)
    _host().fmhost_gen_patterns(t.ctypes.data, t.size, n_pat, min_len, max_len, seed, chars.ctypes.data, off.ctypes.data)
    return chars[: int(off[-1])].copy(), off


class FmIndexBuilder:
    """Mirror of ``com.dynatrace.fm.FmIndexBuilder`` (defaults sampleRate=32, extraction on)."""

    def __init__(self):
        self._sample_rate = 32
        self._enable_extraction = True

    def setSampleRate(self, sample_rate: int) -> "FmIndexBuilder":
        self._sample_rate = int(sample_rate)
        return self

    def setEnableExtraction(self, enable: bool) -> "FmIndexBuilder":
        self._enable_extraction = bool(enable)
        return self

    def build_serialized(self, text) -> bytes:
        return build_index(text, self._sample_rate, self._enable_extraction)

    def build(self, text, device: int | None = None):
        from .fm_index import FmIndex
        return FmIndex.read(self.build_serialized(text), device=device)
