"""ctypes binding of the CPU oracle + naive text-scan checkers.  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs — never by the
product package.  The naive scanners restate the reference's own test oracle
(indices/src/test/java/com/dynatrace/util/Util.java:108-279).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

STATUS_MESSAGES = {
    1: "Text recovery not enabled at build time",
    2: "Requested position less than 0",
    3: "Stop position longer than index string",
    4: "Requested position longer than index string",
    5: "Supplied destination is not large enough",
    6: "Supplied destination for extraction has size zero",
    7: "Boundary does not exist",
    8: "Extraction does not fit in the supplied destination. Currently extracted: {n}",
    9: "ArrayIndexOutOfBoundsException",
    10: "Found a character that exceeds (32767): it was {n}",
    12: "locate does not terminate in the reference (LF walk in a cycle behind a truncated run-block symbol)",
}

NATIVE_LIB = os.path.join(HERE, "liboracle_native.so")

_lib = None
_lib_path = LIB


def use_native() -> str:
    """bench.py's CPU arm: build the oracle ON this box with -march=native and without the work counters (`make native`) and
    make lib() load that build.  Must be called before the first use of the oracle in the process; falls back to the portable
    build if the compile fails.  Returns a description of the build that will be timed."""
    global _lib_path
    if _lib is not None:
        return "portable build (-march=x86-64-v3, thread-local work counters)" if _lib_path == LIB else "native"
    try:
        subprocess.check_call(["make", "-s", "-B", "-C", HERE, "native"])
        _lib_path = NATIVE_LIB
        return "g++ -O3 -march=native, work counters compiled out"
    except Exception:  # noqa: BLE001
        return "portable build (-march=x86-64-v3, thread-local work counters)"


def lib():
    global _lib
    if _lib is None:
        if _lib_path == LIB and not os.path.exists(LIB):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = C.CDLL(_lib_path)
        L.orc_last_error.restype = C.c_char_p
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        L.orc_fm_load.argtypes = [vp, u64, C.POINTER(vp)]
        L.orc_fm_free.argtypes = [vp]
        for f in ("orc_fm_input_length", "orc_fm_alphabet_length", "orc_fm_sample_rate", "orc_fm_extract_enabled"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = i32
        L.orc_fm_block_size_log.argtypes = [vp, i64]
        L.orc_fm_num_superblocks.argtypes = [vp]
        L.orc_fm_num_superblocks.restype = i64
        L.orc_fm_count.argtypes = [vp, vp, i32, i32, C.POINTER(i32)]
        L.orc_fm_locate.argtypes = [vp, vp, i32, i32, vp, i64, i32, C.POINTER(i32)]
        L.orc_fm_extract.argtypes = [vp, i32, i32, vp, i64, i32, C.POINTER(i32)]
        L.orc_fm_extract_until_boundary.argtypes = [vp, i32, vp, i64, i32, C.c_uint16, i32, C.POINTER(i32)]
        L.orc_fm_count_batch.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, i32]
        L.orc_fm_locate_batch.argtypes = [vp, vp, vp, C.c_uint32, i32, vp, vp, i64, vp, i32]
        L.orc_fm_count_batch_utf8.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, i32]
        L.orc_fm_locate_batch_utf8.argtypes = [vp, vp, vp, C.c_uint32, i32, vp, vp, i64, vp, i32]
        L.orc_convert_utf8.argtypes = [vp, i64, i64, i64, vp, C.POINTER(i32)]
        L.orc_convert_utf8.restype = i64
        L.orc_fm_extract_batch.argtypes = [vp, vp, vp, C.c_uint32, vp, i64, vp, vp, i32, i32]
        L.orc_fm_extract_until_boundary_batch.argtypes = [vp, vp, C.c_uint32, C.c_uint16, i32, i32, vp, vp, vp, i32, i32]
        L.orc_fm_stats.argtypes = [vp, vp, i32]
        L.orc_fm_wfbb_rank.argtypes = [vp, i64, i32, C.POINTER(i64)]
        L.orc_fm_wfbb_inverse_select.argtypes = [vp, i64, C.POINTER(i64)]
        L.orc_fm_sampled_rank.argtypes = [vp, i32]
        L.orc_fm_sampled_access.argtypes = [vp, i32]
        L.orc_wfbb_load.argtypes = [vp, u64, C.POINTER(vp)]
        L.orc_wfbb_free.argtypes = [vp]
        L.orc_wfbb_rank.argtypes = [vp, i64, i32, C.POINTER(i64)]
        L.orc_wfbb_inverse_select.argtypes = [vp, i64, C.POINTER(i64)]
        L.orc_rrr_load.argtypes = [vp, u64, C.POINTER(vp)]
        L.orc_rrr_free.argtypes = [vp]
        L.orc_rrr_rank_ones.argtypes = [vp, i32]
        L.orc_rrr_access.argtypes = [vp, i32]
        L.orc_rrr_inverse_table.argtypes = [vp]
        _lib = L
    return _lib


class JavaException(Exception):
    def __init__(self, status: int, n: int = 0):
        self.status = status
        self.n = n
        super().__init__(STATUS_MESSAGES.get(status, "status %d" % status).format(n=n))


def _u16(a) -> np.ndarray:
    if isinstance(a, str):
        return np.frombuffer(a.encode("utf-16-le", "surrogatepass"), dtype=np.uint16).copy()
    return np.ascontiguousarray(a, dtype=np.uint16)


class OracleFmIndex:
    """The Java ``FmIndex`` API, restated on the CPU (single calls raise JavaException)."""

    def __init__(self, serialized: bytes):
        self._h = C.c_void_p()
        buf = np.frombuffer(serialized, dtype=np.uint8)
        rc = lib().orc_fm_load(buf.ctypes.data, buf.size, C.byref(self._h))
        if rc != 0:
            raise IOError(lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_fm_free(self._h)
            self._h = None

    def getInputLength(self):
        return lib().orc_fm_input_length(self._h)

    def getAlphabetLength(self):
        return lib().orc_fm_alphabet_length(self._h)

    @property
    def sample_rate(self):
        return lib().orc_fm_sample_rate(self._h)

    def block_size_logs(self):
        return [lib().orc_fm_block_size_log(self._h, i) for i in range(lib().orc_fm_num_superblocks(self._h))]

    def count(self, pattern, offset=0, length=None):
        p = _u16(pattern)
        if length is None:
            length = p.size
        out = C.c_int32()
        st = lib().orc_fm_count(self._h, p.ctypes.data, offset, length, C.byref(out))
        if st:
            raise JavaException(st)
        return out.value

    def locate(self, pattern, offset=0, length=None, cap=None, max_matches=-1):
        p = _u16(pattern)
        if length is None:
            length = p.size
        if cap is None:
            cap = max(1, self.count(pattern, offset, length)) if max_matches <= 0 else max_matches
        loc = np.zeros(cap, dtype=np.int32)
        out = C.c_int32()
        st = lib().orc_fm_locate(self._h, p.ctypes.data, offset, length, loc.ctypes.data, cap, max_matches, C.byref(out))
        if st:
            raise JavaException(st)
        return loc[: out.value].copy()

    def extract(self, start, stop, dst_len=None, offset=0):
        if dst_len is None:
            dst_len = max(0, stop - start) + offset
        dst = np.zeros(max(dst_len, 1), dtype=np.uint16)
        out = C.c_int32()
        st = lib().orc_fm_extract(self._h, start, stop, dst.ctypes.data, dst_len, offset, C.byref(out))
        if st:
            raise JavaException(st, out.value)
        return dst[offset: offset + out.value].copy()

    def extract_until_boundary(self, frm, dst_len, boundary, mode=0, offset=0):
        dst = np.zeros(max(dst_len, 1), dtype=np.uint16)
        out = C.c_int32()
        st = lib().orc_fm_extract_until_boundary(self._h, frm, dst.ctypes.data, dst_len, offset, boundary, mode, C.byref(out))
        if st:
            raise JavaException(st, out.value)
        return dst[offset: offset + out.value].copy()

    # --- batch forms ----------------------------------------------------------------------
    def count_batch(self, chars, pat_off, threads=1):
        chars = _u16(chars)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        lib().orc_fm_count_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, n, counts.ctypes.data, status.ctypes.data, threads)
        return counts, status

    def count_batch_utf8(self, data, pat_off, threads=1):
        """convertBytePatternToCharPattern + count per pattern (every pattern its own byte[])."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        lib().orc_fm_count_batch_utf8(self._h, data.ctypes.data, pat_off.ctypes.data, n, counts.ctypes.data, status.ctypes.data, threads)
        return counts, status

    def locate_batch_utf8(self, data, pat_off, max_matches, stride, threads=1):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        n_hits = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        pos = np.zeros((n, stride), dtype=np.int32)
        lib().orc_fm_locate_batch_utf8(self._h, data.ctypes.data, pat_off.ctypes.data, n, max_matches, n_hits.ctypes.data,
                                       pos.ctypes.data, stride, status.ctypes.data, threads)
        return n_hits, pos, status

    def locate_batch(self, chars, pat_off, max_matches, stride, threads=1):
        chars = _u16(chars)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        n_hits = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        pos = np.zeros((n, stride), dtype=np.int32)
        lib().orc_fm_locate_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, n, max_matches, n_hits.ctypes.data,
                                  pos.ctypes.data, stride, status.ctypes.data, threads)
        return n_hits, pos, status

    def extract_batch(self, start, stop, stride, threads=1, offset=0):
        start = np.ascontiguousarray(start, dtype=np.int32)
        stop = np.ascontiguousarray(stop, dtype=np.int32)
        n = start.size
        arena = np.zeros((n, stride), dtype=np.uint16)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        lib().orc_fm_extract_batch(self._h, start.ctypes.data, stop.ctypes.data, n, arena.ctypes.data, stride, ln.ctypes.data,
                                   st.ctypes.data, threads, offset)
        return arena, ln, st

    def extract_until_boundary_batch(self, frm, boundary, dst_len, mode=0, threads=1, offset=0):
        frm = np.ascontiguousarray(frm, dtype=np.int32)
        n = frm.size
        arena = np.zeros((n, max(dst_len, 1)), dtype=np.uint16)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        lib().orc_fm_extract_until_boundary_batch(self._h, frm.ctypes.data, n, boundary, dst_len, mode, arena.ctypes.data,
                                                  ln.ctypes.data, st.ctypes.data, threads, offset)
        return arena, ln, st

    def stats(self, reset=True):
        out = np.zeros(4, dtype=np.uint64)
        lib().orc_fm_stats(self._h, out.ctypes.data, int(reset))
        return dict(ranks=int(out[0]), rank_levels=int(out[1]), lf_steps=int(out[2]), lf_levels=int(out[3]))

    def wfbb_rank(self, pos, sym):
        out = C.c_int64()
        st = lib().orc_fm_wfbb_rank(self._h, pos, sym, C.byref(out))
        if st:
            raise JavaException(st)
        return out.value

    def wfbb_inverse_select(self, pos):
        out = C.c_int64()
        st = lib().orc_fm_wfbb_inverse_select(self._h, pos, C.byref(out))
        if st:
            raise JavaException(st)
        return out.value

    def sampled_rank(self, pos):
        return lib().orc_fm_sampled_rank(self._h, pos)

    def sampled_access(self, pos):
        return lib().orc_fm_sampled_access(self._h, pos)


def convert_byte_pattern_to_char_pattern(pattern, offset=0, length=None) -> np.ndarray:
    """FmIndex.convertBytePatternToCharPattern (fm/FmIndex.java:239-298); raises JavaException like the reference."""
    data = np.ascontiguousarray(np.frombuffer(bytes(pattern), dtype=np.uint8))
    length = data.size - offset if length is None else length
    dst = np.zeros(max(1, length), dtype=np.uint16)
    val = C.c_int32(0)
    n = lib().orc_convert_utf8(data.ctypes.data, data.size, offset, length, dst.ctypes.data, C.byref(val))
    if n < 0:
        raise JavaException(int(-100 - n), val.value)
    return dst[:n].copy()


def rrr_inverse_table() -> np.ndarray:
    t = np.zeros(32768, dtype=np.uint16)
    lib().orc_rrr_inverse_table(t.ctypes.data)
    return t


# ---------------------------------------------------------------------------------------------
# Naive text scanners (Util.java:108-279) — independent of every index structure.
# ---------------------------------------------------------------------------------------------
def naive_locations(text: np.ndarray, pattern: np.ndarray) -> np.ndarray:
    """Sorted start positions of all (overlapping) occurrences (Util.java:124-138)."""
    text = _u16(text)
    pattern = _u16(pattern)
    m = pattern.size
    if m == 0 or m > text.size:
        return np.zeros(0, dtype=np.int64)
    cand = np.flatnonzero(text[: text.size - m + 1] == pattern[0])
    for k in range(1, m):
        if cand.size == 0:
            break
        cand = cand[text[cand + k] == pattern[k]]
    return cand.astype(np.int64)


def naive_count(text, pattern) -> int:
    return int(naive_locations(text, pattern).size)


def naive_extract_until_boundary(text: np.ndarray, seed: int, boundary: int, mode: int = 0) -> np.ndarray:
    """Util.java:168-257 (mode 0 both sides, 1 left, 2 right)."""
    text = _u16(text)
    if text[seed] == boundary:
        return np.zeros(0, dtype=np.uint16)
    lo = seed
    while lo - 1 >= 0 and text[lo - 1] != boundary:
        lo -= 1
    hi = seed + 1
    while hi < text.size and text[hi] != boundary:
        hi += 1
    if mode == 0:
        return text[lo:hi].copy()
    if mode == 1:
        return text[lo: seed + 1].copy()
    return text[seed + 1: hi].copy()
