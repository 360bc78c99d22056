"""ctypes binding of tests/support/libflatcheck.so (host replay of the device layout + lane logic)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libflatcheck.so")
ROOT = os.path.dirname(os.path.dirname(HERE))
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, "flatcheck.cpp")] + [
        os.path.join(ROOT, "index4j_b200", "csrc", f) for f in ("flatten.hpp", "jstream.hpp", "walk_lane.h", "lane_logic.h", "lf_lane.h", "ldrec.h", "layout.h", "count_lane.h", "utf8_lane.h", "host_pack.hpp")
    ]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-unknown-pragmas",
                               "-o", LIB, srcs[0]])
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i32, u32, u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
        L.fc_last_error.restype = C.c_char_p
        L.fc_load.argtypes = [vp, u64, i32, C.POINTER(vp)]
        L.fc_free.argtypes = [vp]
        L.fc_alphabet_length.argtypes = [vp]
        L.fc_sizes.argtypes = [vp, vp]
        L.fc_rank.argtypes = [vp, u32, u32, C.POINTER(C.c_int64)]
        L.fc_count_batch.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp]
        L.fc_inverse_select.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64)]
        L.fc_count_batch_table.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp]
        L.fc_build_start_table.argtypes = [vp, u32]
        L.fc_build_start_table.restype = C.c_uint64
        L.fc_locate_rows.argtypes = [vp, vp, u32, vp]
        L.fc_cell_kinds8.argtypes = [vp, vp]
        L.fc_pack_threads.restype = i32
        L.fc_pack_narrow.argtypes = [vp, u64, u32, u32, vp, vp]
        L.fc_find_cells.argtypes = [vp, u32, u32, vp]
        L.fc_find_cells.restype = C.c_uint32
        L.fc_dense_build.argtypes = [vp, u32, vp, vp]
        L.fc_dense_build.restype = C.c_int64
        L.fc_locate_rows_dense.argtypes = [vp, vp, vp, vp, u32, vp]
        L.fc_extract.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, vp, i32]
        L.fc_eub.argtypes = [vp, vp, u32, C.c_uint16, i32, i32, vp, vp, vp, vp, i32]
        L.fc_records.argtypes = [vp, vp, u32, C.c_uint16, i32, vp, vp, vp, vp, vp]
        L.fc_records.restype = C.c_uint64
        L.fc_sampled.argtypes = [vp, u32, C.POINTER(i32), C.POINTER(i32)]
        L.fc_unrank_table.argtypes = [vp]
        L.fc_product_rrr_table.argtypes = [vp, vp, vp]
        L.fc_utf8_convert.argtypes = [vp, C.c_uint64, vp, C.POINTER(i32)]
        L.fc_utf8_convert.restype = C.c_int64
        _lib = L
    return _lib


class FlatIndexHost:
    def __init__(self, serialized: bytes, threads: int = 4):
        self._h = C.c_void_p()
        buf = np.frombuffer(serialized, dtype=np.uint8)
        rc = lib().fc_load(buf.ctypes.data, buf.size, threads, C.byref(self._h))
        if rc:
            raise IOError(lib().fc_last_error().decode())
        self.counters = np.zeros(16, dtype=np.uint64)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fc_free(self._h)
            self._h = None

    def sizes(self):
        out = np.zeros(8, dtype=np.uint64)
        lib().fc_sizes(self._h, out.ctypes.data)
        return out

    def cell_kinds(self):
        """(block, symbol) cells by kind: [normal, const, run, throw, range-1K, list, bits, range-4K]"""
        out = np.zeros(8, dtype=np.uint64)
        lib().fc_cell_kinds8(self._h, out.ctypes.data)
        return [int(x) for x in out]

    def find_cells(self, kind: int, max_cells: int = 8):
        """-> [(symbol, first row of the block, block size)] of up to max_cells cells of one CellKind"""
        out = np.zeros(4 * max_cells, dtype=np.uint32)
        n = lib().fc_find_cells(self._h, kind, max_cells, out.ctypes.data)
        return [(int(out[4 * i]), int(out[4 * i + 1]), int(out[4 * i + 2])) for i in range(n)]

    def rank(self, pos, sym):
        out = C.c_int64()
        st = lib().fc_rank(self._h, pos, sym, C.byref(out))
        return st, out.value

    def count_batch(self, chars, pat_off):
        chars = np.ascontiguousarray(chars, dtype=np.uint16)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        ranges = np.zeros(2 * n, dtype=np.uint32)
        lib().fc_count_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, n, counts.ctypes.data, status.ctypes.data,
                             ranges.ctypes.data, self.counters.ctypes.data)
        return counts, status, ranges.reshape(n, 2)

    def inverse_select(self, pos: int):
        out = C.c_int64(0)
        st = lib().fc_inverse_select(self._h, int(pos), C.byref(out))
        return st, int(out.value)

    def build_start_table(self, q: int) -> int:
        return int(lib().fc_build_start_table(self._h, q))

    def count_batch_table(self, chars, pat_off):
        """count_batch starting from the q-gram start table (build_start_table first)"""
        chars = np.ascontiguousarray(chars, dtype=np.uint16)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        ranges = np.zeros(2 * n, dtype=np.uint32)
        lib().fc_count_batch_table(self._h, chars.ctypes.data, pat_off.ctypes.data, n, counts.ctypes.data, status.ctypes.data,
                                   ranges.ctypes.data, self.counters.ctypes.data)
        return counts, status, ranges.reshape(n, 2)

    def locate_rows(self, rows):
        rp = np.ascontiguousarray(rows, dtype=np.uint32).copy()
        lib().fc_locate_rows(self._h, rp.ctypes.data, rp.size, self.counters.ctypes.data)
        return rp.astype(np.int64)

    def dense_build(self, rate: int, length: int):
        """Host replay of the device-side dense-sample build: -> (marks uint32[n_rec, 8], dsa uint32[n_dense]) or None."""
        n_rec = (length + 223) // 224
        marks = np.zeros((n_rec, 8), dtype=np.uint32)
        dsa = np.full((length - 1) // rate + 1, 0xFFFFFFFF, dtype=np.uint32)
        n = lib().fc_dense_build(self._h, rate, marks.ctypes.data, dsa.ctypes.data)
        return None if n < 0 else (marks, dsa)

    def locate_rows_dense(self, marks, dsa, rows):
        rp = np.ascontiguousarray(rows, dtype=np.uint32).copy()
        steps = np.zeros(1, dtype=np.uint64)
        lib().fc_locate_rows_dense(self._h, marks.ctypes.data, dsa.ctypes.data, rp.ctypes.data, rp.size, steps.ctypes.data)
        return rp.astype(np.int64), int(steps[0])

    def extract(self, start, stop, arena_off, offset=0):
        start = np.ascontiguousarray(start, dtype=np.int32)
        stop = np.ascontiguousarray(stop, dtype=np.int32)
        arena_off = np.ascontiguousarray(arena_off, dtype=np.uint64)
        n = start.size
        arena = np.zeros(int(arena_off[-1]) + 1, dtype=np.uint16)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        lib().fc_extract(self._h, start.ctypes.data, stop.ctypes.data, n, arena.ctypes.data, arena_off.ctypes.data, ln.ctypes.data,
                         st.ctypes.data, self.counters.ctypes.data, offset)
        return arena, ln, st

    def eub(self, frm, boundary, dst_len, mode, offset=0):
        frm = np.ascontiguousarray(frm, dtype=np.int32)
        n = frm.size
        arena = np.zeros((n, max(dst_len, 1)), dtype=np.uint16)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        lib().fc_eub(self._h, frm.ctypes.data, n, boundary, dst_len, mode, arena.ctypes.data, ln.ctypes.data, st.ctypes.data,
                     self.counters.ctypes.data, offset)
        return arena, ln, st

    def records(self, frm, boundary, dst_len):
        """host replay of the fused locate -> extractUntilBoundary pipeline (kernels_records.cuh)"""
        frm = np.ascontiguousarray(frm, dtype=np.int32)
        n = frm.size
        idx = np.zeros(n, dtype=np.int32)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        arena = np.zeros((max(n, 1), max(dst_len, 1)), dtype=np.uint16)
        n_rec = lib().fc_records(self._h, frm.ctypes.data, n, boundary, dst_len, idx.ctypes.data, ln.ctypes.data, st.ctypes.data, arena.ctypes.data,
                                 self.counters.ctypes.data)
        return idx, ln, st, arena[: int(n_rec)]

    def sampled(self, pos):
        b, r = C.c_int32(), C.c_int32()
        lib().fc_sampled(self._h, pos, C.byref(b), C.byref(r))
        return b.value, r.value


def product_rrr_table():
    """-> (inverse uint16[32768], class_base uint16[16], bits_needed uint8[16]) of the product's loader (flatten.hpp)"""
    t, cb, bits = np.zeros(32768, dtype=np.uint16), np.zeros(16, dtype=np.uint16), np.zeros(16, dtype=np.uint8)
    lib().fc_product_rrr_table(t.ctypes.data, cb.ctypes.data, bits.ctypes.data)
    return t, cb, bits


def unrank_table():
    t = np.zeros(32768, dtype=np.uint16)
    lib().fc_unrank_table(t.ctypes.data)
    return t


def utf8_convert(data: bytes):
    """The device-side decoder of utf8_lane.h run on the host: -> (chars uint16[n], 0, 0) or (None, status, value)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    out = np.zeros(max(1, buf.size), dtype=np.uint16)
    val = C.c_int32(0)
    n = lib().fc_utf8_convert(buf.ctypes.data if buf.size else None, buf.size, out.ctypes.data, C.byref(val))
    if n < 0:
        return None, int(-n), int(val.value)
    return out[:n].copy(), 0, 0


def pack_narrow(chars: np.ndarray, n_groups: int, parts: int):
    """host_pack.hpp: the pack pool narrows `chars` (uint16) to bytes in n_groups chunks -> (bytes uint8[n], OR of each chunk's chars)"""
    chars = np.ascontiguousarray(chars, dtype=np.uint16) if chars.dtype != np.uint16 else chars
    out = np.zeros(chars.size, dtype=np.uint8)
    wide = np.zeros(n_groups, dtype=np.uint32)
    lib().fc_pack_narrow(chars.ctypes.data, chars.size, n_groups, parts, out.ctypes.data, wide.ctypes.data)
    return out, wide
