"""Host-side mirror of ``com.dynatrace.fm.FmIndex`` over the C ABI of ``libfmgpu.so``.

The reference class (indices/src/main/java/com/dynatrace/fm/FmIndex.java) exposes
``count`` :443/:455, ``locate`` :487/:504, ``extract`` :564, ``extractUntilBoundary`` :640,
``extractUntilBoundaryLeft`` :772, ``extractUntilBoundaryRight`` :844, ``getInputLength`` :929,
``getAlphabetLength`` :939 and ``read`` :983.  This module keeps those names, argument meanings
and error behaviour (Java exceptions become :class:`FmIndexError` subclasses carrying the same
message) and adds the batched forms the GPU engine is built for.  Everything goes through the
``extern "C"`` functions of ``include/fmgpu.h`` with plain pointers — exactly what a Java host binds
through Panama FFM.  There is no CPU path: if the CUDA library or a GPU is missing
the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

MODE_BOTH, MODE_LEFT, MODE_RIGHT = 0, 1, 2

_MESSAGES = {
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
    11: "Out of range access",
    12: "locate does not terminate in the reference (LF walk in a cycle behind a truncated run-block symbol)",
}


class FmIndexError(RuntimeError):
    """A Java RuntimeException of the reference, or a call-level failure of the native library."""

    def __init__(self, message: str, status: int = 0, n: int = 0):
        super().__init__(message)
        self.status = status
        self.n = n


class FmIndexIllegalArgument(FmIndexError, ValueError):
    """IllegalArgumentException of the reference (status 6 and 7)."""


class FmIndexOutOfBounds(FmIndexError, IndexError):
    """ArrayIndexOutOfBoundsException of the reference (status 9)."""


def raise_status(status: int, n: int = 0):
    msg = _MESSAGES.get(int(status), "status %d" % status).format(n=int(n))
    if status in (6, 7, 11):
        raise FmIndexIllegalArgument(msg, int(status), int(n))
    if status == 9:
        raise FmIndexOutOfBounds(msg, int(status), int(n))
    raise FmIndexError(msg, int(status), int(n))


_lib = None


def native():
    """The loaded ``libfmgpu.so`` (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_build.GPU_LIB):
            raise FmIndexError("libfmgpu.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(_build.GPU_LIB)
        vp, i32, u32, u64, u16 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_uint16
        L.fmgpu_last_error.restype = C.c_char_p
        L.fmgpu_version.restype = C.c_char_p
        L.fmgpu_index_load_serialized.argtypes = [vp, C.c_size_t, vp, C.POINTER(vp)]
        L.fmgpu_index_free.argtypes = [vp]
        for f in ("fmgpu_input_length", "fmgpu_alphabet_length", "fmgpu_sample_rate", "fmgpu_extract_enabled", "fmgpu_device"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = i32
        L.fmgpu_device_bytes.argtypes = [vp]
        L.fmgpu_device_bytes.restype = u64
        L.fmgpu_layout_bytes.argtypes = [vp, vp]
        L.fmgpu_count_batch.argtypes = [vp, vp, vp, u32, vp, vp]
        L.fmgpu_count_batch_device.argtypes = [vp, vp, vp, u64, u32, vp, vp, vp]
        L.fmgpu_locate_batch.argtypes = [vp, vp, vp, u32, i32, vp, vp, vp, u64, vp]
        L.fmgpu_count_batch_utf8.argtypes = [vp, vp, vp, u32, vp, vp]
        L.fmgpu_count_batch_utf8_device.argtypes = [vp, vp, vp, u64, u32, vp, vp, vp]
        L.fmgpu_locate_batch_utf8.argtypes = [vp, vp, vp, u32, i32, vp, vp, vp, u64, vp]
        L.fmgpu_locate_batch_device.argtypes = [vp, vp, vp, u64, u32, i32, vp, vp, vp, u64, vp, C.POINTER(u64), vp]
        L.fmgpu_extract_batch.argtypes = [vp, vp, vp, u32, vp, vp, i32, vp, vp]
        L.fmgpu_extract_batch_device.argtypes = [vp, vp, vp, u32, vp, vp, i32, vp, vp, vp]
        L.fmgpu_extract_until_boundary_batch.argtypes = [vp, vp, u32, u16, i32, i32, i32, vp, vp, vp]
        L.fmgpu_extract_until_boundary_batch_device.argtypes = [vp, vp, u32, u16, i32, i32, i32, vp, vp, vp, vp]
        L.fmgpu_extract_records_batch.argtypes = [vp, vp, u32, u16, i32, vp, vp, vp, vp, u64, C.POINTER(u64)]
        L.fmgpu_extract_records_batch_device.argtypes = [vp, vp, u32, u16, i32, vp, vp, vp, vp, u64, C.POINTER(u64), vp]
        L.fmgpu_locate_records_batch.argtypes = [vp, vp, vp, u32, i32, u16, i32, vp, vp, vp, vp, vp, vp, vp, u64, vp, u64, C.POINTER(u64)]
        L.fmgpu_num_devices.argtypes = [vp]
        L.fmgpu_num_devices.restype = i32
        L.fmgpu_device_at.argtypes = [vp, i32]
        L.fmgpu_device_at.restype = i32
        L.fmgpu_host_pack_threads.restype = i32
        L.fmgpu_host_register.argtypes = [vp, C.c_size_t]
        L.fmgpu_host_unregister.argtypes = [vp]
        L.fmgpu_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.fmgpu_host_free.argtypes = [vp]
        L.fmgpu_wavelet_load_serialized.argtypes = [vp, u64, vp, vp]
        L.fmgpu_rrr_load_serialized.argtypes = [vp, u64, vp, vp]
        L.fmgpu_rrr_rank_access_batch.argtypes = [vp, vp, u32, vp, vp, vp]
        L.fmgpu_wavelet_rank_batch.argtypes = [vp, vp, vp, u32, vp, vp]
        L.fmgpu_wavelet_inverse_select_batch.argtypes = [vp, vp, u32, vp, vp]
        L.fmgpu_last_stats.argtypes = [vp, vp]
        L.fmgpu_last_stats_ex.argtypes = [vp, vp, C.c_uint32]
        L.fmgpu_set_timing.argtypes = [vp, i32]
        L.fmgpu_set_stats.argtypes = [vp, i32]
        L.fmgpu_set_start_table.argtypes = [vp, i32]
        L.fmgpu_start_table_q.argtypes = [vp]
        L.fmgpu_set_locate_dense.argtypes = [vp, i32]
        L.fmgpu_locate_sample_rate.argtypes = [vp]
        L.fmgpu_locate_sample_rate.restype = i32
        L.fmgpu_dense_sample_bytes.argtypes = [vp]
        L.fmgpu_dense_sample_bytes.restype = u64
        L.fmgpu_search_kernel_ms.argtypes = [vp, u32, C.POINTER(C.c_float)]
        L.fmgpu_kernel_ms.argtypes = [vp, i32, u32, u32, C.POINTER(C.c_float)]
        _lib = L
    return _lib


class _Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("host_threads", C.c_int32), ("n_devices", C.c_int32), ("locate_sample_rate", C.c_int32),
                ("devices", C.POINTER(C.c_int32)), ("reserved", C.c_uint64 * 1)]


def pinned_empty(n: int, dtype) -> np.ndarray:
    """A page-locked numpy array from ``fmgpu_host_alloc`` (what a host hands to the ``*_into`` calls so that the copies overlap
    with the kernels and run at the PCIe rate).  The memory is released when the array (and every view of it) is gone."""
    lib = native()
    dt = np.dtype(dtype)
    nbytes = max(int(n) * dt.itemsize, 1)
    p = C.c_void_p()
    if lib.fmgpu_host_alloc(nbytes, C.byref(p)) != 0:
        raise FmIndexError(lib.fmgpu_last_error().decode())
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    _PINNED[p.value] = buf  # keeps the ctypes view alive; freed by pinned_free / at exit
    return arr


_PINNED = {}


def pinned_free(arr: np.ndarray):
    addr = arr.ctypes.data
    if _PINNED.pop(addr, None) is not None:
        native().fmgpu_host_free(addr)


def _u16(a) -> np.ndarray:
    if isinstance(a, str):
        return np.frombuffer(a.encode("utf-16-le", "surrogatepass"), dtype=np.uint16).copy()
    return np.ascontiguousarray(a, dtype=np.uint16)


def concat_patterns(patterns):
    """list of str / uint16 arrays -> (chars uint16[total], pat_off uint64[n+1])"""
    arrs = [_u16(p) for p in patterns]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([a.size for a in arrs])
    chars = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint16)
    return np.ascontiguousarray(chars, dtype=np.uint16), off


class FmIndex:
    """GPU-resident FM-index with the reference's query API."""

    def __init__(self, handle, lib):
        self._h = handle
        self._lib = lib

    # --- construction --------------------------------------------------------------------
    @classmethod
    def read(cls, serialized, device: int | None = None, host_threads: int = 0, devices=None, locate_sample_rate: int = 0) -> "FmIndex":
        """``Serialization.readFromByteArray(FmIndex::read, bytes)`` (Serialization.java:89, FmIndex.java:983).

        ``locate_sample_rate``: device-side denser sampling of the SA rows for locate (include/fmgpu.h): > 0 the requested rate,
        0 = the library's default, < 0 = the index's own sampling only.

        ``devices``: a list of CUDA ordinals (or ``"all"``) — the index is replicated on each of them and every host-pointer batch
        call is cut into one slice per device."""
        lib = native()
        buf = np.frombuffer(serialized, dtype=np.uint8)
        opts = _Opts(-1 if device is None else int(device), int(host_threads))
        opts.locate_sample_rate = int(locate_sample_rate)
        if devices is not None:
            if isinstance(devices, str):
                opts.n_devices = -1
            else:
                arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
                opts.n_devices = len(devices)
                opts.devices = C.cast(arr, C.POINTER(C.c_int32))
        h = C.c_void_p()
        rc = lib.fmgpu_index_load_serialized(buf.ctypes.data, buf.size, C.byref(opts), C.byref(h))
        if rc != 0:
            msg = lib.fmgpu_last_error().decode()
            if rc == -2:
                raise IOError(msg)  # IOException of the reference (incompatible serial version, truncated stream)
            raise FmIndexError(msg)
        return cls(h, lib)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fmgpu_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise FmIndexError(self._lib.fmgpu_last_error().decode(), status=rc)

    # --- metadata --------------------------------------------------------------------------
    def getInputLength(self) -> int:
        return self._lib.fmgpu_input_length(self._h)

    def getAlphabetLength(self) -> int:
        return self._lib.fmgpu_alphabet_length(self._h)

    @property
    def sample_rate(self) -> int:
        return self._lib.fmgpu_sample_rate(self._h)

    @property
    def device(self) -> int:
        return self._lib.fmgpu_device(self._h)

    @property
    def devices(self) -> list:
        return [int(self._lib.fmgpu_device_at(self._h, i)) for i in range(int(self._lib.fmgpu_num_devices(self._h)))]

    def device_bytes(self) -> int:
        return int(self._lib.fmgpu_device_bytes(self._h))

    def layout_bytes(self) -> dict:
        out = np.zeros(8, dtype=np.uint64)
        self._lib.fmgpu_layout_bytes(self._h, out.ctypes.data)
        names = ["cells", "level_sectors", "node_records", "block_descriptors", "occurrence_records", "sampled_rows", "sa_samples", "isa_samples"]
        return {k: int(v) for k, v in zip(names, out)}

    def last_stats(self) -> dict:
        out = np.zeros(16, dtype=np.uint64)
        self._check(self._lib.fmgpu_last_stats_ex(self._h, out.ctypes.data, out.size))
        names = ["ranks", "rank_levels", "lf_steps", "lf_levels", "sampled_tests", "launches", "search_records_loaded", "level_records",
                 "kind0", "rank_cells_const", "rank_cells_run", "rank_cells_throw", "rank_cells_range1k", "rank_cells_list", "rank_cells_bits",
                 "rank_cells_range4k"]
        return {k: int(v) for k, v in zip(names, out)}

    def set_stats(self, enable: bool = True):
        """Work counters of the kernels (``last_stats``) on/off; off by default (production kernels carry none)."""
        self._check(self._lib.fmgpu_set_stats(self._h, int(enable)))

    def set_locate_dense(self, enable: bool = True):
        """locate walks end at the device-side dense samples (default when the index has them) or at the index's own samples."""
        self._check(self._lib.fmgpu_set_locate_dense(self._h, int(enable)))

    @property
    def locate_sample_rate(self) -> int:
        """Rate of the samples locate walks to: the dense rate when dense samples were built, else the index's sampleRate."""
        return int(self._lib.fmgpu_locate_sample_rate(self._h))

    def dense_sample_bytes(self) -> int:
        return int(self._lib.fmgpu_dense_sample_bytes(self._h))

    def set_start_table(self, enable: bool = True):
        """Use (default) / bypass the q-gram start table of the backward search; results are identical either way."""
        self._check(self._lib.fmgpu_set_start_table(self._h, int(enable)))

    def start_table_q(self) -> int:
        return int(self._lib.fmgpu_start_table_q(self._h))

    def set_timing(self, enable: bool = True):
        self._check(self._lib.fmgpu_set_timing(self._h, int(enable)))

    def search_kernel_ms(self, calls_back: int = 0) -> float:
        ms = C.c_float()
        self._check(self._lib.fmgpu_search_kernel_ms(self._h, calls_back, C.byref(ms)))
        return float(ms.value)

    def kernel_ms(self, kind: int, calls_back: int = 0, device_index: int = 0) -> float:
        """Device time of a launch of k_count (kind 0), k_locate (1) or k_extract (2) on replica ``device_index``; needs set_timing."""
        ms = C.c_float()
        self._check(self._lib.fmgpu_kernel_ms(self._h, kind, device_index, calls_back, C.byref(ms)))
        return float(ms.value)

    def count_batch_into(self, chars: np.ndarray, pat_off: np.ndarray, counts: np.ndarray, status: np.ndarray | None = None):
        """``count_batch`` writing into caller-owned (e.g. pinned) buffers — the raw C-ABI call."""
        self._check(self._lib.fmgpu_count_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, pat_off.size - 1, counts.ctypes.data,
                                                status.ctypes.data if status is not None else None))

    def locate_batch_into(self, chars, pat_off, max_hits, n_hits, hit_off, positions, status=None):
        """``locate_batch`` writing into caller-owned (e.g. pinned) buffers — the raw C-ABI call; ``positions`` must hold all hits."""
        self._check(self._lib.fmgpu_locate_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, pat_off.size - 1, max_hits,
                                                 n_hits.ctypes.data, hit_off.ctypes.data, positions.ctypes.data, positions.size,
                                                 status.ctypes.data if status is not None else None))
        return int(hit_off[-1])

    # --- batched API (host buffers) ----------------------------------------------------------
    def count_batch(self, chars, pat_off, return_status: bool = False):
        chars = _u16(chars)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self._check(self._lib.fmgpu_count_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, n, counts.ctypes.data, status.ctypes.data))
        return (counts, status) if return_status else counts

    def count_batch_utf8(self, data, pat_off, return_status: bool = False):
        """UTF-8 byte patterns (``convertBytePatternToCharPattern`` + ``count``); ``pat_off`` are byte offsets.  Where the status
        is 10 the count slot holds the offending code point."""
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data, dtype=np.uint8)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self._check(self._lib.fmgpu_count_batch_utf8(self._h, data.ctypes.data, pat_off.ctypes.data, n, counts.ctypes.data, status.ctypes.data))
        return (counts, status) if return_status else counts

    def count_batch_utf8_into(self, data: np.ndarray, pat_off: np.ndarray, counts: np.ndarray, status: np.ndarray | None = None):
        """``count_batch_utf8`` writing into caller-owned (e.g. pinned) buffers — the raw C-ABI call."""
        self._check(self._lib.fmgpu_count_batch_utf8(self._h, data.ctypes.data, pat_off.ctypes.data, pat_off.size - 1, counts.ctypes.data,
                                                     status.ctypes.data if status is not None else None))

    def count_batch_utf8_device(self, d_bytes, d_pat_off, d_counts, d_status=None, stream: int | None = None):
        """Device-resident UTF-8 patterns (torch uint8 / int64 tensors on this index's GPU)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(d_bytes.device).cuda_stream
        self._check(self._lib.fmgpu_count_batch_utf8_device(self._h, d_bytes.data_ptr(), d_pat_off.data_ptr(), d_bytes.numel(),
                                                            d_pat_off.numel() - 1, d_counts.data_ptr(),
                                                            d_status.data_ptr() if d_status is not None else None, stream))

    def locate_batch(self, chars, pat_off, max_hits: int = -1):
        """-> (n_hits int32[n], hit_off uint64[n+1], positions int32[total], status int32[n])"""
        return self._locate_batch(self._lib.fmgpu_locate_batch, _u16(chars), pat_off, max_hits)

    def locate_batch_utf8(self, data, pat_off, max_hits: int = -1):
        """``locate_batch`` for UTF-8 byte patterns (``pat_off`` are byte offsets)."""
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data, dtype=np.uint8)
        return self._locate_batch(self._lib.fmgpu_locate_batch_utf8, data, pat_off, max_hits)

    def _locate_batch(self, fn, chars, pat_off, max_hits):
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        n_hits = np.zeros(n, dtype=np.int32)
        hit_off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        self._check(fn(self._h, chars.ctypes.data, pat_off.ctypes.data, n, max_hits, n_hits.ctypes.data, hit_off.ctypes.data, None, 0,
                       status.ctypes.data))
        total = int(hit_off[-1])
        positions = np.zeros(max(total, 1), dtype=np.int32)
        if total:
            self._check(fn(self._h, chars.ctypes.data, pat_off.ctypes.data, n, max_hits, n_hits.ctypes.data, hit_off.ctypes.data,
                           positions.ctypes.data, total, status.ctypes.data))
        return n_hits, hit_off, positions[:total], status

    # --- the wavelet structure itself (WaveletFixedBlockBoosting.rank / inverseSelect over alphabet codes) ---------
    def wavelet_rank_batch(self, pos, sym):
        """``WaveletFixedBlockBoosting.rank(position, symbol)`` per query -> (ranks int64[n], status int32[n])"""
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        sym = np.ascontiguousarray(sym, dtype=np.int32)
        out = np.zeros(pos.size, dtype=np.int64)
        st = np.zeros(pos.size, dtype=np.int32)
        self._check(self._lib.fmgpu_wavelet_rank_batch(self._h, pos.ctypes.data, sym.ctypes.data, pos.size, out.ctypes.data, st.ctypes.data))
        return out, st

    def wavelet_inverse_select_batch(self, pos):
        """``WaveletFixedBlockBoosting.inverseSelect(position)`` per query -> (packed int64[n] = rank << 32 | symbol, status)"""
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        out = np.zeros(pos.size, dtype=np.int64)
        st = np.zeros(pos.size, dtype=np.int32)
        self._check(self._lib.fmgpu_wavelet_inverse_select_batch(self._h, pos.ctypes.data, pos.size, out.ctypes.data, st.ctypes.data))
        return out, st

    def extract_batch(self, start, stop, arena_off=None, offset: int = 0):
        """-> (arena uint16[..], arena_off, len int32[n], status int32[n]); slot i = arena[arena_off[i]:arena_off[i+1]] is the
        reference's ``destination``, the chars land ``offset`` slots into it"""
        start = np.ascontiguousarray(start, dtype=np.int32)
        stop = np.ascontiguousarray(stop, dtype=np.int32)
        n = start.size
        if arena_off is None:
            arena_off = np.zeros(n + 1, dtype=np.uint64)
            arena_off[1:] = np.cumsum(np.maximum(stop.astype(np.int64) - start.astype(np.int64), 0) + int(offset))
        arena_off = np.ascontiguousarray(arena_off, dtype=np.uint64)
        arena = np.zeros(max(int(arena_off[-1]), 1), dtype=np.uint16)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        self._check(self._lib.fmgpu_extract_batch(self._h, start.ctypes.data, stop.ctypes.data, n, arena.ctypes.data, arena_off.ctypes.data,
                                                  int(offset), ln.ctypes.data, st.ctypes.data))
        return arena, arena_off, ln, st

    def extract_until_boundary_batch(self, frm, boundary, dst_len: int, mode: int = MODE_BOTH, offset: int = 0):
        """-> (arena uint16[n, dst_len], len int32[n], status int32[n]); row i is the reference's ``destination`` (dst_len chars),
        the record starts at column ``offset``"""
        frm = np.ascontiguousarray(frm, dtype=np.int32)
        n = frm.size
        b = ord(boundary) if isinstance(boundary, str) else int(boundary)
        arena = np.zeros((n, max(dst_len, 1)), dtype=np.uint16)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        self._check(self._lib.fmgpu_extract_until_boundary_batch(self._h, frm.ctypes.data, n, b, dst_len, int(offset), mode,
                                                                 arena.ctypes.data if dst_len > 0 else None, ln.ctypes.data, st.ctypes.data))
        return arena, ln, st

    # --- fused locate -> extractUntilBoundary: every distinct record read once ----------------------------------
    def extract_records_batch(self, frm, boundary, dst_len: int, rec_cap: int | None = None):
        """``extractUntilBoundary(from, new char[dst_len], 0, boundary)`` for every hit, each distinct record extracted once.
        -> (rec_index int32[n], len int32[n], status int32[n], records uint16[n_records, dst_len]); the record of hit h is
        ``records[rec_index[h], :len[h]]``."""
        frm = np.ascontiguousarray(frm, dtype=np.int32)
        n = frm.size
        b = ord(boundary) if isinstance(boundary, str) else int(boundary)
        cap = n if rec_cap is None else int(rec_cap)
        idx = np.zeros(n, dtype=np.int32)
        ln = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        arena = np.zeros((max(cap, 1), max(dst_len, 1)), dtype=np.uint16)
        n_rec = C.c_uint64()
        self._check(self._lib.fmgpu_extract_records_batch(self._h, frm.ctypes.data, n, b, dst_len, idx.ctypes.data, ln.ctypes.data, st.ctypes.data,
                                                          arena.ctypes.data, cap, C.byref(n_rec)))
        return idx, ln, st, arena[: int(n_rec.value)]

    def extract_records_batch_device(self, d_from, boundary, dst_len: int, d_rec_index, d_len, d_status, d_rec_arena, stream: int | None = None) -> int:
        """device-resident form; returns the number of distinct records (synchronizes the stream once)"""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(d_from.device).cuda_stream
        b = ord(boundary) if isinstance(boundary, str) else int(boundary)
        n_rec = C.c_uint64()
        self._check(self._lib.fmgpu_extract_records_batch_device(self._h, d_from.data_ptr(), d_from.numel(), b, dst_len, d_rec_index.data_ptr(),
                                                                 d_len.data_ptr(), d_status.data_ptr(), d_rec_arena.data_ptr(),
                                                                 d_rec_arena.shape[0], C.byref(n_rec), stream))
        return int(n_rec.value)

    def locate_records_batch(self, chars, pat_off, max_hits: int, boundary, dst_len: int):
        """``locate`` + ``extractUntilBoundary`` of every hit in one call.
        -> dict(n_hits, hit_off, pat_status, positions, rec_index, len, status, records)"""
        chars = _u16(chars)
        pat_off = np.ascontiguousarray(pat_off, dtype=np.uint64)
        n = pat_off.size - 1
        b = ord(boundary) if isinstance(boundary, str) else int(boundary)
        n_hits = np.zeros(n, dtype=np.int32)
        hit_off = np.zeros(n + 1, dtype=np.uint64)
        pst = np.zeros(n, dtype=np.int32)
        n_rec = C.c_uint64()
        self._check(self._lib.fmgpu_locate_records_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, n, max_hits, b, dst_len, n_hits.ctypes.data,
                                                         hit_off.ctypes.data, pst.ctypes.data, None, None, None, None, 0, None, 0, C.byref(n_rec)))
        total = int(hit_off[-1])
        pos = np.zeros(max(total, 1), dtype=np.int32)
        idx = np.zeros(max(total, 1), dtype=np.int32)
        ln = np.zeros(max(total, 1), dtype=np.int32)
        st = np.zeros(max(total, 1), dtype=np.int32)
        arena = np.zeros((max(total, 1), max(dst_len, 1)), dtype=np.uint16)
        if total:
            self._check(self._lib.fmgpu_locate_records_batch(self._h, chars.ctypes.data, pat_off.ctypes.data, n, max_hits, b, dst_len,
                                                             n_hits.ctypes.data, hit_off.ctypes.data, pst.ctypes.data, pos.ctypes.data, idx.ctypes.data,
                                                             ln.ctypes.data, st.ctypes.data, total, arena.ctypes.data, total, C.byref(n_rec)))
        return dict(n_hits=n_hits, hit_off=hit_off, pat_status=pst, positions=pos[:total], rec_index=idx[:total], len=ln[:total], status=st[:total],
                    records=arena[: int(n_rec.value)])

    # --- the reference's single-query methods -----------------------------------------------------
    def count(self, pattern, offset: int = 0, length: int | None = None) -> int:
        """``FmIndex.count(char[] pattern, int offset, int length)`` (FmIndex.java:443,455)."""
        p = _u16(pattern)
        if length is None:
            length = p.size - offset
        counts, status = self.count_batch(p[offset: offset + length], np.array([0, max(length, 0)], dtype=np.uint64), True)
        if status[0]:
            raise_status(status[0])
        return int(counts[0])

    def count_utf8(self, pattern: bytes) -> int:
        """``count`` of a UTF-8 byte pattern: ``convertBytePatternToCharPattern`` (FmIndex.java:239-298) then ``count``."""
        counts, status = self.count_batch_utf8(pattern, np.array([0, len(pattern)], dtype=np.uint64), True)
        if status[0]:
            raise_status(status[0], counts[0])
        return int(counts[0])

    def locate(self, pattern, offset: int = 0, length: int | None = None, locations=None, max_matches: int = -1):
        """``FmIndex.locate(char[] pattern, int offset, int length, int[] locations, int maxMatches)`` (:487,:504).

        With ``locations`` (an int32 array) the hits are written into it and their number is returned,
        like Java; without it the located positions are returned as an array.
        """
        p = _u16(pattern)
        if length is None:
            length = p.size - offset
        n_hits, _, pos, status = self.locate_batch(p[offset: offset + length], np.array([0, max(length, 0)], dtype=np.uint64), max_matches)
        if status[0]:
            raise_status(status[0])
        if locations is None:
            return pos
        if pos.size > len(locations):
            raise FmIndexOutOfBounds(_MESSAGES[9], 9)
        locations[: pos.size] = pos
        return int(n_hits[0])

    def extract(self, start: int, stop: int, destination=None, offset: int = 0):
        """``FmIndex.extract(int start, int stop, char[] destination, int offset)`` (:564)."""
        room = len(destination) if destination is not None else max(stop - start, 0) + offset
        arena, _, ln, st = self.extract_batch([start], [stop], np.array([0, max(room, 0)], dtype=np.uint64), offset)
        if st[0]:
            raise_status(st[0])
        k = max(int(ln[0]), 0)
        if destination is None:
            return arena[offset: offset + k].copy()
        destination[offset: offset + k] = arena[offset: offset + k]
        return int(ln[0])

    def _eub(self, frm, destination, offset, boundary, mode):
        if isinstance(destination, int):  # destination = new char[destination]
            dst_len, dest = destination, None
        else:
            dst_len, dest = len(destination), destination
        arena, ln, st = self.extract_until_boundary_batch([frm], boundary, max(dst_len, 0), mode, offset)
        if st[0]:
            raise_status(st[0], ln[0])
        k = max(int(ln[0]), 0)
        if dest is None:
            return arena[0, offset: offset + k].copy()
        dest[offset: offset + k] = arena[0, offset: offset + k]
        return int(ln[0])

    def extractUntilBoundary(self, frm: int, destination, offset: int = 0, boundary="\n"):
        """``FmIndex.extractUntilBoundary(int from, char[] destination, int offset, char boundary)`` (:640)."""
        return self._eub(frm, destination, offset, boundary, MODE_BOTH)

    def extractUntilBoundaryLeft(self, frm: int, destination, offset: int = 0, boundary="\n"):
        return self._eub(frm, destination, offset, boundary, MODE_LEFT)

    def extractUntilBoundaryRight(self, frm: int, destination, offset: int = 0, boundary="\n"):
        return self._eub(frm, destination, offset, boundary, MODE_RIGHT)

    # --- device-resident forms (torch tensors on this index's GPU; asynchronous on the current stream) ----
    def count_batch_device(self, d_chars, d_pat_off, d_counts, d_status=None, stream: int | None = None):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(d_chars.device).cuda_stream
        n = d_pat_off.numel() - 1
        self._check(self._lib.fmgpu_count_batch_device(self._h, d_chars.data_ptr(), d_pat_off.data_ptr(), d_chars.numel(), n,
                                                       d_counts.data_ptr(), d_status.data_ptr() if d_status is not None else None, stream))

    def locate_batch_device(self, d_chars, d_pat_off, max_hits, d_n_hits, d_hit_off, d_positions, d_status=None, stream: int | None = None) -> int:
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(d_chars.device).cuda_stream
        n = d_pat_off.numel() - 1
        total = C.c_uint64()
        self._check(self._lib.fmgpu_locate_batch_device(self._h, d_chars.data_ptr(), d_pat_off.data_ptr(), d_chars.numel(), n, max_hits,
                                                        d_n_hits.data_ptr(), d_hit_off.data_ptr(),
                                                        d_positions.data_ptr() if d_positions is not None else None,
                                                        d_positions.numel() if d_positions is not None else 0,
                                                        d_status.data_ptr() if d_status is not None else None, C.byref(total), stream))
        return int(total.value)

    def extract_until_boundary_batch_device(self, d_from, boundary, dst_len, mode, d_arena, d_len, d_status, stream: int | None = None,
                                            offset: int = 0):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(d_from.device).cuda_stream
        b = ord(boundary) if isinstance(boundary, str) else int(boundary)
        self._check(self._lib.fmgpu_extract_until_boundary_batch_device(self._h, d_from.data_ptr(), d_from.numel(), b, dst_len, int(offset), mode,
                                                                        d_arena.data_ptr(), d_len.data_ptr(), d_status.data_ptr(), stream))

    def extract_batch_device(self, d_start, d_stop, d_arena, d_arena_off, d_len, d_status, stream: int | None = None, offset: int = 0):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(d_start.device).cuda_stream
        self._check(self._lib.fmgpu_extract_batch_device(self._h, d_start.data_ptr(), d_stop.data_ptr(), d_start.numel(), d_arena.data_ptr(),
                                                         d_arena_off.data_ptr(), int(offset), d_len.data_ptr(), d_status.data_ptr(), stream))
