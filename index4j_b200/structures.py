"""GPU-resident mirrors of the reference's stand-alone succinct structures, loaded from their own serialized streams:

* ``WaveletFixedBlockBoosting`` — ``rank(position, symbol)`` / ``inverseSelect(position)``
  (indices/src/main/java/com/dynatrace/wavelet/WaveletFixedBlockBoosting.java:286, :1010, :1305);
* ``RrrVector`` — ``rankOnes`` / ``rankZeroes`` / ``access``
  (indices/src/main/java/com/dynatrace/bitsequence/RrrVector.java:314, :358, :405, :448).

Same C-ABI library as ``FmIndex`` (``fmgpu_wavelet_load_serialized`` / ``fmgpu_rrr_load_serialized``); batched forms take
numpy arrays, the single-query forms raise like the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .fm_index import FmIndexError, _Opts, native, raise_status


class _Handle:
    _loader = ""

    def __init__(self, handle, lib):
        self._h = handle
        self._lib = lib

    @classmethod
    def read(cls, serialized, device: int | None = None, host_threads: int = 0):
        lib = native()
        buf = np.frombuffer(serialized, dtype=np.uint8)
        opts = _Opts(-1 if device is None else int(device), int(host_threads))
        h = C.c_void_p()
        rc = getattr(lib, cls._loader)(buf.ctypes.data, buf.size, C.byref(opts), C.byref(h))
        if rc != 0:
            msg = lib.fmgpu_last_error().decode()
            if rc == -2:
                raise IOError(msg)
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


class WaveletFixedBlockBoosting(_Handle):
    _loader = "fmgpu_wavelet_load_serialized"

    def size(self) -> int:
        return self._lib.fmgpu_input_length(self._h)

    def rank_batch(self, pos, sym):
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        sym = np.ascontiguousarray(sym, dtype=np.int32)
        out = np.zeros(pos.size, dtype=np.int64)
        st = np.zeros(pos.size, dtype=np.int32)
        self._check(self._lib.fmgpu_wavelet_rank_batch(self._h, pos.ctypes.data, sym.ctypes.data, pos.size, out.ctypes.data, st.ctypes.data))
        return out, st

    def inverse_select_batch(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        out = np.zeros(pos.size, dtype=np.int64)
        st = np.zeros(pos.size, dtype=np.int32)
        self._check(self._lib.fmgpu_wavelet_inverse_select_batch(self._h, pos.ctypes.data, pos.size, out.ctypes.data, st.ctypes.data))
        return out, st

    def rank(self, position: int, symbol: int) -> int:
        out, st = self.rank_batch([position], [symbol])
        if st[0]:
            raise_status(st[0])
        return int(out[0])

    def inverseSelect(self, position: int) -> int:
        out, st = self.inverse_select_batch([position])
        if st[0]:
            raise_status(st[0])
        return int(out[0])


class RrrVector(_Handle):
    _loader = "fmgpu_rrr_load_serialized"

    def rank_access_batch(self, pos):
        """-> (rankOnes int32[n], access int32[n] (0/1), status int32[n]: 11 where access throws)"""
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        rk = np.zeros(pos.size, dtype=np.int32)
        bit = np.zeros(pos.size, dtype=np.int32)
        st = np.zeros(pos.size, dtype=np.int32)
        self._check(self._lib.fmgpu_rrr_rank_access_batch(self._h, pos.ctypes.data, pos.size, rk.ctypes.data, bit.ctypes.data, st.ctypes.data))
        return rk, bit, st

    def rankOnes(self, position: int) -> int:
        return int(self.rank_access_batch([position])[0][0])

    def rankZeroes(self, position: int) -> int:  # RrrVector.java:405
        return int(position) - self.rankOnes(position)

    def access(self, position: int) -> bool:
        _, bit, st = self.rank_access_batch([position])
        if st[0]:
            raise_status(st[0])
        return bool(bit[0])
