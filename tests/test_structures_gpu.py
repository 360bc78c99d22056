"""The stand-alone succinct structures on the GPU (fmgpu_wavelet_load_serialized / fmgpu_rrr_load_serialized): the reference's own
known answers (WaveletFixedBlockBoostingTest.java:50-132, RrrVectorTest.java:70-122, tests/golden/known_answers.json) and
randomised parity with the oracle's restatement of WaveletFixedBlockBoosting / RrrVector over the same serialized bytes."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import get_case

import pyoracle
from index4j_b200 import RrrVector, WaveletFixedBlockBoosting
from index4j_b200.builder import build_rrr, build_wfbb, map_text
from index4j_b200.fm_index import FmIndexError

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "known_answers.json")) as fh:
    G = json.load(fh)


def u16(s):
    return np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16)


class OracleWfbb:
    def __init__(self, blob):
        self.h = C.c_void_p()
        buf = np.frombuffer(blob, dtype=np.uint8)
        assert pyoracle.lib().orc_wfbb_load(buf.ctypes.data, buf.size, C.byref(self.h)) == 0

    def rank(self, pos, sym):
        out = C.c_int64()
        st = pyoracle.lib().orc_wfbb_rank(self.h, pos, sym, C.byref(out))
        return out.value, st

    def inverse_select(self, pos):
        out = C.c_int64()
        st = pyoracle.lib().orc_wfbb_inverse_select(self.h, pos, C.byref(out))
        return out.value, st

    def __del__(self):
        pyoracle.lib().orc_wfbb_free(self.h)


def gpu_wfbb(symbols, rate=64):
    blob = build_wfbb(symbols, rate)
    return WaveletFixedBlockBoosting.read(blob), OracleWfbb(blob)


def test_wfbb_known_answers_on_gpu():
    g = G["wfbb"]
    t = u16(g["smaller_text"])
    w, _ = gpu_wfbb(t)
    for pos, ch, want in g["rank"]:
        assert w.rank(t.size if pos == "len" else pos, ord(ch)) == want
    s, _ = gpu_wfbb(u16(g["single_symbol"]["text"]))
    for pos, ch, want in g["single_symbol"]["rank"]:
        assert s.rank(pos, ord(ch)) == want
    a, _ = gpu_wfbb(np.full(g["all_ones"]["n"], g["all_ones"]["value"], dtype=np.uint16))
    for pos, want in g["all_ones"]["inverse_select"]:
        assert (a.inverseSelect(pos) & 0xFFFF) == want
    r = g["rank_out_of_bounds"]
    arr = np.full(r["n"], r["fill"], dtype=np.uint16)
    arr[r["at"]] = r["value"]
    assert gpu_wfbb(arr)[0].rank(*r["query"]) == r["expected"]
    for case in g["large_blocks"]:
        arr = np.full(case["n"], 1, dtype=np.uint16)
        arr[case["at"]] = 2
        assert gpu_wfbb(arr)[0].rank(case["query_pos"], 2) == case["expected"]


@pytest.mark.parametrize("kind", ["log", "multiscript", "tiny"])
def test_wfbb_matches_oracle(kind):
    rng = np.random.default_rng(3)
    if kind == "log":
        t = map_text(get_case("log1m_sr32").text[:600_000])[0][:-1]
    elif kind == "multiscript":
        t = map_text(get_case("multi400k_sr8").text[:300_000])[0][:-1]
    else:
        t = rng.integers(1, 5, 400_000).astype(np.uint16)
        t[100_000:180_000] = 3  # long runs: single-symbol blocks
    w, o = gpu_wfbb(t, 16)
    assert w.size() == t.size
    n = t.size
    pos = np.concatenate([rng.integers(0, n + 1, 20000), [0, 1, n - 1, n, n + 7, -1]]).astype(np.int64)
    sym = t[rng.integers(0, n, pos.size)].astype(np.int32)
    sym[:200] = rng.integers(0, int(t.max()) + 3, 200)
    sym[200:220] = -1
    got, st = w.rank_batch(pos, sym)
    for i in range(pos.size):
        want, wst = o.rank(int(pos[i]), int(sym[i]))
        assert st[i] == wst and (wst or got[i] == want), (int(pos[i]), int(sym[i]), want, int(got[i]), wst, int(st[i]))
    ipos = np.concatenate([rng.integers(0, n, 20000), [0, n - 1, -1, n]]).astype(np.int64)
    got, st = w.inverse_select_batch(ipos)
    for i in range(ipos.size):
        p = int(ipos[i])
        if 0 <= p < n:
            want, wst = o.inverse_select(p)
            assert wst == 0 and st[i] == 0 and int(got[i]) == want, p
            assert (int(got[i]) & 0xFFFF) == int(t[p]) or int(t[p]) > 255  # the symbol itself (low byte only in run blocks, Q1)
        else:
            assert st[i] == 9
    # an FmIndex call on a wavelet handle is refused, not answered
    from index4j_b200 import FmIndex
    alias = FmIndex(w._h, w._lib)
    try:
        with pytest.raises(FmIndexError):
            alias.count_batch(np.array([1], dtype=np.uint16), np.array([0, 1], dtype=np.uint64))
    finally:
        alias._h = None  # the handle belongs to `w`


def test_rrr_known_answers_on_gpu():
    g = G["rrr"]["small"]
    bits = np.zeros(g["length"], dtype=np.uint8)
    bits[g["ones"]] = 1
    r = RrrVector.read(build_rrr(bits, g["sample"]))
    for p, want in g["access"]:
        assert int(r.access(p)) == want
    for p, want in g["rank_ones"]:
        assert r.rankOnes(p) == want
    for p, want in g["rank_zeroes"]:
        assert r.rankZeroes(p) == want
    c = G["rrr"]["corner"]
    bits = np.unpackbits(np.array(c["ints"], dtype="<u4").view(np.uint8), bitorder="little")
    r = RrrVector.read(build_rrr(bits, c["sample"]))
    for p, want in c["rank_ones"]:
        assert r.rankOnes(p) == want
    for p, want in c["rank_zeroes"]:
        assert r.rankZeroes(p) == want


@pytest.mark.parametrize("sample", [1, 15, 32, 256])
def test_rrr_matches_oracle_and_plain_rank(sample):
    rng = np.random.default_rng(sample)
    for n, dens in ((32, 0.5), (1000, 0.03), (77777, 0.5), (320000, 0.9), (500000, 1 / 32)):
        bits = (rng.random(n) < dens).astype(np.uint8)
        blob = build_rrr(bits, sample)
        r = RrrVector.read(blob)
        cum = np.concatenate([[0], np.cumsum(bits)])
        pos = np.concatenate([rng.integers(0, n, 3000), [0, n - 1, n, n + 100, -3]]).astype(np.int32)
        rk, bit, st = r.rank_access_batch(pos)
        for i, p in enumerate(pos):
            p = int(p)
            if 0 <= p < n:
                assert st[i] == 0 and rk[i] == cum[p] and bit[i] == bits[p], (n, p)
            else:
                assert st[i] == 11 and rk[i] == (0 if p < 0 else cum[n])  # RrrVector.java:316-323, :360-365
        with pytest.raises(ValueError):
            r.access(n)
        r.close()
