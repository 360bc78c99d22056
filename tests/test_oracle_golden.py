"""Pins the CPU oracle (oracle/oracle.cpp) against the known answers of the reference's own unit
tests (tests/golden/known_answers.json, transcribed by tests/golden/make_golden.py) and against
naive text scans in the style of the reference's test oracle (util/Util.java:108-279)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import CASE_NAMES, get_case, make_patterns

import pyoracle
from index4j_b200.builder import build_index, build_rrr, build_wfbb

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "known_answers.json")) as fh:
    G = json.load(fh)
FIXTURE = os.path.join("/root/reference", G["fm"]["fixture"]["path"])


class Wfbb:
    def __init__(self, symbols, rate=64):
        self.blob = build_wfbb(symbols, rate)
        self.h = C.c_void_p()
        buf = np.frombuffer(self.blob, dtype=np.uint8)
        assert pyoracle.lib().orc_wfbb_load(buf.ctypes.data, buf.size, C.byref(self.h)) == 0

    def rank(self, pos, sym):
        out = C.c_int64()
        assert pyoracle.lib().orc_wfbb_rank(self.h, pos, sym, C.byref(out)) == 0
        return out.value

    def inverse_select(self, pos):
        out = C.c_int64()
        assert pyoracle.lib().orc_wfbb_inverse_select(self.h, pos, C.byref(out)) == 0
        return out.value

    def __del__(self):
        pyoracle.lib().orc_wfbb_free(self.h)


class Rrr:
    def __init__(self, bits, sample):
        self.blob = build_rrr(bits, sample)
        self.h = C.c_void_p()
        buf = np.frombuffer(self.blob, dtype=np.uint8)
        assert pyoracle.lib().orc_rrr_load(buf.ctypes.data, buf.size, C.byref(self.h)) == 0

    def rank_ones(self, p):
        return pyoracle.lib().orc_rrr_rank_ones(self.h, p)

    def access(self, p):
        return pyoracle.lib().orc_rrr_access(self.h, p)

    def __del__(self):
        pyoracle.lib().orc_rrr_free(self.h)


def u16(s):
    return np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16)


def test_wfbb_known_answers():
    g = G["wfbb"]
    t = u16(g["smaller_text"])
    w = Wfbb(t)
    for pos, ch, want in g["rank"]:
        assert w.rank(t.size if pos == "len" else pos, ord(ch)) == want
    s = Wfbb(u16(g["single_symbol"]["text"]))
    for pos, ch, want in g["single_symbol"]["rank"]:
        assert s.rank(pos, ord(ch)) == want
    a = Wfbb(np.full(g["all_ones"]["n"], g["all_ones"]["value"], dtype=np.uint16))
    for pos, want in g["all_ones"]["inverse_select"]:
        assert (a.inverse_select(pos) & 0xFFFF) == want
    r = g["rank_out_of_bounds"]
    arr = np.full(r["n"], r["fill"], dtype=np.uint16)
    arr[r["at"]] = r["value"]
    assert Wfbb(arr).rank(*r["query"]) == r["expected"]
    for case in g["large_blocks"]:
        arr = np.full(case["n"], 1, dtype=np.uint16)  # 'b' -> 1, 'a' -> 2 (any monotonic mapping gives the same count)
        arr[case["at"]] = 2
        assert Wfbb(arr).rank(case["query_pos"], 2) == case["expected"]


def test_wfbb_inverse_select_recovers_text_and_rank_matches_scan():
    rng = np.random.default_rng(0)
    from index4j_b200.builder import map_text
    t = map_text(get_case("multi400k_sr8").text[:150_000])[0][:-1]  # first-appearance codes, like the reference's tests (:194-235)
    w = Wfbb(t, 16)
    for i in rng.integers(0, t.size, 3000):
        v = w.inverse_select(int(i))
        assert (v & 0xFFFF) == int(t[i])
        if i > 0:
            assert (v >> 32) == int((t[:i] == t[i]).sum())
    for _ in range(1000):
        p = int(rng.integers(0, t.size + 1))
        c = int(t[int(rng.integers(0, t.size))])
        assert w.rank(p, c) == int((t[:p] == c).sum())


def test_rrr_known_answers():
    g = G["rrr"]["small"]
    bits = np.zeros(g["length"], dtype=np.uint8)
    bits[g["ones"]] = 1
    r = Rrr(bits, g["sample"])
    for p, want in g["access"]:
        assert r.access(p) == want
    for p, want in g["rank_ones"]:
        assert r.rank_ones(p) == want
    for p, want in g["rank_zeroes"]:
        assert p - r.rank_ones(p) == want
    c = G["rrr"]["corner"]
    bits = np.unpackbits(np.array(c["ints"], dtype="<u4").view(np.uint8), bitorder="little")
    r = Rrr(bits, c["sample"])
    for p, want in c["rank_ones"]:
        assert r.rank_ones(p) == want
    for p, want in c["rank_zeroes"]:
        assert p - r.rank_ones(p) == want


@pytest.mark.parametrize("sample", [1, 4, 15, 32, 64, 256])
def test_rrr_matches_plain_rank(sample):
    rng = np.random.default_rng(sample)
    for n, dens in ((32, 0.5), (1000, 0.03), (77777, 0.5), (320000, 0.9)):
        bits = (rng.random(n) < dens).astype(np.uint8)
        r = Rrr(bits, sample)
        cum = np.concatenate([[0], np.cumsum(bits)])
        for p in np.concatenate([rng.integers(0, n, 400), [0, n - 1]]):
            assert r.rank_ones(int(p)) == cum[p] and r.access(int(p)) == bits[p]
        assert r.rank_ones(n) == cum[n] and r.rank_ones(n + 100) == cum[n] and r.rank_ones(-3) == 0
        assert r.access(n) == -1  # IllegalArgumentException in the reference


def test_rrr_tables_sha256():
    import hashlib
    t = pyoracle.rrr_inverse_table()
    assert hashlib.sha256(t.astype("<u2").tobytes()).hexdigest().startswith("314a5a51")  # SURVEY.md §8(a) a12, hashed from the Java literal


def test_three_line_string_all_sample_rates():
    g = G["fm"]["three_lines"]
    text = u16(g["text"])
    for sr in g["sample_rates"]:
        o = pyoracle.OracleFmIndex(build_index(text, sr))
        for seed in g["seeds"]:
            for mode in (0, 1, 2):
                got = o.extract_until_boundary(seed, 100, 10, mode)
                assert np.array_equal(got, pyoracle.naive_extract_until_boundary(text, seed, 10, mode)), (sr, seed, mode)


def _two_line_text():
    lines = G["fm"]["two_log_lines"]
    rest = "\n081109 203615 148 INFO dfs.DataNode$PacketResponder: PacketResponder 2 for block blk_-6952295868487656571 terminating\n"
    return lines, u16(lines[0] + "\n" + lines[1] + rest * 40)


def test_two_golden_log_lines_and_overflow_messages():
    lines, text = _two_line_text()
    o = pyoracle.OracleFmIndex(build_index(text, 32))
    first = o.extract_until_boundary(5, 300, 10, 0)
    assert first.tobytes().decode("utf-16-le") == lines[0]
    second = o.extract_until_boundary(first.size + 1 + 2, 300, 10, 0)
    assert second.tobytes().decode("utf-16-le") == lines[1]
    d = G["fm"]["does_not_fit"]
    for mode, key in ((0, "both"), (1, "left"), (2, "right")):
        with pytest.raises(pyoracle.JavaException) as e:
            o.extract_until_boundary(d["from"], d["dst_len"], 10, mode)
        assert e.value.status == 8 and e.value.n == d[key]
        assert str(e.value) == "Extraction does not fit in the supplied destination. Currently extracted: %d" % d[key]
    with pytest.raises(pyoracle.JavaException, match="size zero"):
        o.extract_until_boundary(50, 0, 10, 0)
    with pytest.raises(pyoracle.JavaException, match="Boundary does not exist"):
        o.extract_until_boundary(50, 100, 0x4E01, 0)


@pytest.mark.skipif(not os.path.exists(FIXTURE), reason="reference fixture not mounted (read in place, never copied)")
def test_reference_log_fixture_known_answers():
    with open(FIXTURE, encoding="utf-8") as fh:
        s = fh.read()
    text = u16(s)
    assert len(set(text.tolist())) == G["fm"]["fixture"]["distinct_units"]
    o = pyoracle.OracleFmIndex(build_index(text, 32))
    assert o.getInputLength() == text.size + 1                      # FmIndexTest.java:567
    assert o.getAlphabetLength() == len(set(text.tolist())) + 1     # :568-576
    li = G["fm"]["locate_info"]
    assert o.locate(li["pattern"], 0, 4, li["max"], li["max"]).size == li["expected"]  # :195-200
    lines = G["fm"]["two_log_lines"]
    first = o.extract_until_boundary(5, 300, 10, 0)
    assert first.tobytes().decode("utf-16-le") == lines[0]
    second = o.extract_until_boundary(first.size + 1 + 2, 300, 10, 0)
    assert second.tobytes().decode("utf-16-le") == lines[1]
    d = G["fm"]["does_not_fit"]
    for mode, key in ((0, "both"), (1, "left"), (2, "right")):
        with pytest.raises(pyoracle.JavaException) as e:
            o.extract_until_boundary(d["from"], d["dst_len"], 10, mode)
        assert (e.value.status, e.value.n) == (8, d[key])
    # the reference's property tests on the fixture (FmIndexTest.java:104-129,181-193,360-374,498-562), seeds < len-100
    rng = np.random.default_rng(42)
    for sr in (1, 2, 4, 8, 16):
        o = pyoracle.OracleFmIndex(build_index(text, sr))
        for _ in range(40):
            ln = int(rng.integers(1, 32))
            st = int(rng.integers(0, text.size - ln))
            p = text[st: st + ln]
            loc = pyoracle.naive_locations(text, p)
            assert o.count(p) == loc.size
            assert np.array_equal(np.sort(o.locate(p, max_matches=10000, cap=max(loc.size, 1))), loc)
            a = int(rng.integers(0, text.size - 100))
            assert np.array_equal(o.extract(a, a + ln), text[a: a + ln])
            for mode in (0, 1, 2):
                assert np.array_equal(o.extract_until_boundary(a, 32768, 10, mode), pyoracle.naive_extract_until_boundary(text, a, 10, mode))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_text_scans(name):
    """The oracle restates Java line by line INCLUDING its rank fallback bugs (SURVEY.md Q3 and the
    run-block variant), so on texts that trigger them a count may legitimately differ from a text
    scan; such differences must be rare and are reported, never hidden."""
    case = get_case(name)
    chars, off = make_patterns(case.text, 400, 1, 31, seed=101, absent_frac=0.1)
    counts, status = case.oracle.count_batch(chars, off, threads=4)
    assert not status.any()
    diff = 0
    for i in range(counts.size):
        p = chars[int(off[i]): int(off[i + 1])]
        diff += int(counts[i] != pyoracle.naive_count(case.text, p))
    if name.startswith("tiny"):
        assert diff <= counts.size // 4
    else:
        assert diff == 0
    rng = np.random.default_rng(7)
    n = case.text.size
    for _ in range(60):
        a = int(rng.integers(0, n - 120))
        b = a + int(rng.integers(0, 100))
        assert np.array_equal(case.oracle.extract(a, b), case.text[a:b])
