"""CPU tests of the product's host re-layout (flatten.hpp) and per-lane logic (lane_logic.h,
lf_lane.h): tests/support/libflatcheck.so replays exactly the functions every GPU lane runs, one
lane at a time with plain memory reads, and the results must be bit-exact with the oracle.
"""
import numpy as np
import pytest

from conftest import CASE_NAMES, QUIRK_CASE_NAMES, get_case, make_patterns

ALL_CASES = CASE_NAMES + QUIRK_CASE_NAMES

import flatcheck
import pyoracle


@pytest.fixture(scope="module")
def flats():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = flatcheck.FlatIndexHost(get_case(name).blob)
        return cache[name]

    return get


def test_unranking_reproduces_reference_table():
    assert np.array_equal(flatcheck.unrank_table(), pyoracle.rrr_inverse_table())


def test_product_rrr_tables_match_the_java_literals():
    """The loader's own tables (flatten.hpp, computed by combinatorial unranking; uploaded to the device) against the sha256 of the
    reference's INVERSE_VALUES literal, its CARDINALITY_OFFSETS (RrrVector.java:8692-8698) and BITS_NEEDED (:111-129)."""
    import hashlib
    inv, cbase, bits = flatcheck.product_rrr_table()
    assert hashlib.sha256(inv.astype("<u2").tobytes()).hexdigest().startswith("314a5a51")
    assert np.array_equal(inv, pyoracle.rrr_inverse_table())
    assert cbase.tolist() == [0, 1, 16, 121, 576, 1941, 4944, 9949, 16384, 22819, 27824, 30827, 32192, 32647, 32752, 32767]
    assert bits.tolist() == [1, 4, 7, 9, 11, 12, 13, 13, 13, 13, 12, 11, 9, 7, 4, 1]


@pytest.mark.parametrize("name", ALL_CASES)
def test_rank_cells_match_oracle(flats, name):
    case, f = get_case(name), flats(name)
    rng = np.random.default_rng(1)
    L = case.oracle.getInputLength()
    sigma = case.oracle.getAlphabetLength() + 2
    pos = np.concatenate([rng.integers(0, L + 1, 30000), [0, 1, L - 1, L, L + 5]])
    # positions around block boundaries, where the absent-from-block fallbacks live
    pos = np.concatenate([pos, (rng.integers(1, max(2, L >> 9), 4000) << 9) + rng.integers(-1, 2, 4000)])
    pos = np.clip(pos, 0, L + 5)
    for p in pos:
        s = int(rng.integers(0, sigma))
        try:
            want, st = case.oracle.wfbb_rank(int(p), s), 0
        except pyoracle.JavaException as e:
            want, st = None, e.status
        got_st, got = f.rank(int(p), s)
        assert got_st == st and (st or got == want), (int(p), s, want, got, st, got_st)


@pytest.mark.parametrize("name", ALL_CASES)
def test_inverse_select_records_match_oracle(flats, name):
    """WaveletFixedBlockBoosting.inverseSelect through the block descriptors / level + node records (incl. the low-byte symbol of
    single-symbol blocks, quirk Q1) against the oracle's restatement of :1305-1537."""
    case, f = get_case(name), flats(name)
    rng = np.random.default_rng(5)
    L = case.oracle.getInputLength()
    pos = np.concatenate([rng.integers(0, L, 20000), [0, 1, L - 1], (rng.integers(1, max(2, L >> 9), 2000) << 9) + rng.integers(-1, 2, 2000)])
    for p in np.clip(pos, 0, L - 1):
        st, got = f.inverse_select(int(p))
        assert st == 0 and got == case.oracle.wfbb_inverse_select(int(p)), int(p)
    assert f.inverse_select(-1)[0] == 9 and f.inverse_select(L)[0] == 9


@pytest.mark.parametrize("name", ALL_CASES)
def test_sampled_rows_match_oracle(flats, name):
    case, f = get_case(name), flats(name)
    rng = np.random.default_rng(2)
    L = case.oracle.getInputLength()
    for p in np.concatenate([rng.integers(0, L, 20000), [0, L - 1]]):
        b, r = f.sampled(int(p))
        assert b == case.oracle.sampled_access(int(p)) and r == case.oracle.sampled_rank(int(p)), int(p)


@pytest.mark.parametrize("name", ALL_CASES)
def test_count_and_locate_lanes(flats, name):
    case, f = get_case(name), flats(name)
    chars, off = make_patterns(case.text, 2500, 1, 40, seed=3)
    want, want_st = case.oracle.count_batch(chars, off, threads=4)
    got, got_st, ranges = f.count_batch(chars, off)
    assert np.array_equal(got_st, want_st) and np.array_equal(got, want)
    n_hits, pos, st = case.oracle.locate_batch(chars, off, 50, 50, threads=4)
    rows, exp = [], []
    for i in range(n_hits.size):
        rows += list(range(int(ranges[i, 0]), int(ranges[i, 0]) + int(n_hits[i])))
        exp += list(pos[i, : n_hits[i]])
    got_pos = f.locate_rows(np.array(rows, dtype=np.uint32))  # lane code of k_locate (lf_lane.h)
    assert np.array_equal(got_pos, np.array(exp, dtype=np.int64))


@pytest.mark.parametrize("name", ALL_CASES)
def test_occurrence_cells_cover_all_forms(flats, name):
    """rank through the (block, symbol) occurrence structures (short position list / range lists / bit vector, layout.h) at
    EVERY position of a few blocks and for every symbol of the alphabet — all forms and their record boundaries."""
    case, f = get_case(name), flats(name)
    kinds = f.cell_kinds()  # [normal, const, run, throw, range-1K, list, bits, range-4K]
    assert kinds[0] == 0 and kinds[5] > 0 and kinds[6] > 0
    if name in ("log1m_sr32", "log3m_sr16", "multi400k_sr8"):
        assert kinds[4] > 0 and kinds[7] > 0
    L = case.oracle.getInputLength()
    sigma = case.oracle.getAlphabetLength() + 1
    rng = np.random.default_rng(8)
    starts = [0, max(0, L - 5000)] + [int(x) for x in rng.integers(0, max(1, L - 3000), 3)]
    syms = list(range(min(sigma, 40))) + [int(x) for x in rng.integers(0, sigma, 25)]
    for s0 in starts:
        for p in range(s0, min(L, s0 + 2500), 3):
            for s in syms[:: 1 if p % 30 == 0 else 7]:
                try:
                    want, st = case.oracle.wfbb_rank(p, s), 0
                except pyoracle.JavaException as e:
                    want, st = None, e.status
                got_st, got = f.rank(p, s)
                assert got_st == st and (st or got == want), (p, s, want, got)


@pytest.mark.parametrize("name,q", [("log300k_sr64", 2), ("log300k_sr64", 3), ("tiny600k_sr4", 5), ("log200k_sr1", 2)])
def test_start_table_lanes(name, q):
    """The q-gram start table (pattern_start / start_table_lookup of count_lane.h, table built by the step-by-step search itself):
    counts, statuses and SA ranges equal the oracle's / the table-less search's for patterns of every length around q with known
    and unknown chars in their last q positions."""
    case = get_case(name)
    f = flatcheck.FlatIndexHost(case.blob)
    assert f.build_start_table(q) > 0
    rng = np.random.default_rng(13)
    t = case.text
    pats = []
    for ln in range(1, q + 4):
        for _ in range(250):
            a = int(rng.integers(0, t.size - ln))
            p = t[a: a + ln].copy()
            if rng.random() < 0.3:
                p[int(rng.integers(0, ln))] = rng.choice([0xFFFE, 0, int(t[int(rng.integers(0, t.size))])])
            pats.append(p)
    chars0, off0 = make_patterns(t, 1500, 1, 40, seed=14)
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in pats])
    chars = np.concatenate(pats + [chars0]).astype(np.uint16)
    off = np.concatenate([off, off[-1] + off0[1:]]).astype(np.uint64)
    want, want_st = case.oracle.count_batch(chars, off, threads=4)
    got, got_st, ranges = f.count_batch_table(chars, off)
    assert np.array_equal(got_st, want_st) and np.array_equal(got, want)
    plain, plain_st, plain_ranges = f.count_batch(chars, off)
    hit = want > 0
    assert np.array_equal(ranges[hit], plain_ranges[hit])  # locate starts from these rows


@pytest.mark.parametrize("name", CASE_NAMES)
def test_extract_lanes(flats, name):
    case, f = get_case(name), flats(name)
    rng = np.random.default_rng(4)
    n = case.text.size
    m = 1500
    start = rng.integers(0, n - 200, m).astype(np.int32)
    ln = rng.integers(0, 150, m).astype(np.int32)
    stop = start + ln
    start[:3] = [0, n - 40, n - 1]
    stop[:3] = [70, n, n]
    aoff = np.zeros(m + 1, dtype=np.uint64)
    aoff[1:] = np.cumsum(stop - start)
    arena, got_len, st = f.extract(start, stop, aoff)
    assert not st.any() and np.array_equal(got_len, stop - start)
    for i in range(m):
        assert np.array_equal(arena[int(aoff[i]): int(aoff[i + 1])], case.text[start[i]: stop[i]]), i


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_extract_until_boundary_lanes(flats, name, mode):
    case, f = get_case(name), flats(name)
    n = case.text.size
    rng = np.random.default_rng(5 + mode)
    for dst_len in (512, 40, 10, 3, 1):
        frm = np.concatenate([rng.integers(0, n, 700), np.arange(n - 12, n + 2), np.arange(-1, 5)]).astype(np.int32)
        a1, l1, s1 = case.oracle.extract_until_boundary_batch(frm, 10, dst_len, mode, threads=4)
        a2, l2, s2 = f.eub(frm, 10, dst_len, mode)
        assert np.array_equal(s1, s2), (dst_len, np.flatnonzero(s1 != s2)[:5])
        ok = (s1 == 0) | (s1 == 8)
        assert np.array_equal(l1[ok], l2[ok]), dst_len
        for i in np.flatnonzero(s1 == 0):
            if frm[i] >= n:
                continue
            assert np.array_equal(a1[i, : l1[i]], a2[i, : l1[i]]), (dst_len, i, int(frm[i]))


@pytest.mark.parametrize("name", ["log300k_sr64", "tiny600k_sr4", "nul1m_sr32"])
@pytest.mark.parametrize("offset", [1, 7, 30, 39, 40, 64])
def test_offset_lanes(flats, name, offset):
    """The reference's `offset` argument of extract / extractUntilBoundary* (fm/FmIndex.java:564, :640, :772, :844): where the
    chars land, the capacity tests, the N of "does not fit" and the arraycopy that throws when the left part does not fit."""
    case, f = get_case(name), flats(name)
    n = case.text.size
    rng = np.random.default_rng(70 + offset)
    for mode in (0, 1, 2):
        for dst_len in (40, 512):
            frm = np.concatenate([rng.integers(0, n, 500), np.arange(n - 9, n + 1), np.arange(0, 4)]).astype(np.int32)
            a1, l1, s1 = case.oracle.extract_until_boundary_batch(frm, 10, dst_len, mode, threads=4, offset=offset)
            a2, l2, s2 = f.eub(frm, 10, dst_len, mode, offset=offset)
            assert np.array_equal(s1, s2), (mode, dst_len, np.flatnonzero(s1 != s2)[:5], s1[s1 != s2][:5], s2[s1 != s2][:5])
            ok = (s1 == 0) | (s1 == 8)
            assert np.array_equal(l1[ok], l2[ok]), (mode, dst_len)
            for i in np.flatnonzero(s1 == 0):
                if frm[i] >= n:
                    continue
                assert np.array_equal(a1[i, offset: offset + l1[i]], a2[i, offset: offset + l1[i]]), (mode, dst_len, i, int(frm[i]))
    m = 600
    start = rng.integers(0, n - 100, m).astype(np.int32)
    stop = (start + rng.integers(0, 80, m)).astype(np.int32)
    stride = 90  # destination.length: some items fit behind the offset, some do not
    aoff = (np.arange(m + 1) * stride).astype(np.uint64)
    arena, got_len, st = f.extract(start, stop, aoff, offset=offset)
    w_arena, w_len, w_st = case.oracle.extract_batch(start, stop, stride, threads=4, offset=offset)
    assert np.array_equal(st, w_st) and ((w_st == 5).sum() > 0 or offset < 20) and (w_st == 0).sum() > 0
    for i in np.flatnonzero(w_st == 0):
        assert got_len[i] == w_len[i]
        assert np.array_equal(arena[i * stride + offset: i * stride + offset + w_len[i]], w_arena[i, offset: offset + w_len[i]]), i


def _check_records(case, frm, dst_len, idx, ln, st, records):
    """per hit: exactly what the oracle's extractUntilBoundary(from, new char[dst_len], 0, '\\n') returns / throws"""
    n = case.text.size
    w_arena, w_len, w_st = case.oracle.extract_until_boundary_batch(frm, 10, dst_len, 0, threads=4)
    assert np.array_equal(st, w_st), np.flatnonzero(st != w_st)[:10]
    ok = (w_st == 0) | (w_st == 8)
    assert np.array_equal(ln[ok], w_len[ok])
    for i in np.flatnonzero(w_st == 0):
        if frm[i] >= n or w_len[i] == 0:
            continue
        assert idx[i] >= 0
        assert np.array_equal(records[idx[i], : w_len[i]], w_arena[i, : w_len[i]]), (i, int(frm[i]), int(w_len[i]))
    return int((w_st == 0).sum())


@pytest.mark.parametrize("name", ["log300k_sr64", "log1m_sr32", "tiny600k_sr4", "multi400k_sr8", "nul1m_sr32"])
@pytest.mark.parametrize("dst_len", [512, 130, 37, 5])
def test_record_lanes(flats, name, dst_len):
    """Fused locate -> extractUntilBoundary (scan + claim, one extraction per distinct record, per-hit chunk arithmetic): every hit
    gets what the reference's extractUntilBoundary returns or throws for it — hits clustered in the same records, at boundary
    chars, at both ends of the text."""
    case, f = get_case(name), flats(name)
    n = case.text.size
    rng = np.random.default_rng(90 + dst_len)
    centers = rng.integers(0, n, 150)
    near = (centers[:, None] + rng.integers(-60, 60, (150, 8))).reshape(-1)  # several hits per record
    nl = np.flatnonzero(case.text == 10)
    at_nl = nl[rng.integers(0, nl.size, 60)]
    frm = np.concatenate([near, at_nl, at_nl + 1, at_nl - 1, rng.integers(0, n, 400), np.arange(n - 14, n + 2), np.arange(-1, 8)])
    frm = np.clip(frm, -1, n + 1).astype(np.int32)
    idx, ln, st, records = f.records(frm, 10, dst_len)
    good = _check_records(case, frm, dst_len, idx, ln, st, records)
    if dst_len >= 130 and name.startswith("log"):
        assert good > 500 and records.shape[0] < 0.7 * frm.size  # the clustered hits share records


def _naive_sa(text):
    """suffix array of text + sentinel (sentinel smallest), symbols ordered by first appearance (fm/FmIndex.java:396-435)"""
    order = {}
    for c in text:
        order.setdefault(int(c), len(order) + 1)
    codes = [order[int(c)] for c in text] + [0]
    return sorted(range(len(codes)), key=lambda i: codes[i:])


@pytest.mark.parametrize("n,sr,rate", [(5003, 32, 8), (4096, 16, 4), (1000, 32, 8), (33, 32, 8), (7, 4, 2), (2000, 8, 1), (640, 64, 16)])
def test_dense_samples_build_small(n, sr, rate):
    """Dense-sample build (kernels_dense.cuh) replayed on the host against a naive suffix array: exactly the rows whose suffix
    starts at a multiple of the rate are marked, each with its position — incl. the rows behind the last multiple of
    sampleRate (walk from the sentinel's row) and position 0."""
    from index4j_b200.builder import build_index, gen_log_text
    text = gen_log_text(n, seed=31 + n)
    f = flatcheck.FlatIndexHost(build_index(text, sr))
    built = f.dense_build(rate, n + 1)
    assert built is not None
    marks, dsa = built
    sa = np.array(_naive_sa(text), dtype=np.int64)
    bits = np.unpackbits(marks[:, 1:].copy().view(np.uint8), bitorder="little")[: n + 1].astype(bool)
    assert np.array_equal(bits, sa % rate == 0)
    assert np.array_equal(dsa.astype(np.int64), sa[bits])
    assert np.array_equal(marks[:, 0], np.concatenate([[0], np.cumsum(bits)])[np.arange(marks.shape[0]) * 224])
    got, _ = f.locate_rows_dense(marks, dsa, np.arange(n + 1, dtype=np.uint32))
    assert np.array_equal(got, sa)


@pytest.mark.parametrize("name", ["log1m_sr32", "log3m_sr16", "log300k_sr64", "tiny600k_sr4", "nul1m_sr32"])
def test_dense_samples_locate_lanes(flats, name):
    """locate through the dense marks (lane code of k_locate<., DENSE>) = locate through the index's own samples = the oracle."""
    case, f = get_case(name), flats(name)
    L = case.oracle.getInputLength()
    rate = 8 if case.sample_rate > 8 else 2
    built = f.dense_build(rate, L)
    assert built is not None
    marks, dsa = built
    assert int(marks[-1, 0]) + int(np.unpackbits(marks[-1, 1:].copy().view(np.uint8)).sum()) == (L - 1) // rate + 1
    rows = np.concatenate([np.random.default_rng(4).integers(0, L, 30000), [0, L - 1]]).astype(np.uint32)
    f.counters[:] = 0
    want = f.locate_rows(rows)
    got, steps = f.locate_rows_dense(marks, dsa, rows)
    assert np.array_equal(got, want)
    assert steps < int(f.counters[2])


@pytest.mark.parametrize("name", ["log1m_sr32", "log3m_sr16", "multi400k_sr8", "tiny600k_sr4"])
def test_every_occurrence_form_over_whole_blocks(flats, name):
    """For a few (block, symbol) cells of EVERY occurrence form (position list, range lists over 1024 / 4096 positions, bit
    vector): rank at every record boundary of the block and at a stride of positions in between, against the oracle."""
    case, f = get_case(name), flats(name)
    L = case.oracle.getInputLength()
    for kind in (4, 5, 6, 7):
        for sym, row0, bsize in f.find_cells(kind, 4):
            ps = set(range(row0, min(L, row0 + bsize) + 1, 37))
            for step in (224, 1024, 4096):
                for q in range(row0, min(L, row0 + bsize) + 1, step):
                    ps.update(p for p in (q - 1, q, q + 1) if row0 <= p <= min(L, row0 + bsize))
            for p in sorted(ps):
                got_st, got = f.rank(p, sym)
                assert got_st == 0 and got == case.oracle.wfbb_rank(p, sym), (kind, sym, row0, bsize, p)


@pytest.mark.parametrize("sr", [2, 4, 32])
def test_extract_until_boundary_exhaustive_on_small_texts(sr):
    """extractUntilBoundary{,Left,Right} from EVERY position (and just outside the text) of small multi-line texts, for every
    destination length 1..26 and offsets 0..4: status, returned length / N of "does not fit", contents — the reference's 4-char
    chunk arithmetic incl. its end-of-text rule (quirks Q5 / Q6), which the lanes evaluate in closed form (eub_right_chunks)."""
    import pyoracle
    from index4j_b200.builder import build_index
    texts = ["ab\ncde\n\nfghij\nk", "\n\nxy\nabcdefghijklmnopqrstuvw\nz\n", "no boundary at all in this one", "a\nbb\nccc\ndddd\neeeee\nffffff\nggggggg\nhhhhhhhh"]
    for t in texts:
        text = np.frombuffer(t.encode("utf-16-le"), dtype=np.uint16)
        blob = build_index(text, sr)
        f, o = flatcheck.FlatIndexHost(blob), pyoracle.OracleFmIndex(blob)
        n = text.size
        frm = np.arange(-1, n + 2, dtype=np.int32)
        for boundary in (10, ord("z") + 1):  # '\n', and a char that is not in the alphabet ("Boundary does not exist")
            for mode in (0, 1, 2):
                for dst_len in list(range(0, 27)) + [64]:
                    for offset in range(0, 5):
                        a1, l1, s1 = o.extract_until_boundary_batch(frm, boundary, dst_len, mode, threads=1, offset=offset)
                        a2, l2, s2 = f.eub(frm, boundary, dst_len, mode, offset=offset)
                        assert np.array_equal(s1, s2), (t, mode, dst_len, offset, frm[s1 != s2][:5], s1[s1 != s2][:5], s2[s1 != s2][:5])
                        ok = (s1 == 0) | (s1 == 8)
                        assert np.array_equal(l1[ok], l2[ok]), (t, mode, dst_len, offset, frm[ok][l1[ok] != l2[ok]][:5])
                        for i in np.flatnonzero(s1 == 0):
                            if frm[i] >= n:  # the terminator's own position: the reference returns a length that covers a
                                continue     # slot it never writes (end-of-text rule) — only status and length are defined
                            assert np.array_equal(a1[i, offset: offset + l1[i]], a2[i, offset: offset + l1[i]]), (t, mode, dst_len, offset, int(frm[i]))


@pytest.mark.parametrize("sr", [1, 3, 32])
def test_count_and_locate_exhaustive_on_small_texts(sr):
    """Every substring (up to 7 chars) of small texts, plus every 1- and 2-char string over the alphabet + an unknown char: count and
    the located position sets through the lane code, against the oracle and a naive scan."""
    import pyoracle
    from index4j_b200.builder import build_index
    texts = ["abracadabra abracadabra\nabra", "aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaab", "\n\n\n", "the quick brown fox\njumps over the lazy dog\nthe end"]
    for t in texts:
        text = np.frombuffer(t.encode("utf-16-le"), dtype=np.uint16)
        blob = build_index(text, sr)
        f, o = flatcheck.FlatIndexHost(blob), pyoracle.OracleFmIndex(blob)
        pats = {t[i: i + k] for i in range(len(t)) for k in range(1, 8) if i + k <= len(t)}
        alpha = sorted(set(t)) + ["中"]
        pats |= {a for a in alpha} | {a + b for a in alpha for b in alpha}
        pats = sorted(pats)
        arrs = [np.frombuffer(p.encode("utf-16-le"), dtype=np.uint16) for p in pats]
        off = np.zeros(len(arrs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([a.size for a in arrs])
        chars = np.concatenate(arrs)
        want, want_st = o.count_batch(chars, off, threads=1)
        got, got_st, ranges = f.count_batch(chars, off)
        assert np.array_equal(got_st, want_st) and np.array_equal(got, want)
        for i, p in enumerate(pats):
            naive = sum(1 for j in range(len(t) - len(p) + 1) if t[j: j + len(p)] == p)
            assert got[i] == naive, (t, p)
            rows = np.arange(int(ranges[i, 0]), int(ranges[i, 0]) + int(got[i]), dtype=np.uint32)
            if rows.size:
                pos = f.locate_rows(rows)
                assert sorted(pos.tolist()) == [j for j in range(len(t) - len(p) + 1) if t[j: j + len(p)] == p], (t, p)
