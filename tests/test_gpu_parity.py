"""GPU parity: the CUDA engine, called through the C ABI, against the CPU oracle (bit-exact).

Mirrors the reference's FmIndexTest matrix (indices/src/test/java/com/dynatrace/fm/FmIndexTest.java):
count :43-129, locate :181-282, extract :350-374, extractUntilBoundary* :376-562, error contract
:284-348,:402-475 — on synthetic texts (the reference's log fixture cannot travel to the GPU box).
"""
import numpy as np
import pytest

from conftest import CASE_NAMES, get_case, make_patterns

import pyoracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASE_NAMES)
def test_count_matches_oracle(gpu_indexes, name):
    case = get_case(name)
    g = gpu_indexes(name)
    assert g.getInputLength() == case.oracle.getInputLength() == case.text.size + 1
    assert g.getAlphabetLength() == case.oracle.getAlphabetLength()
    chars, off = make_patterns(case.text, 6000, 1, 48, seed=11)
    want, want_st = case.oracle.count_batch(chars, off, threads=4)
    got, got_st = g.count_batch(chars, off, return_status=True)  # production kernel (no work counters)
    assert np.array_equal(got_st, want_st)
    assert np.array_equal(got, want)
    assert int((want > 0).sum()) > 1000
    # the same without the q-gram start table (every pattern from its last char), with the instrumented kernel: identical
    # results, and then the kernel performs exactly the rank queries of the reference's loop (their code lengths summed)
    g.set_stats(True)
    g.set_start_table(False)
    try:
        got, got_st = g.count_batch(chars, off, return_status=True)
        mine = g.last_stats()
    finally:
        g.set_stats(False)
        g.set_start_table(True)
    assert np.array_equal(got_st, want_st) and np.array_equal(got, want)
    case.oracle.stats(reset=True)
    case.oracle.count_batch(chars, off, threads=1)
    st = case.oracle.stats()
    assert mine["launches"] == 4
    assert mine["rank_levels"] == st["rank_levels"]  # work counters agree with the oracle's instrumentation
    assert mine["level_records"] <= mine["ranks"]     # at most one occurrence record per rank


def test_start_table_is_transparent(gpu_indexes):
    """Patterns of every length around q, known / unknown chars at every position of the last q chars: the q-gram start table gives
    the counts, statuses and located positions of the step-by-step search."""
    for name in ("log1m_sr32", "tiny600k_sr4"):
        case, g = get_case(name), gpu_indexes(name)
        q = g.start_table_q()
        assert q >= 2
        rng = np.random.default_rng(9)
        t = case.text
        pats = []
        for ln in range(1, q + 4):
            for _ in range(300):
                a = int(rng.integers(0, t.size - ln))
                p = t[a: a + ln].copy()
                if rng.random() < 0.3:
                    p[int(rng.integers(0, ln))] = rng.choice([0xFFFE, 0, int(t[int(rng.integers(0, t.size))])])
                pats.append(p)
        off = np.zeros(len(pats) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([p.size for p in pats])
        chars = np.concatenate(pats).astype(np.uint16)
        want, want_st = case.oracle.count_batch(chars, off, threads=4)
        got, got_st = g.count_batch(chars, off, return_status=True)
        assert np.array_equal(got_st, want_st) and np.array_equal(got, want)
        n1, o1, p1, s1 = g.locate_batch(chars, off, 20)
        g.set_start_table(False)
        try:
            got0 = g.count_batch(chars, off)
            n0, o0, p0, s0 = g.locate_batch(chars, off, 20)
        finally:
            g.set_start_table(True)
        assert np.array_equal(got0, want) and np.array_equal(n0, n1) and np.array_equal(p0, p1) and np.array_equal(s0, s1)


def test_count_edge_cases(gpu_indexes):
    case = get_case("log1m_sr32")
    g = gpu_indexes("log1m_sr32")
    t = case.text
    pats = [t[0:1], t[-1:], t[:64], t[-64:], t[100:100], np.array([0xFFFE], dtype=np.uint16), t[5:300], np.array([0], dtype=np.uint16)]
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in pats])
    chars = np.concatenate(pats).astype(np.uint16)
    want, want_st = case.oracle.count_batch(chars, off)
    got, got_st = g.count_batch(chars, off, return_status=True)
    assert np.array_equal(got_st, want_st) and np.array_equal(got, want)
    assert got_st[4] == 9  # empty pattern: pattern[-1] in the reference
    # empty batch
    c = g.count_batch(np.zeros(0, dtype=np.uint16), np.zeros(1, dtype=np.uint64))
    assert c.size == 0
    # single-query API
    assert g.count("INFO") == case.oracle.count("INFO")
    assert g.count(t[1000:1040], 3, 20) == case.oracle.count(t[1000:1040], 3, 20)


@pytest.mark.parametrize("name", CASE_NAMES)
@pytest.mark.parametrize("max_hits", [-1, 0, 1, 100])
def test_locate_matches_oracle(gpu_indexes, name, max_hits):
    case = get_case(name)
    g = gpu_indexes(name)
    chars, off = make_patterns(case.text, 1500, 2, 32, seed=21)
    counts, _ = case.oracle.count_batch(chars, off, threads=4)
    stride = int(max(1, counts.max() if max_hits <= 0 else min(max_hits, max(counts.max(), 1))))
    want_n, want_pos, want_st = case.oracle.locate_batch(chars, off, max_hits, stride, threads=4)
    n_hits, hit_off, pos, st = g.locate_batch(chars, off, max_hits)
    assert np.array_equal(st, want_st)
    assert np.array_equal(n_hits, want_n)
    if max_hits == 0:  # maxMatches <= 0 = unlimited (fm/FmIndex.java:544)
        assert np.array_equal(n_hits[st == 0], counts[st == 0])
    assert int(hit_off[-1]) == int(want_n.sum()) == pos.size
    for i in range(want_n.size):
        a = pos[int(hit_off[i]): int(hit_off[i + 1])]
        assert np.array_equal(a, want_pos[i, : want_n[i]]), (i, a[:5], want_pos[i, :5])  # same order as Java fills its array
    # the instrumented kernels give the same answer and count their LF steps
    g.set_stats(True)
    try:
        n2, off2, pos2, st2 = g.locate_batch(chars, off, max_hits)
        assert g.last_stats()["lf_steps"] > 0 or case.sample_rate == 1
    finally:
        g.set_stats(False)
    assert np.array_equal(n2, n_hits) and np.array_equal(pos2, pos) and np.array_equal(st2, st)


def test_locate_single_query_api(gpu_indexes):
    case = get_case("log1m_sr32")
    g = gpu_indexes("log1m_sr32")
    loc = np.zeros(100, dtype=np.int32)
    assert g.locate("INFO", 0, 4, loc, 100) == 100  # FmIndexTest.java:195-200
    want = case.oracle.locate("INFO", 0, 4, 100, 100)
    assert np.array_equal(loc, want)
    allpos = g.locate("WARN dfs.")
    import pyoracle
    assert np.array_equal(np.sort(allpos), pyoracle.naive_locations(case.text, "WARN dfs."))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_extract_matches_text(gpu_indexes, name):
    case = get_case(name)
    g = gpu_indexes(name)
    rng = np.random.default_rng(31)
    n = case.text.size
    m = 3000
    start = rng.integers(0, n - 200, m).astype(np.int32)
    ln = rng.integers(0, 160, m).astype(np.int32)
    stop = start + ln
    start[:4] = [0, 0, n - 50, n - 1]
    stop[:4] = [n, 1, n, n]
    arena, aoff, got_len, st = g.extract_batch(start, stop)
    assert not st.any()
    assert np.array_equal(got_len, stop - start)
    for i in range(m):
        assert np.array_equal(arena[int(aoff[i]): int(aoff[i + 1])], case.text[start[i]: stop[i]]), i
    # against the oracle too, including error statuses (FmIndexTest.java:304-348)
    start2 = np.array([-1, 5, 5, 10, 0], dtype=np.int32)
    stop2 = np.array([5, n + 1, 20, 5, n], dtype=np.int32)
    aoff2 = np.array([0, 10, 20, 25, 30, 30 + n], dtype=np.uint64)  # third slot too small
    _, _, l2, s2 = g.extract_batch(start2, stop2, aoff2)
    for i in range(5):
        room = int(aoff2[i + 1] - aoff2[i])
        try:
            want = case.oracle.extract(int(start2[i]), int(stop2[i]), room)
            assert s2[i] == 0 and l2[i] == stop2[i] - start2[i], i
        except Exception as e:
            assert s2[i] == e.status, (i, s2[i], e.status)


@pytest.mark.parametrize("name", CASE_NAMES)
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("dst_len", [512, 37, 10, 2])
def test_extract_until_boundary_matches_oracle(gpu_indexes, name, mode, dst_len):
    case = get_case(name)
    g = gpu_indexes(name)
    n = case.text.size
    rng = np.random.default_rng(41 + mode)
    frm = np.concatenate([rng.integers(0, n, 1200), np.arange(n - 12, n + 2), np.arange(-1, 6)]).astype(np.int32)
    want_arena, want_len, want_st = case.oracle.extract_until_boundary_batch(frm, 10, dst_len, mode, threads=4)
    arena, got_len, st = g.extract_until_boundary_batch(frm, "\n", dst_len, mode)
    assert np.array_equal(st, want_st)
    ok = (want_st == 0) | (want_st == 8)
    assert np.array_equal(got_len[ok], want_len[ok])
    for i in np.flatnonzero(want_st == 0):
        if frm[i] >= n:  # the sentinel position: the reference returns one char it never wrote
            continue
        k = want_len[i]
        assert np.array_equal(arena[i, :k], want_arena[i, :k]), (i, frm[i], k)


@pytest.mark.parametrize("name", ["log300k_sr64", "tiny600k_sr4", "multi400k_sr8"])
@pytest.mark.parametrize("offset", [1, 30, 39, 40, 64])
def test_offset_argument_matches_oracle(gpu_indexes, name, offset):
    """The reference's `offset` argument (fm/FmIndex.java:564, :640, :772, :844): where the chars land, the capacity tests, the
    N of "Currently extracted: N", and System.arraycopy throwing when the left part does not fit behind the offset."""
    case, g = get_case(name), gpu_indexes(name)
    n = case.text.size
    rng = np.random.default_rng(80 + offset)
    for mode in (0, 1, 2):
        for dst_len in (40, 512):
            frm = np.concatenate([rng.integers(0, n, 800), np.arange(n - 9, n + 1), np.arange(0, 4)]).astype(np.int32)
            w_arena, w_len, w_st = case.oracle.extract_until_boundary_batch(frm, 10, dst_len, mode, threads=4, offset=offset)
            arena, got_len, st = g.extract_until_boundary_batch(frm, "\n", dst_len, mode, offset=offset)
            assert np.array_equal(st, w_st), (mode, dst_len)
            ok = (w_st == 0) | (w_st == 8)
            assert np.array_equal(got_len[ok], w_len[ok])
            for i in np.flatnonzero(w_st == 0):
                if frm[i] >= n:
                    continue
                assert np.array_equal(arena[i, offset: offset + w_len[i]], w_arena[i, offset: offset + w_len[i]]), (mode, dst_len, i)
    m = 1000
    start = rng.integers(0, n - 100, m).astype(np.int32)
    stop = (start + rng.integers(0, 80, m)).astype(np.int32)
    stride = 90
    aoff = (np.arange(m + 1) * stride).astype(np.uint64)
    arena, _, got_len, st = g.extract_batch(start, stop, aoff, offset=offset)
    w_arena, w_len, w_st = case.oracle.extract_batch(start, stop, stride, threads=4, offset=offset)
    assert np.array_equal(st, w_st) and ((w_st == 5).sum() > 0 or offset < 20) and (w_st == 0).sum() > 0
    for i in np.flatnonzero(w_st == 0):
        assert got_len[i] == w_len[i]
        assert np.array_equal(arena[i * stride + offset: i * stride + offset + w_len[i]], w_arena[i, offset: offset + w_len[i]]), i
    # the single-query forms with a destination array and an offset, like the Java calls
    import pyoracle
    from index4j_b200.fm_index import FmIndexError
    for f in frm[:40]:
        dst = np.zeros(600, dtype=np.uint16)
        try:
            want = case.oracle.extract_until_boundary(int(f), 600, 10, 0, offset=offset)
        except pyoracle.JavaException as e:
            with pytest.raises(FmIndexError) as ei:
                g.extractUntilBoundary(int(f), dst, offset, "\n")
            assert ei.value.status == e.status and (e.status != 8 or ei.value.n == e.n)
            continue
        k = g.extractUntilBoundary(int(f), dst, offset, "\n")
        assert k == want.size and np.array_equal(dst[offset: offset + k], want)
    with pytest.raises(Exception):
        g.extractUntilBoundary(50, dst, -1, "\n")  # negative offset: rejected (the reference indexes out of bounds)


def test_extract_until_boundary_error_contract(gpu_indexes):
    from index4j_b200 import FmIndex
    from index4j_b200.fm_index import FmIndexError, FmIndexIllegalArgument
    case = get_case("log1m_sr32")
    g = gpu_indexes("log1m_sr32")
    with pytest.raises(FmIndexIllegalArgument, match="Boundary does not exist"):
        g.extractUntilBoundary(50, 100, 0, "一")
    with pytest.raises(FmIndexIllegalArgument, match="size zero"):
        g.extractUntilBoundary(50, 0, 0, "\n")
    with pytest.raises(FmIndexError, match="Requested position less than 0"):
        g.extractUntilBoundary(-1, 100, 0, "\n")
    with pytest.raises(FmIndexError, match="Requested position longer than index string"):
        g.extractUntilBoundary(case.text.size + 1, 100, 0, "\n")
    with pytest.raises(FmIndexError, match=r"Currently extracted: \d+"):
        g.extractUntilBoundary(50, 10, 0, "\n")
    # same N as the oracle for the three variants (FmIndexTest.java:430-475 pins 13 / 10 / 11 on its fixture)
    for mode, fn in ((0, g.extractUntilBoundary), (1, g.extractUntilBoundaryLeft), (2, g.extractUntilBoundaryRight)):
        try:
            fn(50, 10, 0, "\n")
            got = None
        except FmIndexError as e:
            got = (e.status, e.n)
        try:
            case.oracle.extract_until_boundary(50, 10, 10, mode)
            want = None
        except Exception as e:
            want = (e.status, e.n)
        assert got == want
    noex = FmIndex.read(get_case("noextract").blob)
    with pytest.raises(FmIndexError, match="Text recovery not enabled at build time"):
        noex.extract(0, 5)
    with pytest.raises(FmIndexError, match="Text recovery not enabled at build time"):
        noex.extractUntilBoundary(5, 100, 0, "\n")
    assert noex.count("INFO") == get_case("noextract").oracle.count("INFO")
    noex.close()


def test_device_resident_api(gpu_indexes):
    import torch
    case = get_case("log1m_sr32")
    g = gpu_indexes("log1m_sr32")
    chars, off = make_patterns(case.text, 4000, 4, 64, seed=51)
    dev = torch.device("cuda", g.device)
    d_chars = torch.from_numpy(chars.view(np.int16)).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    d_counts = torch.empty(off.size - 1, dtype=torch.int32, device=dev)
    d_status = torch.empty(off.size - 1, dtype=torch.int32, device=dev)
    g.count_batch_device(d_chars, d_off, d_counts, d_status)
    torch.cuda.synchronize()
    want, want_st = case.oracle.count_batch(chars, off, threads=4)
    assert np.array_equal(d_counts.cpu().numpy(), want)
    assert np.array_equal(d_status.cpu().numpy(), want_st)


def test_malformed_streams_rejected():
    from index4j_b200 import FmIndex
    blob = bytearray(get_case("log300k_sr64").blob)
    with pytest.raises(IOError):
        FmIndex.read(bytes(blob[: len(blob) // 2]))
    bad = bytearray(blob)
    # the first primitive after the 4-byte stream header and the 0x7A block header is the serial version
    assert bad[:4] == b"\xac\xed\x00\x05"
    bad[9] = 3
    with pytest.raises(IOError, match="Incompatible serial versions! Expected version 0 but was 3."):
        FmIndex.read(bytes(bad))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_wavelet_rank_and_inverse_select_match_oracle(gpu_indexes, name):
    """The wavelet structure queried directly (fmgpu_wavelet_rank_batch / fmgpu_wavelet_inverse_select_batch)."""
    case, g = get_case(name), gpu_indexes(name)
    rng = np.random.default_rng(17)
    L = case.oracle.getInputLength()
    sigma = case.oracle.getAlphabetLength() + 2
    pos = np.concatenate([rng.integers(0, L + 1, 20000), [0, 1, L - 1, L, L + 5, -1, -7],
                          (rng.integers(1, max(2, L >> 9), 3000) << 9) + rng.integers(-1, 2, 3000)]).astype(np.int64)
    sym = rng.integers(0, sigma, pos.size).astype(np.int32)
    sym[:50] = -1
    got, st = g.wavelet_rank_batch(pos, sym)
    for i in range(pos.size):
        try:
            want, wst = case.oracle.wfbb_rank(int(pos[i]), int(sym[i])), 0
        except pyoracle.JavaException as e:
            want, wst = 0, e.status
        assert st[i] == wst and (wst or got[i] == want), (int(pos[i]), int(sym[i]), want, int(got[i]), wst, int(st[i]))
    ipos = np.concatenate([rng.integers(0, L, 20000), [0, 1, L - 1, -1, L, L + 9]]).astype(np.int64)
    got, st = g.wavelet_inverse_select_batch(ipos)
    for i in range(ipos.size):
        p = int(ipos[i])
        if p < 0 or p >= L:
            assert st[i] == 9
        else:
            assert st[i] == 0 and int(got[i]) == case.oracle.wfbb_inverse_select(p), p
