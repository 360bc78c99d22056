"""GPU parity inside the reference's quirk regions (SURVEY.md section 8, Q1 / Q4 / text containing \\0 / maxMatches == 0) and at the
size of BASELINE.json configs[0] — every call through the C ABI, every result against the CPU oracle (bit-exact).

  q4_2m_sr32   n + 1 = 2 * 2^20: WaveletFixedBlockBoosting.rank(size, .) indexes past the superblock arrays and throws
               ArrayIndexOutOfBounds (wavelet/WaveletFixedBlockBoosting.java:1022-1026) -> status 9 from count, from locate (in the
               backward search AND inside the LF walk of a hit) and from the wavelet rank entry point
  nul1m_sr32   a log text with 1000 chars overwritten by \\0 (FmIndexTest.java:53-65, 202-217): the text's own \\0 is an ordinary
               symbol, the sentinel keeps code 0
  q1_runs_sr4  308 symbols, BWT runs of symbols with alphabet code >= 256: inverseSelect on a single-symbol block keeps only the
               LOW BYTE of the symbol (:1329-1332), so LF walks through such blocks continue from the wrong row — in Java and here
"""
import numpy as np
import pytest

from conftest import QUIRK_CASE_NAMES, get_case, make_patterns

import pyoracle

pytestmark = pytest.mark.gpu

QUIRK_IDS = {"q4_2m_sr32": "Q4-size-multiple-of-2^20", "nul1m_sr32": "NUL-in-text", "q1_runs_sr4": "Q1-run-symbol-ge-256"}
quirk_cases = pytest.mark.parametrize("name", QUIRK_CASE_NAMES, ids=[QUIRK_IDS[n] for n in QUIRK_CASE_NAMES])


def _patterns(case, n_pat, seed):
    chars, off = make_patterns(case.text, n_pat, 1, 40, seed=seed)
    if case.name == "nul1m_sr32":  # make sure patterns that contain \0 are in the batch: substrings around the text's NULs
        t = case.text
        at = np.flatnonzero(t == 0)[:400]
        extra = [t[max(0, int(a) - 3): int(a) + 4] for a in at] + [np.array([0], dtype=np.uint16), np.array([0, 0], dtype=np.uint16)]
        eo = np.cumsum([p.size for p in extra]).astype(np.uint64)
        chars = np.concatenate([chars] + extra).astype(np.uint16)
        off = np.concatenate([off, off[-1] + eo]).astype(np.uint64)
    return chars, off


@quirk_cases
def test_quirk_count(gpu_indexes, name):
    case, g = get_case(name), gpu_indexes(name)
    assert g.getInputLength() == case.oracle.getInputLength() == case.text.size + 1
    assert g.getAlphabetLength() == case.oracle.getAlphabetLength()
    chars, off = _patterns(case, 6000, 5)
    want, want_st = case.oracle.count_batch(chars, off, threads=4)
    for table in (True, False):  # with and without the q-gram start table
        g.set_start_table(table)
        try:
            got, got_st = g.count_batch(chars, off, return_status=True)
        finally:
            g.set_start_table(True)
        assert np.array_equal(got_st, want_st) and np.array_equal(got, want)
    if name == "q4_2m_sr32":
        assert (want_st == 9).sum() > 0  # the quirk fires: rank(size, c) is reached when end == C[c + 1] == length
    if name == "nul1m_sr32":
        nul_pats = [i for i in range(off.size - 1) if (chars[int(off[i]): int(off[i + 1])] == 0).any()]
        assert len(nul_pats) >= 400 and (want[nul_pats] > 0).sum() >= 400  # \0 is searchable like any other char
        assert g.count(np.array([0], dtype=np.uint16)) == int((case.text == 0).sum())


@quirk_cases
@pytest.mark.parametrize("max_hits", [-1, 0, 1, 100])
def test_quirk_locate(gpu_indexes, name, max_hits):
    case, g = get_case(name), gpu_indexes(name)
    chars, off = _patterns(case, 1500 if name != "q1_runs_sr4" or max_hits > 0 else 400, 6)
    counts, _ = case.oracle.count_batch(chars, off, threads=4)
    if name == "q1_runs_sr4" and max_hits <= 0:  # unlimited: keep the patterns with a moderate number of hits
        keep = np.flatnonzero(counts < 20_000)
        assert keep.size > 50
        pats = [chars[int(off[i]): int(off[i + 1])] for i in keep]
        off = np.concatenate([[0], np.cumsum([p.size for p in pats])]).astype(np.uint64)
        chars = np.concatenate(pats).astype(np.uint16)
        counts = counts[keep]
    stride = int(max(1, counts.max() if max_hits <= 0 else min(max_hits, max(counts.max(), 1))))
    want_n, want_pos, want_st = case.oracle.locate_batch(chars, off, max_hits, stride, threads=4)
    n_hits, hit_off, pos, st = g.locate_batch(chars, off, max_hits)
    assert np.array_equal(st, want_st)
    ok = want_st == 0
    assert np.array_equal(n_hits[ok], want_n[ok])
    for i in np.flatnonzero(ok):
        assert np.array_equal(pos[int(hit_off[i]): int(hit_off[i + 1])], want_pos[i, : want_n[i]]), i
    if max_hits == 0:  # maxMatches <= 0 means unlimited (fm/FmIndex.java:544): every occurrence is reported
        assert np.array_equal(n_hits[ok], counts[ok])
    if name == "q4_2m_sr32" and max_hits <= 0:
        # more patterns throw in locate than in count: the extra ones throw inside the LF walk of one of their hits
        _, cst = case.oracle.count_batch(chars, off, threads=4)
        assert (want_st == 9).sum() > (cst == 9).sum()
    if name == "nul1m_sr32":
        for i in np.flatnonzero(ok)[:200]:
            p = chars[int(off[i]): int(off[i + 1])]
            if max_hits <= 0:
                assert np.array_equal(np.sort(pos[int(hit_off[i]): int(hit_off[i + 1])]), pyoracle.naive_locations(case.text, p))


@quirk_cases
def test_quirk_extract(gpu_indexes, name):
    case, g = get_case(name), gpu_indexes(name)
    rng = np.random.default_rng(8)
    n = case.text.size
    m = 2500
    start = rng.integers(0, n - 200, m).astype(np.int32)
    stop = (start + rng.integers(0, 160, m)).astype(np.int32)
    start[:4] = [0, 0, n - 50, n - 1]
    stop[:4] = [n, 1, n, n]
    arena, aoff, got_len, st = g.extract_batch(start, stop)
    w_arena, w_len, w_st = case.oracle.extract_batch(start, stop, int((stop - start).max()), threads=4)
    assert np.array_equal(st, w_st) and np.array_equal(got_len[w_st == 0], w_len[w_st == 0])
    differs_from_text = 0
    for i in np.flatnonzero(w_st == 0):
        a = arena[int(aoff[i]): int(aoff[i + 1])]
        assert np.array_equal(a, w_arena[i, : w_len[i]]), i
        differs_from_text += int(not np.array_equal(a, case.text[start[i]: stop[i]]))
    if name == "q1_runs_sr4":
        assert differs_from_text > 0  # the reference itself returns other chars than the text here (truncated run symbols)
    else:
        assert differs_from_text == 0


@quirk_cases
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_quirk_extract_until_boundary(gpu_indexes, name, mode):
    case, g = get_case(name), gpu_indexes(name)
    n = case.text.size
    rng = np.random.default_rng(9 + mode)
    boundaries = [10, 0] if name == "nul1m_sr32" else [10]
    for boundary in boundaries:
        for dst_len in (512, 37, 4):
            frm = np.concatenate([rng.integers(0, n, 1000), np.arange(n - 12, n + 2), np.arange(-1, 6)]).astype(np.int32)
            w_arena, w_len, w_st = case.oracle.extract_until_boundary_batch(frm, boundary, dst_len, mode, threads=4)
            arena, got_len, st = g.extract_until_boundary_batch(frm, boundary, dst_len, mode)
            assert np.array_equal(st, w_st), (boundary, dst_len)
            ok = (w_st == 0) | (w_st == 8)
            assert np.array_equal(got_len[ok], w_len[ok])
            for i in np.flatnonzero(w_st == 0):
                if frm[i] >= n:
                    continue
                assert np.array_equal(arena[i, : w_len[i]], w_arena[i, : w_len[i]]), (boundary, dst_len, i)


@quirk_cases
def test_quirk_wavelet_entry_points(gpu_indexes, name):
    case, g = get_case(name), gpu_indexes(name)
    rng = np.random.default_rng(10)
    L = case.oracle.getInputLength()
    sigma = case.oracle.getAlphabetLength() + 2
    pos = np.concatenate([rng.integers(0, L + 1, 8000), [0, 1, L - 1, L, L, L, L + 5, -1], np.full(200, L)]).astype(np.int64)
    sym = rng.integers(0, sigma, pos.size).astype(np.int32)
    got, st = g.wavelet_rank_batch(pos, sym)
    n_throw = 0
    for i in range(pos.size):
        try:
            want, wst = case.oracle.wfbb_rank(int(pos[i]), int(sym[i])), 0
        except pyoracle.JavaException as e:
            want, wst = 0, e.status
        n_throw += int(wst == 9)
        assert st[i] == wst and (wst or got[i] == want), (int(pos[i]), int(sym[i]))
    if name == "q4_2m_sr32":
        assert n_throw >= 150  # rank(size, sym) throws for every symbol of the alphabet
    ipos = np.concatenate([rng.integers(0, L, 12000), [0, 1, L - 1]]).astype(np.int64)
    got, st = g.wavelet_inverse_select_batch(ipos)
    for i in range(ipos.size):
        p = int(ipos[i])
        assert st[i] == 0 and int(got[i]) == case.oracle.wfbb_inverse_select(p), p
    if name == "q1_runs_sr4":
        # the truncated value itself: the late chars have alphabet codes 301..306 (codes go by first appearance,
        # fm/FmIndex.java:396-435); inside single-symbol blocks inverseSelect answers with their LOW BYTE (301 & 0xff = 45, ...)
        code_of = {}
        for ch in case.text:
            if int(ch) not in code_of:
                code_of[int(ch)] = len(code_of) + 1
        late = sorted(code_of[c] for c in range(0x5000, 0x5006) if c in code_of)
        assert len(late) >= 2 and min(late) >= 301
        alias = {c & 0xFF for c in late}
        codes = np.array([code_of[int(c)] for c in case.text])
        real = int(np.isin(codes, list(alias)).sum())  # occurrences of the symbols that really have codes 45..50
        answered = int(np.isin((got & 0xFFFF).astype(np.int64), list(alias)).sum())
        assert real < 100 and answered > 1000, (real, answered)


def test_cfg1_16mib_index(gpu_indexes):
    """BASELINE.json configs[0] at full size: FmIndex(sampleRate 32, extraction on) over 16 MiB of synthetic log text, 10,000
    substring patterns of length 8-32: count, locate (max 1000) and extractUntilBoundary of located hits against the oracle."""
    from index4j_b200.builder import gen_patterns
    case, g = get_case("cfg1_16m_sr32"), gpu_indexes("cfg1_16m_sr32")
    chars, off = gen_patterns(case.text, 10_000, 8, 32, 42)
    want, want_st = case.oracle.count_batch(chars, off, threads=8)
    got, got_st = g.count_batch(chars, off, return_status=True)
    assert np.array_equal(got, want) and np.array_equal(got_st, want_st) and (want > 0).all()
    k = 2000
    w_n, w_pos, w_st = case.oracle.locate_batch(chars[: int(off[k])], off[: k + 1], 1000, 1000, threads=8)
    n_hits, hit_off, pos, st = g.locate_batch(chars[: int(off[k])], off[: k + 1], 1000)
    assert np.array_equal(n_hits, w_n) and np.array_equal(st, w_st)
    for i in range(k):
        assert np.array_equal(pos[int(hit_off[i]): int(hit_off[i + 1])], w_pos[i, : w_n[i]]), i
    frm = pos[:: max(1, pos.size // 3000)].astype(np.int32)
    w_arena, w_len, w_st = case.oracle.extract_until_boundary_batch(frm, 10, 512, 0, threads=8)
    arena, got_len, st = g.extract_until_boundary_batch(frm, "\n", 512, 0)
    assert np.array_equal(st, w_st) and np.array_equal(got_len, w_len)
    for i in np.flatnonzero(w_st == 0):
        assert np.array_equal(arena[i, : w_len[i]], w_arena[i, : w_len[i]]), i
