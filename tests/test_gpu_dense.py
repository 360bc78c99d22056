"""Device-side denser SA sampling for locate (include/fmgpu.h: fmgpu_opts.locate_sample_rate, fmgpu_set_locate_dense):
the located positions, their order and the statuses are those of the reference's walk to the index's own samples
(fm/FmIndex.java:526-548) — checked against the oracle, and against the same handle with the dense samples switched off."""
import numpy as np
import pytest

from conftest import get_case, make_patterns

pytestmark = pytest.mark.gpu


def _locate_all(ix, chars, off, max_hits):
    n, o, p, s = ix.locate_batch(chars, off, max_hits)
    return n, o, p, s


@pytest.mark.parametrize("name,want,rate", [
    ("log1m_sr32", 0, 8),      # default request: 8
    ("log1m_sr32", 4, 4),
    ("log1m_sr32", 2, 2),
    ("log1m_sr32", 1, 1),      # every row carries its position
    ("log1m_sr32", 12, 8),     # largest divisor of 32 <= 12
    ("log1m_sr32", 32, 16),    # never the sampleRate itself
    ("log3m_sr16", 0, 8),
    ("log300k_sr64", 0, 8),
    ("tiny600k_sr4", 0, 2),
    ("nul1m_sr32", 0, 8),
    ("noextract", 0, 8),       # no inverse-SA samples needed: the walk starts from the SA samples
])
def test_dense_samples_do_not_change_locate(name, want, rate):
    from index4j_b200 import FmIndex
    case = get_case(name)
    ix = FmIndex.read(case.blob, locate_sample_rate=want)
    try:
        assert ix.locate_sample_rate == rate
        assert ix.dense_sample_bytes() > 0
        chars, off = make_patterns(case.text, 1200, 1, 24, seed=77)
        for max_hits in (0, 7, 300):
            n1, o1, p1, s1 = _locate_all(ix, chars, off, max_hits)
            ix.set_locate_dense(False)
            n0, o0, p0, s0 = _locate_all(ix, chars, off, max_hits)
            ix.set_locate_dense(True)
            assert np.array_equal(n1, n0) and np.array_equal(o1, o0) and np.array_equal(s1, s0)
            assert np.array_equal(p1, p0)
        counts, _ = case.oracle.count_batch(chars, off, threads=4)
        stride = int(max(1, min(300, counts.max())))
        want_n, want_pos, want_st = case.oracle.locate_batch(chars, off, 300, stride, threads=4)
        assert np.array_equal(n1, want_n) and np.array_equal(s1, want_st)
        for i in range(want_n.size):
            assert np.array_equal(p1[int(o1[i]): int(o1[i + 1])], want_pos[i, : want_n[i]])
        # fewer LF steps per hit: the walks end at the dense samples
        ix.set_stats(True)
        _locate_all(ix, chars, off, 300)
        dense_steps = ix.last_stats()["lf_steps"]
        ix.set_locate_dense(False)
        _locate_all(ix, chars, off, 300)
        own_steps = ix.last_stats()["lf_steps"]
        ix.set_stats(False)
        assert dense_steps < own_steps
    finally:
        ix.close()


def test_every_text_position_is_located():
    """Single chars as patterns: together they locate EVERY text position — every row's walk, incl. the rows behind the last
    multiple of sampleRate (covered by the walk from the sentinel's row) and position 0."""
    from index4j_b200 import FmIndex
    from index4j_b200.builder import build_index, gen_log_text
    for n, sr, want in ((5003, 32, 8), (4096, 16, 4), (1000, 32, 0), (33, 32, 8), (7, 4, 2)):
        text = gen_log_text(n, seed=31 + n)
        ix = FmIndex.read(build_index(text, sr), locate_sample_rate=want)
        try:
            assert ix.dense_sample_bytes() > 0
            syms = np.unique(text)
            chars = syms.astype(np.uint16)
            off = np.arange(syms.size + 1, dtype=np.uint64)
            n_hits, hit_off, pos, st = ix.locate_batch(chars, off, 0)
            assert not st.any() and int(n_hits.sum()) == n
            assert np.array_equal(np.sort(pos), np.arange(n))
            for i, c in enumerate(syms):
                assert np.array_equal(np.sort(pos[int(hit_off[i]): int(hit_off[i + 1])]), np.flatnonzero(text == c))
        finally:
            ix.close()


@pytest.mark.parametrize("name", ["q4_2m_sr32", "q1_runs_sr4", "multi400k_sr8", "log200k_sr1"])
def test_no_dense_samples_where_the_reference_walk_can_differ(name):
    """Quirk Q4 (length % 2^20 == 0: LF steps can throw) and Q1 (> 256 symbols: walks can cycle): the reference's full walk is
    what runs; sampleRate 1 has nothing to densify."""
    from index4j_b200 import FmIndex
    ix = FmIndex.read(get_case(name).blob, locate_sample_rate=2)
    try:
        assert ix.dense_sample_bytes() == 0
        assert ix.locate_sample_rate == ix.sample_rate
    finally:
        ix.close()


def test_opt_out():
    from index4j_b200 import FmIndex
    ix = FmIndex.read(get_case("log1m_sr32").blob, locate_sample_rate=-1)
    try:
        assert ix.dense_sample_bytes() == 0 and ix.locate_sample_rate == 32
    finally:
        ix.close()


@pytest.mark.parametrize("name", ["log1m_sr32", "q4_2m_sr32", "tiny600k_sr4"])
def test_host_locate_in_chunks(name, monkeypatch):
    """fmgpu_locate_batch walks the rows in chunks so that the download of chunk k overlaps the walks of chunk k + 1: same
    positions and statuses as the oracle when a small batch is forced into several chunks (incl. the per-pattern status of
    walks on which the reference throws, quirk Q4, which is looked up by the hit's row in the whole call)."""
    from index4j_b200 import FmIndex
    case = get_case(name)
    ix = FmIndex.read(case.blob)
    try:
        chars, off = make_patterns(case.text, 1500, 1, 20, seed=5)
        counts, _ = case.oracle.count_batch(chars, off, threads=4)
        stride = int(max(1, min(200, counts.max())))
        want_n, want_pos, want_st = case.oracle.locate_batch(chars, off, 200, stride, threads=4)
        monkeypatch.setenv("FMGPU_LOCATE_CHUNK_ROWS", "997")
        n, o, p, s = ix.locate_batch(chars, off, 200)
        monkeypatch.delenv("FMGPU_LOCATE_CHUNK_ROWS")
        n1, o1, p1, s1 = ix.locate_batch(chars, off, 200)
        ok = want_st == 0  # where the reference throws, only the status is defined
        assert np.array_equal(s, want_st) and np.array_equal(n[ok], want_n[ok])
        if name == "q4_2m_sr32":
            assert (want_st == 9).any()
        assert np.array_equal(n1, n) and np.array_equal(o1, o) and np.array_equal(p1, p) and np.array_equal(s1, s)
        for i in range(want_n.size):
            if want_st[i] == 0:
                assert np.array_equal(p[int(o[i]): int(o[i + 1])], want_pos[i, : want_n[i]])
    finally:
        ix.close()
