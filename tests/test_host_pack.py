"""Host half of the packed transport of fmgpu_count_batch (index4j_b200/csrc/host_pack.hpp): the pack pool and the narrowing
loop, replayed through tests/support/libflatcheck.so (no GPU): bytes, the per-chunk "does it fit a byte" verdict, odd
alignments and sizes, concurrent submitters."""
import threading

import numpy as np
import pytest

import flatcheck


def test_pool_exists():
    assert flatcheck.lib().fc_pack_threads() >= 1  # tests/conftest.py sets FMGPU_PACK_THREADS


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 65537, 1 << 20])
@pytest.mark.parametrize("groups,parts", [(1, 1), (3, 16), (8, 16)])
def test_narrow_matches_numpy(n, groups, parts):
    rng = np.random.default_rng(n + groups)
    chars = rng.integers(0, 256, n).astype(np.uint16)
    for shift in (0, 1, 3):  # source / destination alignments
        buf = np.zeros(n + shift, dtype=np.uint16)
        view = buf[shift:]
        view[:] = chars
        got, wide = flatcheck.pack_narrow(view, groups, parts)
        assert np.array_equal(got, chars.astype(np.uint8))
        assert (wide <= 0xFF).all()


def test_wide_chars_are_reported_per_chunk():
    n, groups = 100_000, 8
    chars = np.random.default_rng(1).integers(32, 127, n).astype(np.uint16)
    for pos, val in ((0, 0x0100), (n // 2 + 7, 0x4E2D), (n - 1, 0xFFFF)):
        c = chars.copy()
        c[pos] = val
        got, wide = flatcheck.pack_narrow(c, groups, 16)
        g = min(groups - 1, next(k for k in range(groups) if n * k // groups <= pos < n * (k + 1) // groups))
        assert wide[g] > 0xFF and all(wide[k] <= 0xFF for k in range(groups) if k != g)
        ok = np.ones(n, dtype=bool)
        ok[n * g // groups: n * (g + 1) // groups] = False  # the wide chunk's bytes are not used
        assert np.array_equal(got[ok], c.astype(np.uint8)[ok])


def test_concurrent_submitters():
    """Several callers at once (concurrent batch calls on one handle, or the worker threads of a multi-device handle): jobs queue
    up in the pool and every submitter also works on its own job."""
    n = 300_000
    inputs = [np.random.default_rng(t).integers(0, 256, n).astype(np.uint16) for t in range(6)]
    results = [None] * len(inputs)

    def run(t):
        for _ in range(5):
            results[t] = flatcheck.pack_narrow(inputs[t], 8, 16)[0]

    threads = [threading.Thread(target=run, args=(t,)) for t in range(len(inputs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for t, c in enumerate(inputs):
        assert np.array_equal(results[t], c.astype(np.uint8))
