"""One handle = the reference's one @ThreadSafe FmIndex (fm/FmIndex.java:82): concurrent callers, *_device calls on different
streams, and an index replicated on several GPUs whose host-pointer calls are cut into one slice per device.  Every result is
compared with the single-caller / single-device result, which tests/test_gpu_parity.py pins against the oracle."""
import threading

import numpy as np
import pytest

from conftest import get_case, make_patterns

pytestmark = pytest.mark.gpu


def _ref_results(g, chars, off, frm):
    counts, st = g.count_batch(chars, off, return_status=True)
    n_hits, hit_off, pos, lst = g.locate_batch(chars, off, 50)
    arena, ln, est = g.extract_until_boundary_batch(frm, "\n", 200)
    return counts, st, n_hits, hit_off, pos, lst, arena, ln, est


def _same(a, b):
    counts, st, n_hits, hit_off, pos, lst, arena, ln, est = a
    c2, s2, n2, o2, p2, l2, a2, ln2, e2 = b
    assert np.array_equal(counts, c2) and np.array_equal(st, s2)
    assert np.array_equal(n_hits, n2) and np.array_equal(hit_off, o2) and np.array_equal(pos, p2) and np.array_equal(lst, l2)
    assert np.array_equal(ln, ln2) and np.array_equal(est, e2)
    for i in np.flatnonzero(est == 0):
        assert np.array_equal(arena[i, : ln[i]], a2[i, : ln[i]])


def test_concurrent_callers_of_one_handle(gpu_indexes):
    """Eight host threads hammer one handle with count / locate / extractUntilBoundary batches (ctypes drops the GIL inside the
    calls): every call returns what it returns when it runs alone."""
    case, g = get_case("log1m_sr32"), gpu_indexes("log1m_sr32")
    rng = np.random.default_rng(3)
    work = []
    for t in range(8):
        chars, off = make_patterns(case.text, 3000 + 500 * t, 2, 48, seed=100 + t)
        frm = rng.integers(0, case.text.size, 1500 + 100 * t).astype(np.int32)
        work.append((chars, off, frm, _ref_results(g, chars, off, frm)))
    errors = []

    def run(t):
        try:
            chars, off, frm, want = work[t]
            for _ in range(6):
                _same(want, _ref_results(g, chars, off, frm))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_concurrent_callers_with_the_packed_transport(gpu_indexes, monkeypatch):
    """The same with every count call forced through the packed transport (host_pack.hpp) in several chunks: the callers' pack
    jobs queue up in one pool, each call stages into its own context's pinned buffer."""
    case, g = get_case("log1m_sr32"), gpu_indexes("log1m_sr32")
    monkeypatch.setenv("FMGPU_HOST_PACK_MIN", "1")
    monkeypatch.setenv("FMGPU_PIPE_CHUNK", "1500")
    work = []
    for t in range(6):
        chars, off = make_patterns(case.text, 9000 + 700 * t, 1, 60, seed=200 + t)
        work.append((chars, off, case.oracle.count_batch(chars, off, threads=2)))
    errors = []

    def run(t):
        try:
            chars, off, (want, want_st) = work[t]
            for _ in range(8):
                got, st = g.count_batch(chars, off, return_status=True)
                assert np.array_equal(got, want) and np.array_equal(st, want_st)
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_device_calls_on_two_streams(gpu_indexes):
    """*_device calls return before their work has finished; two of them on different streams may end up sharing the handle's
    internal scratch buffers and must still both be right (the library orders such calls with events)."""
    import torch
    case, g = get_case("log1m_sr32"), gpu_indexes("log1m_sr32")
    dev = torch.device("cuda", g.device)
    batches = []
    for t in range(4):
        chars, off = make_patterns(case.text, 40_000, 4, 64, seed=200 + t)
        want = g.count_batch(chars, off)
        d_chars = torch.from_numpy(chars.view(np.int16)).to(dev)
        d_off = torch.from_numpy(off.view(np.int64)).to(dev)
        batches.append((d_chars, d_off, torch.zeros(off.size - 1, dtype=torch.int32, device=dev), want))
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(5):
        for t, (d_chars, d_off, d_counts, _) in enumerate(batches):
            g.count_batch_device(d_chars, d_off, d_counts, None, stream=streams[t % 2].cuda_stream)
    torch.cuda.synchronize()
    for d_chars, d_off, d_counts, want in batches:
        assert np.array_equal(d_counts.cpu().numpy(), want)


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_replicated_index_slices_the_batch(n_dev):
    """fmgpu_opts.devices: the index on n GPUs, every host-pointer call cut into n slices — same results as on one GPU,
    including a batch too small to cut and the global hit offsets of locate."""
    if _device_count() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    from index4j_b200 import FmIndex
    case = get_case("log3m_sr16")
    one = FmIndex.read(case.blob, device=0)
    many = FmIndex.read(case.blob, devices=list(range(n_dev)))
    try:
        assert many.devices == list(range(n_dev)) and many.getInputLength() == one.getInputLength()
        rng = np.random.default_rng(9)
        for n_pat in (1, 3, 5000, 60_001):
            chars, off = make_patterns(case.text, n_pat, 1, 48, seed=300 + n_pat)
            frm = rng.integers(-2, case.text.size + 3, max(n_pat // 4, 1)).astype(np.int32)
            _same(_ref_results(one, chars, off, frm), _ref_results(many, chars, off, frm))
            c8 = many.count_batch_utf8(np.minimum(chars, 127).astype(np.uint8), off)
            assert np.array_equal(c8, one.count_batch_utf8(np.minimum(chars, 127).astype(np.uint8), off))
            start = rng.integers(0, case.text.size - 100, max(n_pat // 4, 1)).astype(np.int32)
            stop = (start + rng.integers(0, 90, start.size)).astype(np.int32)
            a1, o1, l1, s1 = one.extract_batch(start, stop)
            a2, o2, l2, s2 = many.extract_batch(start, stop)
            assert np.array_equal(a1, a2) and np.array_equal(l1, l2) and np.array_equal(s1, s2)
        # two-phase sizing and the capacity error still work on the sliced call
        chars, off = make_patterns(case.text, 4000, 2, 20, seed=77)
        n_hits, hit_off, pos, st = many.locate_batch(chars, off, -1)
        assert int(hit_off[-1]) == int(n_hits.sum()) == pos.size
        assert np.array_equal(np.diff(hit_off.astype(np.int64)), n_hits)
    finally:
        one.close()
        many.close()
