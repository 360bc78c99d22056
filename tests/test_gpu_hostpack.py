"""Packed transport of fmgpu_count_batch (csrc/host_pack.hpp): char[] chunks whose chars fit a byte cross PCIe as bytes +
chunk-relative uint32 offsets (narrowed by the library's host threads into page-locked staging, widened again on the device);
chunks with a char above 0xFF go as they are.  Counts and statuses are those of the direct path and of the oracle."""
import numpy as np
import pytest

from conftest import get_case, make_patterns

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["log1m_sr32", "multi400k_sr8", "nul1m_sr32", "tiny600k_sr4"])
@pytest.mark.parametrize("chunk", [0, 700])
def test_packed_count_matches_direct_and_oracle(gpu_indexes, name, chunk, monkeypatch):
    case, g = get_case(name), gpu_indexes(name)
    chars, off = make_patterns(case.text, 6000, 0, 48, seed=13)
    if name == "log1m_sr32":  # a few chars above 0xFF inside otherwise Latin-1 chunks: those chunks must go unpacked
        chars = chars.copy()
        chars[[5, 40_000, 90_000]] = [0x4E2D, 0x0100, 0xFFFF]
    want, want_st = case.oracle.count_batch(chars, off, threads=4)
    monkeypatch.setenv("FMGPU_HOST_PACK", "0")
    direct, direct_st = g.count_batch(chars, off, return_status=True)
    monkeypatch.setenv("FMGPU_HOST_PACK", "1")
    monkeypatch.setenv("FMGPU_HOST_PACK_MIN", "1")
    if chunk:
        monkeypatch.setenv("FMGPU_PIPE_CHUNK", str(chunk))  # several chunks per call
    packed, packed_st = g.count_batch(chars, off, return_status=True)
    assert np.array_equal(direct, want) and np.array_equal(direct_st, want_st)
    assert np.array_equal(packed, want) and np.array_equal(packed_st, want_st)
    # an empty batch and a batch of empty patterns
    e_counts, e_st = g.count_batch(np.zeros(0, np.uint16), np.zeros(4, np.uint64), return_status=True)
    assert np.array_equal(e_st, [9, 9, 9])


def test_packed_count_from_unpinned_arrays(gpu_indexes, monkeypatch):
    """The packed path reads the caller's arrays with the CPU: plain (pageable) numpy arrays, odd alignments."""
    case, g = get_case("log1m_sr32"), gpu_indexes("log1m_sr32")
    chars, off = make_patterns(case.text, 20000, 1, 64, seed=14)
    want, want_st = case.oracle.count_batch(chars, off, threads=4)
    monkeypatch.setenv("FMGPU_HOST_PACK_MIN", "1")
    monkeypatch.setenv("FMGPU_PIPE_CHUNK", "3000")
    buf = np.zeros(chars.size + 1, dtype=np.uint16)
    view = buf[1:]  # 2-byte aligned only
    view[:] = chars
    got, got_st = g.count_batch(view, off, return_status=True)
    assert np.array_equal(got, want) and np.array_equal(got_st, want_st)
