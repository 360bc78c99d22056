"""Fused locate -> extractUntilBoundary on the GPU (fmgpu_extract_records_batch / fmgpu_locate_records_batch): the reference's
"extracting whole records" flow (README.md:98-107) with every distinct record read once.  Per hit the result must be exactly
what the oracle's locate + extractUntilBoundary(locations[i], new char[dst_len], 0, '\\n') composition gives."""
import numpy as np
import pytest

from conftest import get_case, make_patterns

pytestmark = pytest.mark.gpu


def _check_hits(case, frm, dst_len, idx, ln, st, records):
    n = case.text.size
    w_arena, w_len, w_st = case.oracle.extract_until_boundary_batch(frm, 10, dst_len, 0, threads=4)
    assert np.array_equal(st, w_st), np.flatnonzero(st != w_st)[:10]
    ok = (w_st == 0) | (w_st == 8)
    assert np.array_equal(ln[ok], w_len[ok])
    good = np.flatnonzero((w_st == 0) & (w_len > 0) & (frm < n))
    assert (idx[good] >= 0).all() and (idx[good] < records.shape[0]).all()
    cols = np.arange(records.shape[1])[None, :]
    rows = records[idx[good]]
    valid = cols < w_len[good][:, None]
    assert np.array_equal(rows[valid], w_arena[good][valid])
    return good.size


@pytest.mark.parametrize("name", ["log1m_sr32", "log300k_sr64", "tiny600k_sr4", "multi400k_sr8", "nul1m_sr32", "q4_2m_sr32"])
@pytest.mark.parametrize("dst_len", [512, 130, 37, 5])
def test_extract_records_matches_oracle(gpu_indexes, name, dst_len):
    case, g = get_case(name), gpu_indexes(name)
    n = case.text.size
    rng = np.random.default_rng(60 + dst_len)
    centers = rng.integers(0, n, 600)
    near = (centers[:, None] + rng.integers(-60, 60, (600, 8))).reshape(-1)  # several hits per record
    nl = np.flatnonzero(case.text == 10)
    at_nl = nl[rng.integers(0, nl.size, 100)]
    frm = np.concatenate([near, at_nl, at_nl + 1, at_nl - 1, rng.integers(0, n, 2000), np.arange(n - 14, n + 2), np.arange(-1, 8)])
    frm = np.clip(frm, -1, n + 1).astype(np.int32)
    idx, ln, st, records = g.extract_records_batch(frm, "\n", dst_len)
    good = _check_hits(case, frm, dst_len, idx, ln, st, records)
    if dst_len == 512 and name.startswith("log"):
        assert good > 4000 and records.shape[0] < 0.6 * frm.size  # clustered hits share their records
    # too few arena rows: the call reports how many it needs
    from index4j_b200.fm_index import FmIndexError
    if records.shape[0] > 10:
        with pytest.raises(FmIndexError, match="distinct records"):
            g.extract_records_batch(frm, "\n", dst_len, rec_cap=records.shape[0] - 1)
    # error contract of the scan (checkBoundsForExtraction / unknown boundary), as the plain call
    i2, l2, s2, r2 = g.extract_records_batch(frm[:50], "一", dst_len)
    a3, l3, s3 = g.extract_until_boundary_batch(frm[:50], "一", dst_len, 0)
    assert np.array_equal(s2, s3) and r2.shape[0] == 0


@pytest.mark.parametrize("name", ["log1m_sr32", "multi400k_sr8"])
@pytest.mark.parametrize("max_hits", [-1, 20])
def test_locate_records_matches_oracle(gpu_indexes, name, max_hits):
    case, g = get_case(name), gpu_indexes(name)
    chars, off = make_patterns(case.text, 800, 3, 30, seed=71)
    out = g.locate_records_batch(chars, off, max_hits, "\n", 512)
    n_hits, hit_off, pos, st = g.locate_batch(chars, off, max_hits)
    assert np.array_equal(out["n_hits"], n_hits) and np.array_equal(out["hit_off"], hit_off) and np.array_equal(out["positions"], pos)
    assert np.array_equal(out["pat_status"], st)
    good = _check_hits(case, pos.astype(np.int32), 512, out["rec_index"], out["len"], out["status"], out["records"])
    assert good > 500
    # empty batch / batch without hits
    e = g.locate_records_batch(np.array([0xFFFE], dtype=np.uint16), np.array([0, 1], dtype=np.uint64), -1, "\n", 64)
    assert e["positions"].size == 0 and e["records"].shape[0] == 0 and e["n_hits"][0] == 0


def test_extract_records_device_api(gpu_indexes):
    import torch
    case, g = get_case("log1m_sr32"), gpu_indexes("log1m_sr32")
    dev = torch.device("cuda", g.device)
    rng = np.random.default_rng(5)
    frm = rng.integers(0, case.text.size, 30_000).astype(np.int32)
    d_from = torch.from_numpy(frm).to(dev)
    d_idx = torch.empty(frm.size, dtype=torch.int32, device=dev)
    d_len = torch.empty(frm.size, dtype=torch.int32, device=dev)
    d_st = torch.empty(frm.size, dtype=torch.int32, device=dev)
    d_arena = torch.zeros((frm.size, 256), dtype=torch.int16, device=dev)
    n_rec = g.extract_records_batch_device(d_from, "\n", 256, d_idx, d_len, d_st, d_arena)
    torch.cuda.synchronize()
    records = d_arena[:n_rec].cpu().numpy().view(np.uint16)
    _check_hits(case, frm, 256, d_idx.cpu().numpy(), d_len.cpu().numpy(), d_st.cpu().numpy(), records)
    assert n_rec < frm.size  # 30,000 hits in ~8,900 lines
