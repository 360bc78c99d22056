"""UTF-8 byte patterns: FmIndex.convertBytePatternToCharPattern (fm/FmIndex.java:239-298) + count / locate.

CPU: the oracle's restatement of the converter against the reference's own known answers (FmIndexTest.java:126-160) and
against Python's codec on well-formed input; the device decoder (utf8_lane.h, replayed on the host) against the oracle on
well-formed, malformed and truncated byte strings.  GPU: fmgpu_count_batch_utf8 / fmgpu_locate_batch_utf8 against the oracle.
"""
import numpy as np
import pytest

from conftest import get_case, make_patterns

import flatcheck
import pyoracle


def test_converter_known_answers_of_the_reference():
    # FmIndexTest.shouldConvertFourByteUtf8 (:126-138): 'a', 11110_000 10_000000 10_000000 10_000000, 'c' -> 3 chars
    four = bytes([ord("a"), 0b11110000, 0b10000000, 0b10000000, 0b10000000, ord("c")])
    assert pyoracle.convert_byte_pattern_to_char_pattern(four).size == 3
    # FmIndexTest.shouldComplainFromTooBigChar (:140-158)
    big = bytes([ord("a"), 0b11110111, 0b10111000, 0b10111000, 0b10111000, ord("c")])
    with pytest.raises(pyoracle.JavaException) as e:
        pyoracle.convert_byte_pattern_to_char_pattern(big)
    assert str(e.value) == "Found a character that exceeds (32767): it was 2068024" and e.value.status == 10


def test_converter_matches_python_codec_on_bmp_text():
    s = "INFO dfs.DataNode: Übergröße ¿qué? Ελληνικά данные 数据 ログ\n"
    s = "".join(c for c in s if ord(c) <= 0xFFFF)
    got = pyoracle.convert_byte_pattern_to_char_pattern(s.encode("utf-8"))
    assert np.array_equal(got, np.array([ord(c) for c in s], dtype=np.uint16))


def _random_byte_strings(rng, n):
    out = [b"", b"a", b"\xc3", b"\xe2\x82", b"\xf0\x80\x80", b"\x80\x80", b"ab\xbf", b"\xff\xff\xff\xff", b"\xf0\x80\x81\xbfz"]
    for _ in range(n):
        k = int(rng.integers(1, 24))
        kind = rng.integers(0, 3)
        if kind == 0:  # well-formed text with multi-byte chars
            cps = rng.choice([0x41, 0x7A, 0xE9, 0x3B1, 0x20AC, 0x4E2D, 0x7FFF, 0xFFFD], size=k)
            out.append("".join(chr(int(c)) for c in cps).encode("utf-8"))
        elif kind == 1:  # arbitrary bytes
            out.append(bytes(rng.integers(0, 256, k, dtype=np.uint8)))
        else:  # mostly ASCII with a few high bytes
            b = rng.integers(32, 127, k, dtype=np.uint8)
            b[rng.integers(0, k)] = rng.integers(128, 256)
            out.append(bytes(b))
    return out


def test_device_decoder_matches_oracle():
    rng = np.random.default_rng(7)
    for data in _random_byte_strings(rng, 3000):
        try:
            want, st, val = pyoracle.convert_byte_pattern_to_char_pattern(data), 0, 0
        except pyoracle.JavaException as e:
            want, st, val = None, e.status, e.n
        got, got_st, got_val = flatcheck.utf8_convert(data)
        assert got_st == st, (data, st, got_st)
        if st == 0:
            assert np.array_equal(got, want), data
        elif st == 10:
            assert got_val == val, data


def _utf8_batch(case, n, seed):
    """Patterns of the case's text as UTF-8, plus malformed / truncated / over-limit ones."""
    chars, off = make_patterns(case.text, n, 1, 40, seed=seed)
    pats = []
    for i in range(off.size - 1):
        u = chars[int(off[i]): int(off[i + 1])]
        pats.append("".join(chr(int(c)) for c in u).encode("utf-8", "surrogatepass"))
    rng = np.random.default_rng(seed)
    pats += _random_byte_strings(rng, 300)
    boff = np.zeros(len(pats) + 1, dtype=np.uint64)
    boff[1:] = np.cumsum([len(p) for p in pats])
    return np.frombuffer(b"".join(pats), dtype=np.uint8).copy(), boff


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["log1m_sr32", "multi400k_sr8"])
def test_count_utf8_matches_oracle(gpu_indexes, name):
    case, g = get_case(name), gpu_indexes(name)
    data, boff = _utf8_batch(case, 3000, 31)
    want, want_st = case.oracle.count_batch_utf8(data, boff, threads=4)
    got, got_st = g.count_batch_utf8(data, boff, return_status=True)
    assert np.array_equal(got_st, want_st)
    assert np.array_equal(got, want)
    assert int((want_st == 0).sum()) > 2500 and int((want_st == 10).sum()) > 0 and int((want_st == 9).sum()) > 0
    assert int((want[want_st == 0] > 0).sum()) > 2000
    # device-resident entry point
    import torch
    dev = torch.device("cuda", g.device)
    d_counts = torch.zeros(boff.size - 1, dtype=torch.int32, device=dev)
    d_status = torch.zeros(boff.size - 1, dtype=torch.int32, device=dev)
    g.count_batch_utf8_device(torch.from_numpy(data).to(dev), torch.from_numpy(boff.view(np.int64)).to(dev), d_counts, d_status)
    torch.cuda.synchronize()
    assert np.array_equal(d_counts.cpu().numpy(), want) and np.array_equal(d_status.cpu().numpy(), want_st)
    # single-query form raises like the reference
    assert g.count_utf8(b"INFO") == case.oracle.count("INFO")
    with pytest.raises(Exception) as e:
        g.count_utf8(bytes([ord("a"), 0b11110111, 0b10111000, 0b10111000, 0b10111000, ord("c")]))
    assert "Found a character that exceeds (32767): it was 2068024" in str(e.value)


@pytest.mark.gpu
def test_locate_utf8_matches_oracle(gpu_indexes):
    case, g = get_case("log1m_sr32"), gpu_indexes("log1m_sr32")
    data, boff = _utf8_batch(case, 800, 33)
    want_n, want_pos, want_st = case.oracle.locate_batch_utf8(data, boff, 50, 50, threads=4)
    n_hits, hit_off, pos, st = g.locate_batch_utf8(data, boff, 50)
    assert np.array_equal(st, want_st) and np.array_equal(n_hits, want_n)
    for i in range(want_n.size):
        assert np.array_equal(pos[int(hit_off[i]): int(hit_off[i + 1])], want_pos[i, : want_n[i]]), i
