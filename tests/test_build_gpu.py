"""Index production with the device stage (suffix array, BWT, sampled structures on the GPU): the serialized index
must be byte-identical with the host-only producer's, and the host producer fed with pieces computed by numpy must agree too."""
import numpy as np
import pytest

from conftest import get_case

from index4j_b200.builder import build_index, build_index_from_parts, map_text, suffix_array_host


def parts_numpy(text, sr):
    codes, sigma = map_text(text)
    sa = suffix_array_host(codes, sigma).astype(np.int64)
    length = codes.size
    bwt = codes[(sa - 1) % length]
    samp = (sa % sr) == 0
    bits = np.zeros(((length + 31) // 32) * 32, dtype=np.uint8)
    bits[:length] = samp
    mask = np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view("<u4").ravel()
    suffixes = sa[samp].astype(np.int32)
    positions = np.zeros(length // sr + 2, dtype=np.int32)
    positions[sa[samp] // sr] = np.nonzero(samp)[0]
    positions[(length - 1) // sr + 1] = positions[0]
    return bwt, mask, suffixes, positions


@pytest.mark.parametrize("name,sr,extract", [("log300k_sr64", 64, True), ("log200k_sr1", 1, True), ("multi400k_sr8", 8, True),
                                             ("tiny600k_sr4", 4, False), ("log300k_sr64", 7, True)])
def test_host_producer_from_parts_is_byte_identical(name, sr, extract):
    text = get_case(name).text[:150_000]
    bwt, mask, suffixes, positions = parts_numpy(text, sr)
    a = build_index(text, sr, extract, framed=False)
    b = build_index_from_parts(text, bwt, mask, suffixes, positions if extract else None, sr, extract, framed=False)
    assert a == b


def _unframe(blob: bytes) -> bytes:
    """Payload of an ObjectOutputStream of block-data records (0x77 len8 / 0x7A len32), Serialization.java:67-78."""
    assert blob[:4] == b"\xac\xed\x00\x05"
    out, i = bytearray(), 4
    while i < len(blob):
        if blob[i] == 0x77:
            n, i = blob[i + 1], i + 2
        else:
            assert blob[i] == 0x7A
            n, i = int.from_bytes(blob[i + 1: i + 5], "big"), i + 5
        out += blob[i: i + n]
        i += n
    return bytes(out)


def test_unframed_stream_is_the_payload_of_the_framed_one():
    """The raw writer (bulk big-endian stores) and the block-data writer (byte by byte) must serialize the same bytes."""
    text = get_case("log300k_sr64").text[:120_000]
    assert _unframe(build_index(text, 16, True, framed=True)) == build_index(text, 16, True, framed=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name,sr,extract", [("log1m_sr32", 32, True), ("log200k_sr1", 1, True), ("multi400k_sr8", 8, True),
                                             ("tiny600k_sr4", 4, False), ("log300k_sr64", 64, True)])
def test_device_stage_is_byte_identical(name, sr, extract):
    from index4j_b200.gpu_sa import build_index_gpu
    text = get_case(name).text
    assert build_index_gpu(text, sr, extract, framed=False) == build_index(text, sr, extract, framed=False)
