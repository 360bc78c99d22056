"""Robustness of the stream parser + re-layout (jstream.hpp, flatten.hpp — the code fmgpu_index_load_serialized runs): truncated and
bit-flipped FmIndex streams are either rejected with a format error (IOException in the reference) or load into a self-consistent
layout on which the lane code answers queries without leaving its arrays.  Runs the product's host code through tests/support."""
import numpy as np

from conftest import get_case, make_patterns

import flatcheck


def test_mutated_streams_never_fault():
    case = get_case("log300k_sr64")
    blob = case.blob
    chars, off = make_patterns(case.text, 200, 1, 24, seed=5)
    rng = np.random.default_rng(1)
    loaded = rejected = 0
    for t in range(160):
        b = bytearray(blob)
        kind = t % 4
        if kind == 0:  # truncation
            b = b[: int(rng.integers(0, len(b)))]
        elif kind == 1:  # header / directory bytes
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, min(len(b), 4096)))] = int(rng.integers(0, 256))
        elif kind == 2:  # bit flips anywhere
            for _ in range(int(rng.integers(1, 20))):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif t % 8 == 3:  # a huge count / length field somewhere
            p = int(rng.integers(0, len(b) - 8))
            b[p: p + 4] = bytes([0x7F, 0xFF, 0xFF, 0xFF])
        else:  # byte noise in the wavelet part (directories, block headers, level bits)
            for _ in range(int(rng.integers(1, 30))):
                b[int(rng.integers(len(b) // 2, len(b)))] = int(rng.integers(0, 256))
        try:
            f = flatcheck.FlatIndexHost(bytes(b), threads=2)
        except IOError:
            rejected += 1
            continue
        loaded += 1
        # wrong answers are fine, faults are not: every lane family walks the mutated index
        counts, st, ranges = f.count_batch(chars, off)
        rows = np.unique(np.clip(ranges.ravel(), 0, case.text.size)).astype(np.uint32)[:200]
        f.locate_rows(rows)
        start = rng.integers(0, case.text.size - 100, 60).astype(np.int32)
        f.extract(start, start + 40, np.arange(61, dtype=np.uint64) * 40)
        f.eub(start, 10, 128, 0)
        del f
    assert rejected > 40 and loaded + rejected == 160
