import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "support")):
    if p not in sys.path:
        sys.path.insert(0, p)


# the packed transport of fmgpu_count_batch needs its host thread pool, whatever the size of the test machine
os.environ.setdefault("FMGPU_PACK_THREADS", "4")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def multiscript_text(n: int, seed: int = 7) -> np.ndarray:
    """Log-like lines mixing ASCII with Greek, Cyrillic, CJK and a few units above 32767 (like the
    reference's HDFS_2k_multichar fixture: > 700 distinct UTF-16 units, max > 32767)."""
    rng = np.random.default_rng(seed)
    pools = [np.arange(0x20, 0x7F), np.arange(0x391, 0x3C9), np.arange(0x410, 0x450), np.arange(0x4E00, 0x4E00 + 500),
             np.arange(0x9E00, 0x9E00 + 120), np.arange(0xAC00, 0xAC00 + 60)]
    words = []
    for _ in range(400):
        pool = pools[int(rng.integers(0, len(pools)))] if rng.random() < 0.5 else pools[0]
        words.append(rng.choice(pool, size=int(rng.integers(2, 9))).astype(np.uint16))
    out = []
    total = 0
    while total < n:
        line = [np.frombuffer(("%06d INFO " % int(rng.integers(0, 999999))).encode("utf-16-le"), dtype=np.uint16)]
        for _ in range(int(rng.integers(3, 12))):
            line.append(words[int(rng.integers(0, len(words)))])
            line.append(np.array([0x20], dtype=np.uint16))
        line.append(np.array([0x0A], dtype=np.uint16))
        a = np.concatenate(line)
        out.append(a)
        total += a.size
    return np.ascontiguousarray(np.concatenate(out)[:n], dtype=np.uint16)


def tiny_alphabet_text(n: int, sigma: int = 6, seed: int = 3) -> np.ndarray:
    """Few symbols, long runs: single-symbol blocks and blocks holding the whole alphabet (quirks Q3 / run blocks)."""
    rng = np.random.default_rng(seed)
    syms = np.array([ord("a") + i for i in range(sigma - 1)] + [0x0A], dtype=np.uint16)
    out = []
    total = 0
    while total < n:
        if rng.random() < 0.5:
            a = np.full(int(rng.integers(1, 3000)), syms[int(rng.integers(0, sigma))], dtype=np.uint16)
        else:
            a = rng.choice(syms, size=int(rng.integers(1, 400)))
        out.append(a)
        total += a.size
    return np.ascontiguousarray(np.concatenate(out)[:n], dtype=np.uint16)


def nul_text(n: int, k: int = 1000, seed: int = 12) -> np.ndarray:
    """Log text with k chars overwritten by \\0 — the reference's FmIndexTest.java:53-65,202-217: the text's own \\0 gets
    alphabet code 1 or later, the appended sentinel keeps code 0 (fm/FmIndex.java:398-415)."""
    from index4j_b200.builder import gen_log_text
    t = gen_log_text(n, seed=seed).copy()
    rng = np.random.default_rng(seed)
    t[rng.integers(0, n, k)] = 0
    return t


def big_code_runs_text(seed: int = 4) -> np.ndarray:
    """More than 256 distinct chars first, then long periodic stretches of chars that appear late (alphabet codes >= 256): the
    BWT then holds single-symbol blocks whose symbol code is >= 256, where inverseSelect keeps only the low byte (quirk Q1,
    wavelet/WaveletFixedBlockBoosting.java:1329-1332)."""
    rng = np.random.default_rng(seed)
    head = np.arange(0x4E00, 0x4E00 + 300, dtype=np.uint16)  # codes 1..300 in order of first appearance
    late = np.arange(0x5000, 0x5000 + 6, dtype=np.uint16)    # codes 301..306
    parts = [head]
    for _ in range(4):  # 100,000 (a b) pairs: BWT runs of 100,000 equal symbols, longer than the largest block (2^16)
        a, b = rng.choice(late, 2, replace=False)
        parts.append(np.tile(np.array([a, b], dtype=np.uint16), 100_000))
        parts.append(rng.choice(head, 200).astype(np.uint16))
        parts.append(np.array([0x0A], dtype=np.uint16))
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.uint16)


class Case:
    def __init__(self, name, text, sample_rate, extraction=True):
        from index4j_b200.builder import build_index
        self.name = name
        self.text = np.ascontiguousarray(text, dtype=np.uint16)
        self.sample_rate = sample_rate
        self.blob = build_index(self.text, sample_rate, extraction)
        self._oracle = None

    @property
    def oracle(self):
        if self._oracle is None:
            import pyoracle
            self._oracle = pyoracle.OracleFmIndex(self.blob)
        return self._oracle


_CASES = {}


def get_case(name: str) -> Case:
    if name not in _CASES:
        from index4j_b200.builder import gen_log_text
        if name == "log1m_sr32":
            c = Case(name, gen_log_text(1 << 20), 32)
        elif name == "log3m_sr16":
            c = Case(name, gen_log_text((3 << 20) + 12345, seed=99), 16)
        elif name == "log300k_sr64":
            c = Case(name, gen_log_text(300_000, seed=5), 64)
        elif name == "log200k_sr1":
            c = Case(name, gen_log_text(200_000, seed=6), 1)
        elif name == "multi400k_sr8":
            c = Case(name, multiscript_text(400_000), 8)
        elif name == "tiny600k_sr4":
            c = Case(name, tiny_alphabet_text(600_000), 4)
        elif name == "q4_2m_sr32":  # n + 1 = 2 * 2^20: rank(size, .) indexes past the superblock arrays (quirk Q4, :1022-1026)
            c = Case(name, gen_log_text((2 << 20) - 1, seed=21), 32)
        elif name == "nul1m_sr32":
            c = Case(name, nul_text(1 << 20), 32)
        elif name == "q1_runs_sr4":
            c = Case(name, big_code_runs_text(), 4)
        elif name == "cfg1_16m_sr32":  # BASELINE.json configs[0]: 16 MiB of log text, sampleRate 32, extraction on
            c = Case(name, gen_log_text(16 << 20), 32)
        elif name == "noextract":
            c = Case(name, gen_log_text(100_000, seed=8), 32, extraction=False)
        else:
            raise KeyError(name)
        _CASES[name] = c
    return _CASES[name]


CASE_NAMES = ["log1m_sr32", "log3m_sr16", "log300k_sr64", "log200k_sr1", "multi400k_sr8", "tiny600k_sr4"]
# the reference's quirk regions, one index each (Q4, text containing \\0, Q1)
QUIRK_CASE_NAMES = ["q4_2m_sr32", "nul1m_sr32", "q1_runs_sr4"]


def make_patterns(text: np.ndarray, n_pat: int, min_len: int, max_len: int, seed: int, absent_frac: float = 0.15):
    """Substrings of the text (the reference's JMH workload) mixed with mutated ones that mostly do not occur."""
    rng = np.random.default_rng(seed)
    pats = []
    for _ in range(n_pat):
        ln = int(rng.integers(min_len, max_len + 1))
        s = int(rng.integers(0, text.size - ln))
        p = text[s: s + ln].copy()
        r = rng.random()
        if r < absent_frac and ln > 0:
            p[int(rng.integers(0, ln))] = text[int(rng.integers(0, text.size))]
        elif r < absent_frac + 0.03 and ln > 0:
            p[int(rng.integers(0, ln))] = 0xFFFE  # not in the alphabet
        pats.append(p)
    off = np.zeros(n_pat + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in pats])
    return np.ascontiguousarray(np.concatenate(pats), dtype=np.uint16), off


@pytest.fixture(scope="module")
def gpu_indexes():
    """GPU-resident indexes of the test cases (loaded through the C ABI), one set per test module."""
    from index4j_b200 import FmIndex
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = FmIndex.read(get_case(name).blob)
        return cache[name]

    yield get
    for v in cache.values():
        v.close()
