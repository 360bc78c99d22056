"""The sharded-index layer (index4j_b200/sharded.py) on the GPU engine: the hand-written composition kernels
(csrc/kernels_shard.cuh: ownership filter, global max_hits cut in rank order, pack, merge) around the two exchanges.  One GPU is
enough: the shards' ranks run as threads of this process and exchange through an in-process communicator with the semantics
of the NCCL calls (`DistComm`), so exactly the product code of the N > 1 path runs — kernels included.  The results are checked
against naive text scans and against the torch-op composition the CPU/gloo test uses (tests/test_sharded_gloo.py)."""
import threading

import numpy as np
import pytest
import torch

from conftest import get_case  # noqa: F401  (path setup)

import pyoracle

pytestmark = pytest.mark.gpu
MAXLEN = 24


class ThreadComm:
    """all_reduce / all_gather / all_gather_v among the threads of one process (all on the same GPU)."""

    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world

    def bind(self, rank):
        c = ThreadComm.__new__(ThreadComm)
        c.__dict__ = dict(self.__dict__)
        c.rank = rank
        c.root = self
        return c

    def _gather(self, t):
        torch.cuda.synchronize()
        self.root.slots[self.rank] = t
        self.root.bar.wait()
        got = [x.clone() for x in self.root.slots]
        self.root.bar.wait()
        return got

    def all_reduce_sum(self, t):
        t.copy_(torch.stack(self._gather(t)).sum(0))
        return t

    def all_gather(self, t):
        return torch.cat(self._gather(t.contiguous()))

    def all_gather_v(self, send, sizes):
        got = self._gather(send)
        assert [g.numel() for g in got] == [int(s) for s in sizes]
        return torch.cat(got) if got else send


def _run_sharded(text, world, chars, off, max_hits, force_torch=False):
    from index4j_b200 import FmIndex
    from index4j_b200.builder import build_index
    from index4j_b200.sharded import GpuEngine, ShardedFmIndex, pattern_tensors, shard_bounds
    dev = torch.device("cuda", 0)
    bounds = shard_bounds(text.size, world, MAXLEN)
    handles, shards = [], []
    root = ThreadComm(world)
    for r, (a, b, e) in enumerate(bounds):
        ix = FmIndex.read(build_index(text[a:e], 8), device=0)
        ov = FmIndex.read(build_index(text[b:e], 8), device=0) if e > b else None
        handles += [ix] + ([ov] if ov is not None else [])
        shards.append(ShardedFmIndex(GpuEngine(ix), GpuEngine(ov) if ov is not None else None, r, world, a, b - a, MAXLEN, comm=root.bind(r)))
    t_chars, t_off = pattern_tensors(chars, off, dev)
    res, errs = [None] * world, []

    def work(r):
        try:
            torch.cuda.set_device(0)
            sh = shards[r]
            counts = sh.count_batch(t_chars, t_off)
            if force_torch:
                n_hits, hit_off, pos = sh.engine.locate(t_chars, t_off, sh._cap(max_hits))
                out = sh._compose_torch(n_hits, hit_off, pos, max_hits)
            else:
                out = sh.locate_batch(t_chars, t_off, max_hits)
            torch.cuda.synchronize()
            res[r] = (counts.cpu().numpy(),) + tuple(x.cpu().numpy() for x in out)
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))
            root.bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for h in handles:
        h.close()
    assert not errs, errs
    return res, [b for (_, b, _) in bounds][:-1]


@pytest.mark.parametrize("world,max_hits", [(2, -1), (2, 5), (3, 3), (4, 40)])
def test_sharded_gpu_engines_one_gpu(world, max_hits):
    from index4j_b200.builder import gen_log_text
    from index4j_b200.sharded import shard_bounds
    text = gen_log_text(120_000, seed=78)
    rng = np.random.default_rng(6)
    cuts = [b for (_, b, _) in shard_bounds(text.size, world, MAXLEN)][:-1]
    pats = []
    for k in range(500):
        ln = int(rng.integers(1, MAXLEN + 1))
        if k % 3 == 0:  # straddling / touching a shard boundary
            cut = cuts[k % len(cuts)]
            s = int(rng.integers(cut - MAXLEN, cut + 2))
        else:
            s = int(rng.integers(0, text.size - ln))
        pats.append(text[s: s + ln])
    pats.append(np.array([0xFFFE], dtype=np.uint16))  # occurs nowhere
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in pats])
    chars = np.concatenate(pats).astype(np.uint16)
    res, cuts = _run_sharded(text, world, chars, off, max_hits)
    ref, _ = _run_sharded(text, world, chars, off, max_hits, force_torch=True)
    for r in range(world):
        for a, b in zip(res[r], res[0]):
            assert np.array_equal(a, b)  # every rank holds the same result
        for a, b in zip(res[r], ref[r]):
            assert np.array_equal(a, b)  # kernels == torch-op composition
    counts, n_hits, hit_off, pos = res[0]
    for i, p in enumerate(pats):
        loc = pyoracle.naive_locations(text, p)
        assert counts[i] == loc.size, i
        got = pos[hit_off[i]: hit_off[i + 1]]
        assert n_hits[i] == got.size
        if max_hits <= 0:
            assert np.array_equal(np.sort(got), loc), i
        else:
            assert got.size == min(loc.size, max_hits) and np.isin(got, loc).all() and np.unique(got).size == got.size, i
            owner = np.searchsorted(np.array(cuts), got, side="right")
            assert (np.diff(owner) >= 0).all(), i  # hits come in rank order
