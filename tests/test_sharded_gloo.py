"""The N>1 path of the sharded-index mode (index4j_b200/sharded.py) on CPU: world_size 2, gloo.

The composition logic (ownership filter, overlap correction, all-reduce of counts, all-gather of
positions, global max_hits cut) is the product code; only the per-shard engine is a stand-in here
(the CPU oracle behind the same tensor interface the GPU engine offers), because no GPU exists in
this container.  tests/test_gpu_sharded.py runs the same layer — with its CUDA composition kernels — on the GPU engine.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

MAXLEN = 24


class OracleEngine:
    def __init__(self, blob):
        import pyoracle
        self.o = pyoracle.OracleFmIndex(blob)

    def _np(self, chars, pat_off):
        return chars.numpy().view(np.uint16), pat_off.numpy().view(np.uint64)

    def count(self, chars, pat_off):
        c, st = self.o.count_batch(*self._np(chars, pat_off))
        assert not st.any()
        return torch.from_numpy(c.astype(np.int64))

    def locate(self, chars, pat_off, max_hits):
        ch, off = self._np(chars, pat_off)
        counts, _ = self.o.count_batch(ch, off)
        stride = int(max(1, counts.max() if max_hits <= 0 else min(max_hits, max(int(counts.max()), 1))))
        n_hits, pos, _ = self.o.locate_batch(ch, off, max_hits, stride)
        flat = np.concatenate([pos[i, : n_hits[i]] for i in range(n_hits.size)]) if n_hits.size else np.zeros(0, np.int32)
        hit_off = np.zeros(n_hits.size + 1, dtype=np.int64)
        hit_off[1:] = np.cumsum(n_hits)
        return torch.from_numpy(n_hits.astype(np.int64)), torch.from_numpy(hit_off), torch.from_numpy(flat.astype(np.int64))


def _worker(rank, world, port, text, chars, off, max_hits, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from index4j_b200.builder import build_index
        from index4j_b200.sharded import ShardedFmIndex, pattern_tensors, shard_bounds
        a, b, e = shard_bounds(text.size, world, MAXLEN)[rank]
        engine = OracleEngine(build_index(text[a:e], 8))
        overlap = OracleEngine(build_index(text[b:e], 8)) if e > b else None
        sh = ShardedFmIndex(engine, overlap, rank, world, a, b - a, MAXLEN)
        t_chars, t_off = pattern_tensors(chars, off, "cpu")
        counts = sh.count_batch(t_chars, t_off)
        n_hits, hit_off, pos = sh.locate_batch(t_chars, t_off, max_hits)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), counts=counts.numpy(), n_hits=n_hits.numpy(), hit_off=hit_off.numpy(), pos=pos.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,max_hits", [(2, -1), (2, 5), (3, 3)])
def test_sharded_count_and_locate(tmp_path, world, max_hits):
    import pyoracle
    from index4j_b200.builder import gen_log_text
    from index4j_b200.sharded import shard_bounds
    text = gen_log_text(60_000, seed=77)
    rng = np.random.default_rng(5)
    cuts = [b for (_, b, _) in shard_bounds(text.size, world, MAXLEN)][:-1]  # the shard boundaries
    pats = []
    for k in range(300):
        ln = int(rng.integers(1, MAXLEN + 1))
        if k % 3 == 0:  # straddling / touching a shard boundary
            cut = cuts[k % len(cuts)]
            s = int(rng.integers(cut - MAXLEN, cut + 2))
        else:
            s = int(rng.integers(0, text.size - ln))
        pats.append(text[s: s + ln])
    off = np.zeros(len(pats) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in pats])
    chars = np.concatenate(pats).astype(np.uint16)
    mp.spawn(_worker, args=(world, _free_port(), text, chars, off, max_hits, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    for k in ("counts", "n_hits", "hit_off", "pos"):
        for other in res[1:]:
            assert np.array_equal(res[0][k], other[k]), k  # every rank holds the same result
    r = res[0]
    for i, p in enumerate(pats):
        loc = pyoracle.naive_locations(text, p)
        assert r["counts"][i] == loc.size, i
        got = r["pos"][r["hit_off"][i]: r["hit_off"][i + 1]]
        if max_hits <= 0:
            assert np.array_equal(np.sort(got), loc), i
        else:
            assert got.size == min(loc.size, max_hits) and np.isin(got, loc).all() and np.unique(got).size == got.size, i
            # hits come in rank order: the owning shard of consecutive hits never decreases
            owner = np.searchsorted(np.array(cuts), got, side="right")
            assert (np.diff(owner) >= 0).all(), i
