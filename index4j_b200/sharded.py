"""Texts beyond Java's 2^31-char array limit: one FmIndex per rank over a text shard, NCCL for the exchange.

The reference cannot index more than 2^31-2 chars (``int length``, ``char[] input``:
indices/src/main/java/com/dynatrace/fm/FmIndex.java:131,155,335-341) and has no multi-device code,
so this layer has no Java counterpart; it composes per-shard results that are each bit-exact with
the Java ``FmIndex`` of that shard (BASELINE.json configs[4]).

Layout: the text is cut into ``world`` shards of ``shard_len`` chars; shard r additionally holds the
first ``max_pattern_len - 1`` chars of shard r+1 (overlap), so every occurrence of a pattern of at
most ``max_pattern_len`` chars lies wholly inside at least one shard.  An occurrence is OWNED by the
shard in which it starts before ``shard_len``.

* count: each rank counts in its shard and subtracts the occurrences lying wholly inside its overlap
  tail (counted again by the next rank) using a tiny FmIndex over just that tail; the per-rank
  vectors are summed with one all-reduce.
* locate: each rank locates ``max_hits + overlap`` rows, drops hits it does not own, adds its text
  offset (positions become int64) and the ranks all-gather hit counts and positions.  The global
  cut at ``max_hits`` keeps hits in rank order (lowest shard first, SA order inside a shard).

One process per GPU; every rank gets the whole pattern batch.  The engine object only has to offer
``count(chars, pat_off) -> int64[n]`` and ``locate(chars, pat_off, max_hits) -> (n_hits, hit_off,
positions)`` on tensors of the process group's device, so the same code runs on NCCL/CUDA with the
GPU engine and (tests) on gloo/CPU with a stand-in engine.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, max_pattern_len: int):
    """[(text_start, owned_end, end_with_overlap)] per rank."""
    shard_len = (n_total + world - 1) // world
    out = []
    for r in range(world):
        a = min(r * shard_len, n_total)
        b = min((r + 1) * shard_len, n_total)
        e = min(b + max_pattern_len - 1, n_total)
        out.append((a, b, e))
    return out


class GpuEngine:
    """Adapter: :class:`index4j_b200.FmIndex` -> the tensor interface used below (device-resident, current stream)."""

    def __init__(self, fm_index):
        self.ix = fm_index
        self.device = torch.device("cuda", fm_index.device)

    def count(self, chars: torch.Tensor, pat_off: torch.Tensor) -> torch.Tensor:
        n = pat_off.numel() - 1
        counts = torch.empty(n, dtype=torch.int32, device=self.device)
        self.ix.count_batch_device(chars, pat_off, counts, None)
        return counts.to(torch.int64)

    def locate(self, chars: torch.Tensor, pat_off: torch.Tensor, max_hits: int):
        n = pat_off.numel() - 1
        n_hits = torch.empty(n, dtype=torch.int32, device=self.device)
        hit_off = torch.empty(n + 1, dtype=torch.int64, device=self.device)
        total = self.ix.locate_batch_device(chars, pat_off, max_hits, n_hits, hit_off, None)  # sizing pass
        pos = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        if total:
            self.ix.locate_batch_device(chars, pat_off, max_hits, n_hits, hit_off, pos)
        return n_hits.to(torch.int64), hit_off, pos[:total].to(torch.int64)


class ShardedFmIndex:
    def __init__(self, engine, overlap_engine, rank: int, world: int, text_start: int, owned_len: int, max_pattern_len: int,
                 group=None):
        """``engine``: this rank's shard index; ``overlap_engine``: index over the shard's overlap tail only
        (``None`` on the last rank / when the tail is empty)."""
        self.engine, self.overlap_engine = engine, overlap_engine
        self.rank, self.world = rank, world
        self.text_start, self.owned_len, self.max_pattern_len = int(text_start), int(owned_len), int(max_pattern_len)
        self.group = group

    # ------------------------------------------------------------------------------------------
    def _check_lengths(self, pat_off: torch.Tensor):
        lens = pat_off[1:] - pat_off[:-1]
        if lens.numel() and int(lens.max()) > self.max_pattern_len:
            raise ValueError("pattern longer than the shard overlap allows (%d)" % self.max_pattern_len)

    def count_batch(self, chars: torch.Tensor, pat_off: torch.Tensor) -> torch.Tensor:
        """Occurrences in the whole text, int64[n]; identical on every rank."""
        self._check_lengths(pat_off)
        local = self.engine.count(chars, pat_off)
        if self.overlap_engine is not None:
            local = local - self.overlap_engine.count(chars, pat_off)
        if self.world > 1:
            dist.all_reduce(local, op=dist.ReduceOp.SUM, group=self.group)
        return local

    def owned_hits(self, chars: torch.Tensor, pat_off: torch.Tensor, max_hits: int):
        """This rank's owned hits as global int64 positions: (n_hits int64[n], positions int64[sum])."""
        self._check_lengths(pat_off)
        n = pat_off.numel() - 1
        cap = max_hits + self.max_pattern_len - 1 if max_hits > 0 else max_hits
        n_hits, hit_off, pos = self.engine.locate(chars, pat_off, cap)
        dev = pos.device
        pat_id = torch.repeat_interleave(torch.arange(n, device=dev), n_hits)
        keep = pos < self.owned_len
        kept_pat = pat_id[keep]
        kept_pos = pos[keep]
        if max_hits > 0 and kept_pos.numel():
            kept_n = torch.bincount(kept_pat, minlength=n)
            kept_off = torch.cumsum(kept_n, 0) - kept_n
            t = torch.arange(kept_pos.numel(), device=dev) - kept_off[kept_pat]
            sel = t < max_hits
            kept_pat, kept_pos = kept_pat[sel], kept_pos[sel]
        out_n = torch.bincount(kept_pat, minlength=n) if kept_pat.numel() else torch.zeros(n, dtype=torch.int64, device=dev)
        return out_n, kept_pos + self.text_start

    def locate_batch(self, chars: torch.Tensor, pat_off: torch.Tensor, max_hits: int = -1):
        """-> (n_hits int64[n], hit_off int64[n+1], positions int64[total]) — identical on every rank.
        Hits of a pattern are ordered by rank, then SA order inside the rank's shard; at most ``max_hits`` are kept."""
        n = pat_off.numel() - 1
        my_n, my_pos = self.owned_hits(chars, pat_off, max_hits)
        dev = my_n.device
        if self.world == 1:
            off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            off[1:] = torch.cumsum(my_n, 0)
            return my_n, off, my_pos
        # exchange: per-rank hit counts, then positions padded to the largest rank
        all_n = torch.empty(self.world * n, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_n, my_n.contiguous(), group=self.group)
        all_n = all_n.view(self.world, n)
        totals = all_n.sum(1)
        pad = int(totals.max().item())
        send = torch.zeros(max(pad, 1), dtype=torch.int64, device=dev)
        send[: my_pos.numel()] = my_pos
        recv = torch.empty(self.world * max(pad, 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        recv = recv.view(self.world, max(pad, 1))
        # global cut at max_hits in rank order
        before = torch.cumsum(all_n, 0) - all_n  # hits of lower ranks, per pattern
        if max_hits > 0:
            keep_n = torch.clamp(torch.minimum(all_n, max_hits - before), min=0)
        else:
            keep_n = all_n
        kept_before = torch.cumsum(keep_n, 0) - keep_n
        n_hits = keep_n.sum(0)
        hit_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        hit_off[1:] = torch.cumsum(n_hits, 0)
        # one pass over all (rank, pattern) segments: segment (r, p) holds all_n[r][p] hits in recv[r] from roff[r][p]; its first
        # keep_n[r][p] go to out[hit_off[p] + kept_before[r][p] ...]  (no per-rank loop, one host sync for the output size)
        total = int(hit_off[-1].item())
        out = torch.empty(total, dtype=torch.int64, device=dev)
        if total:
            stride = max(pad, 1)
            roff = torch.cumsum(all_n, 1) - all_n
            src0 = (roff + torch.arange(self.world, device=dev).unsqueeze(1) * stride).reshape(-1)
            dst0 = (hit_off[:-1].unsqueeze(0) + kept_before).reshape(-1)
            seg_len = keep_n.reshape(-1)
            seg = torch.repeat_interleave(torch.arange(seg_len.numel(), device=dev), seg_len, output_size=total)
            seg_first = torch.cumsum(seg_len, 0) - seg_len
            t = torch.arange(total, device=dev) - seg_first[seg]
            out[dst0[seg] + t] = recv.reshape(-1)[src0[seg] + t]
        return n_hits, hit_off, out


def pattern_tensors(chars: np.ndarray, pat_off: np.ndarray, device):
    """numpy (uint16 chars, uint64 offsets) -> tensors the engines take (int16 / int64 views)."""
    return (torch.from_numpy(np.ascontiguousarray(chars, dtype=np.uint16).view(np.int16)).to(device),
            torch.from_numpy(np.ascontiguousarray(pat_off, dtype=np.uint64).view(np.int64)).to(device))
