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
* locate: each rank locates ``max_hits + overlap`` rows; the hand-written kernels of
  ``csrc/kernels_shard.cuh`` (C ABI ``fmgpu_shard_*``) count the hits the shard owns, plan the global cut
  at ``max_hits`` (lowest shard first, SA order inside a shard) from the all-gathered counts, pack this
  rank's contribution as global int64 positions, and merge every rank's contribution into the final
  per-pattern order.  The positions travel in ONE grouped NCCL send/recv with their exact sizes (no
  padding); the only host read of the whole call is the ``world + 1`` contribution sizes that the
  receive buffers are allocated from.

One process per GPU; every rank gets the whole pattern batch.  On CPU tensors (the gloo test of the
composition logic, with a stand-in engine) the same plan is computed with torch ops
(``_compose_torch``), which the GPU test also uses to cross-check the kernels.
"""
from __future__ import annotations

import ctypes as C

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


class DistComm:
    """The exchanges of the sharded layer over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU test)."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """equal-sized contributions -> [world * n]"""
        if self.world == 1:
            return t.clone()
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def all_gather_v(self, send: torch.Tensor, sizes) -> torch.Tensor:
        """contributions of different sizes (``sizes[r]`` elements from rank r), concatenated in rank order: one grouped send/recv"""
        if self.world == 1:
            return send
        recv = torch.empty(int(sum(sizes)), dtype=send.dtype, device=send.device)
        views, at = [], 0
        for r in range(self.world):
            views.append(recv[at: at + int(sizes[r])])
            at += int(sizes[r])
        views[self.rank].copy_(send)
        ops = []
        for r in range(self.world):
            if r == self.rank:
                continue
            if send.numel():
                ops.append(dist.P2POp(dist.isend, send, r, group=self.group))
            if views[r].numel():
                ops.append(dist.P2POp(dist.irecv, views[r], r, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv


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

    def locate_raw(self, chars: torch.Tensor, pat_off: torch.Tensor, max_hits: int):
        """-> (n_hits int32[n], hit_off int64[n+1], positions int32[total]) straight from fmgpu_locate_batch_device"""
        n = pat_off.numel() - 1
        n_hits = torch.empty(n, dtype=torch.int32, device=self.device)
        hit_off = torch.empty(n + 1, dtype=torch.int64, device=self.device)
        total = self.ix.locate_batch_device(chars, pat_off, max_hits, n_hits, hit_off, None)  # sizing pass
        pos = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        if total:
            self.ix.locate_batch_device(chars, pat_off, max_hits, n_hits, hit_off, pos)
        return n_hits, hit_off, pos[:total]

    def locate(self, chars: torch.Tensor, pat_off: torch.Tensor, max_hits: int):
        n_hits, hit_off, pos = self.locate_raw(chars, pat_off, max_hits)
        return n_hits.to(torch.int64), hit_off, pos.to(torch.int64)


def _shard_lib():
    from .fm_index import native
    L = native()
    if not getattr(L, "_shard_bound", False):
        vp, i32, u32, i64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64
        L.fmgpu_shard_keep_device.argtypes = [vp, vp, u32, i32, i32, vp, vp]
        L.fmgpu_shard_plan_device.argtypes = [vp, u32, u32, u32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.fmgpu_shard_pack_device.argtypes = [vp, vp, u32, i32, i64, vp, vp, vp, vp]
        L.fmgpu_shard_merge_device.argtypes = [vp, vp, vp, vp, vp, u32, u32, vp, vp]
        L._shard_bound = True
    return L


class ShardedFmIndex:
    def __init__(self, engine, overlap_engine, rank: int, world: int, text_start: int, owned_len: int, max_pattern_len: int,
                 group=None, comm=None):
        """``engine``: this rank's shard index; ``overlap_engine``: index over the shard's overlap tail only
        (``None`` on the last rank / when the tail is empty)."""
        self.engine, self.overlap_engine = engine, overlap_engine
        self.rank, self.world = rank, world
        self.text_start, self.owned_len, self.max_pattern_len = int(text_start), int(owned_len), int(max_pattern_len)
        self.comm = comm if comm is not None else DistComm(rank, world, group)

    # ------------------------------------------------------------------------------------------
    def _check_lengths(self, pat_off: torch.Tensor):
        lens = pat_off[1:] - pat_off[:-1]
        if lens.numel() and int(lens.max()) > self.max_pattern_len:
            raise ValueError("pattern longer than the shard overlap allows (%d)" % self.max_pattern_len)

    def count_batch(self, chars: torch.Tensor, pat_off: torch.Tensor, check_lengths: bool = True) -> torch.Tensor:
        """Occurrences in the whole text, int64[n]; identical on every rank."""
        if check_lengths:
            self._check_lengths(pat_off)
        local = self.engine.count(chars, pat_off)
        if self.overlap_engine is not None:
            local = local - self.overlap_engine.count(chars, pat_off)
        return self.comm.all_reduce_sum(local)

    def _cap(self, max_hits: int) -> int:
        # at most max_pattern_len - 1 of a shard's hits start in its overlap tail: locating that many more rows than max_hits
        # leaves max_hits owned ones whenever the shard has them
        return max_hits + self.max_pattern_len - 1 if max_hits > 0 else max_hits

    def locate_batch(self, chars: torch.Tensor, pat_off: torch.Tensor, max_hits: int = -1, check_lengths: bool = True):
        """-> (n_hits int64[n], hit_off int64[n+1], positions int64[total]) — identical on every rank.
        Hits of a pattern are ordered by rank, then SA order inside the rank's shard; at most ``max_hits`` are kept."""
        if check_lengths:
            self._check_lengths(pat_off)
        if chars.is_cuda and hasattr(self.engine, "locate_raw"):
            return self._locate_batch_kernels(chars, pat_off, max_hits)
        n_hits, hit_off, pos = self.engine.locate(chars, pat_off, self._cap(max_hits))
        return self._compose_torch(n_hits, hit_off, pos, max_hits)

    # --- GPU: hand-written kernels (csrc/kernels_shard.cuh) around the two exchanges -----------------------------------
    def _locate_batch_kernels(self, chars, pat_off, max_hits):
        L = _shard_lib()
        dev = chars.device
        st = torch.cuda.current_stream(dev).cuda_stream
        n, W = pat_off.numel() - 1, self.world
        n_loc, hit_off_loc, pos = self.engine.locate_raw(chars, pat_off, self._cap(max_hits))
        i32, i64 = torch.int32, torch.int64
        kept = torch.empty(n, dtype=i32, device=dev)
        rc = L.fmgpu_shard_keep_device(pos.data_ptr(), hit_off_loc.data_ptr(), n, self.owned_len, max_hits, kept.data_ptr(), st)
        all_kept = self.comm.all_gather(kept)                                   # exchange 1: int32[world][n]
        take = torch.empty(W * n, dtype=i32, device=dev)
        n_hits = torch.empty(n, dtype=i32, device=dev)
        hit_off = torch.empty(n + 1, dtype=i64, device=dev)
        roff = torch.empty(W * (n + 1), dtype=i64, device=dev)
        totals = torch.empty(W + 1, dtype=i64, device=dev)
        rank_base = torch.empty(W, dtype=i64, device=dev)
        rc |= L.fmgpu_shard_plan_device(all_kept.data_ptr(), n, W, self.rank, max_hits, take.data_ptr(), n_hits.data_ptr(), hit_off.data_ptr(),
                                        roff.data_ptr(), totals.data_ptr(), rank_base.data_ptr(), st)
        sizes = totals.cpu().tolist()                                            # the one host read: world + 1 sizes
        send = torch.empty(max(int(sizes[self.rank]), 1), dtype=i64, device=dev)[: int(sizes[self.rank])]
        rc |= L.fmgpu_shard_pack_device(pos.data_ptr(), hit_off_loc.data_ptr(), n, self.owned_len, self.text_start,
                                        take[self.rank * n:].data_ptr(), roff[self.rank * (n + 1):].data_ptr(), send.data_ptr(), st)
        recv = self.comm.all_gather_v(send, sizes[:W])                           # exchange 2: exact sizes, one grouped send/recv
        out = torch.empty(max(int(sizes[W]), 1), dtype=i64, device=dev)[: int(sizes[W])]
        rc |= L.fmgpu_shard_merge_device(recv.data_ptr(), rank_base.data_ptr(), roff.data_ptr(), take.data_ptr(), hit_off.data_ptr(), n, W,
                                         out.data_ptr(), st)
        if rc:
            raise RuntimeError(L.fmgpu_last_error().decode())
        return n_hits.to(i64), hit_off, out

    # --- the same plan with torch ops (CPU / gloo test; cross-check of the kernels) --------------------------------------
    def owned_hits(self, n_hits, hit_off, pos, max_hits: int):
        """This rank's owned hits as global int64 positions: (kept int64[n], positions int64[sum])."""
        n = n_hits.numel()
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

    def _compose_torch(self, n_hits_loc, hit_off_loc, pos, max_hits):
        n, W = n_hits_loc.numel(), self.world
        my_n, my_pos = self.owned_hits(n_hits_loc.to(torch.int64), hit_off_loc, pos.to(torch.int64), max_hits)
        dev = my_n.device
        all_n = self.comm.all_gather(my_n.contiguous()).view(W, n)
        before = torch.cumsum(all_n, 0) - all_n  # hits of lower ranks, per pattern
        take = torch.clamp(torch.minimum(all_n, max_hits - before), min=0) if max_hits > 0 else all_n
        n_hits = take.sum(0)
        hit_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        hit_off[1:] = torch.cumsum(n_hits, 0)
        # this rank sends only what survives the global cut: the first take[rank][p] of its kept hits of pattern p
        my_off = torch.cumsum(my_n, 0) - my_n
        pat_id = torch.repeat_interleave(torch.arange(n, device=dev), my_n)
        t = torch.arange(my_pos.numel(), device=dev) - my_off[pat_id]
        send = my_pos[t < take[self.rank][pat_id]]
        sizes = take.sum(1).tolist()
        recv = self.comm.all_gather_v(send.contiguous(), sizes)
        total = int(hit_off[-1])
        out = torch.empty(total, dtype=torch.int64, device=dev)
        if total:
            roff = torch.cumsum(take, 1) - take
            rank_base = torch.tensor(np.concatenate([[0], np.cumsum(sizes)[:-1]]), dtype=torch.int64, device=dev)
            kept_before = torch.cumsum(take, 0) - take
            src0 = (roff + rank_base.unsqueeze(1)).reshape(-1)
            dst0 = (hit_off[:-1].unsqueeze(0) + kept_before).reshape(-1)
            seg_len = take.reshape(-1)
            seg = torch.repeat_interleave(torch.arange(seg_len.numel(), device=dev), seg_len, output_size=total)
            seg_first = torch.cumsum(seg_len, 0) - seg_len
            tt = torch.arange(total, device=dev) - seg_first[seg]
            out[dst0[seg] + tt] = recv[src0[seg] + tt]
        return n_hits, hit_off, out


def pattern_tensors(chars: np.ndarray, pat_off: np.ndarray, device):
    """numpy (uint16 chars, uint64 offsets) -> tensors the engines take (int16 / int64 views)."""
    return (torch.from_numpy(np.ascontiguousarray(chars, dtype=np.uint16).view(np.int16)).to(device),
            torch.from_numpy(np.ascontiguousarray(pat_off, dtype=np.uint64).view(np.int64)).to(device))
