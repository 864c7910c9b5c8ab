"""Collections sharded by sequence across the GPUs of one box.

One process per GPU (torch.distributed; NCCL on the GPUs, gloo in the CPU tests). The collection is cut into
`world` contiguous groups of sequences balanced by length; rank r holds the E2FM index the reference built
for group r. A batch is answered in one exchange step:

    broadcast(patterns from rank 0)  ->  local search on every shard (CUDA, e2fm_b200.capi)
    all_reduce(SUM, counts)          ->  EFMIndex::countOccurrences over the whole collection
    all_reduce(SUM, occurrences per pattern) + all_gather(int32 occurrences per pattern per shard)
                                     ->  merged CSR offsets, and where each shard's occurrences go inside every pattern
    all_reduce(SUM, zero-filled merged arrays written at disjoint places)
                                     ->  per-pattern concatenation in shard order, on every rank

Exact patterns never span a terminator and every sequence starts k-aligned (EFMCollection.cpp:218-225), so an
occurrence has the same (sequence, offset) whether its sequence is indexed alone, in a shard or in the whole
collection; adding the shard's text base gives the very position the unsharded reference index reports, and
because shards are contiguous sequence ranges the concatenation is already sorted (EFMIndex.cpp:2291).

The merge below is device-agnostic tensor code: the CPU tests drive it with world_size-2 gloo groups.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


def padded_length(seq_len: int, k: int) -> int:
    """bytes a sequence occupies in the indexed text: padded to a multiple of k plus a k-long terminator
    (EFMCollection.cpp:187-233)"""
    return (seq_len + k - 1) // k * k + k


def plan_shards(seq_lens, world: int, k: int):
    """Contiguous groups balanced by length -> list of (first_seq, n_seqs, text_base) per rank.

    A group is closed after the sequence that brings the running total to its share, at most one group per sequence, and never
    so late that a later rank would be left without a sequence: with at least `world` sequences every rank gets one or more
    (an empty shard has no index to build), however skewed the lengths are."""
    lens = [int(x) for x in seq_lens]
    n = len(lens)
    total = sum(lens)
    bounds = [0]
    acc = 0
    for i, L in enumerate(lens):
        acc += L
        g = len(bounds)                       # the group being filled is g-1; groups g..world-1 still need sequences
        if g >= world:
            break
        left_after = n - (i + 1)
        must = left_after == world - g        # closing later would starve a later rank
        may = left_after >= world - g         # closing now leaves one sequence or more for every later rank
        if must or (may and acc * world >= total * g):
            bounds.append(i + 1)
    while len(bounds) < world:
        bounds.append(n)
    bounds.append(n)
    out = []
    base = 0
    for r in range(world):
        a, b = bounds[r], bounds[r + 1]
        out.append((a, b - a, base))
        base += sum(padded_length(L, k) for L in lens[a:b])
    return out


@dataclass
class LocalResult:
    """what one shard found, as tensors on the rank's device"""
    counts: torch.Tensor        # int64 [n]
    pos_offsets: torch.Tensor   # int64 [n+1]
    positions: torch.Tensor     # int64 [total]  text positions inside the shard's index
    seq: torch.Tensor           # int64 [total]  sequence index inside the shard
    off: torch.Tensor           # int64 [total]  offset inside the sequence


@dataclass
class MergedResult:
    counts: torch.Tensor        # int64 [n]   whole-collection counts (on every rank)
    pos_offsets: torch.Tensor   # int64 [n+1]
    positions: torch.Tensor     # int64 [total] positions in the unsharded collection text, ascending per pattern
    seq: torch.Tensor           # int64 [total] collection-wide sequence index
    off: torch.Tensor           # int64 [total]


def broadcast_patterns(pats: torch.Tensor | None, shape, device, group=None) -> torch.Tensor:
    """rank 0 holds the (n, L) uint8 batch; every rank returns it on `device`"""
    rank = dist.get_rank(group)
    t = pats.to(device) if rank == 0 else torch.empty(shape, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0, group=group)
    return t


def merge_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    out = counts.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def merge_results(local: LocalResult, first_seq: int, text_base: int, group=None, locate: bool = True) -> MergedResult:
    """The exchange step. Every rank ends up with the whole merged result."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local.counts.device
    n = local.counts.numel()
    counts = merge_counts(local.counts, group)
    if not locate:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return MergedResult(counts, torch.zeros(n + 1, dtype=torch.int64, device=dev), z, z, z)
    # Per pattern, the merged list is rank 0's occurrences, then rank 1's, ...: every rank computes where ITS occurrences go and
    # writes them into a zero-filled array of the merged size; destinations are disjoint, so one all-reduce(SUM) assembles the
    # whole result on every rank (an all-gather-v without padding or a receive-side scatter).
    mine = (local.pos_offsets[1:] - local.pos_offsets[:-1]).contiguous()            # my occurrences per pattern
    per_pattern = mine.clone()
    dist.all_reduce(per_pattern, op=dist.ReduceOp.SUM, group=group)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(per_pattern, dim=0, out=offsets[1:])
    # occurrences of the ranks before me, per pattern (int32 on the wire: a pattern has < 2^31 occurrences in one shard)
    per_rank = torch.empty(world * n, dtype=torch.int32, device=dev)                # flat: gloo only takes 1-D outputs
    dist.all_gather_into_tensor(per_rank, mine.to(torch.int32), group=group)
    before = per_rank.view(world, n)[:rank].sum(dim=0, dtype=torch.int64) if rank > 0 else torch.zeros(n, dtype=torch.int64, device=dev)
    total = int(offsets[-1].item()) if n else 0
    out = torch.zeros((3, total), dtype=torch.int64, device=dev)
    t = local.positions.numel()
    if t:
        pat = torch.repeat_interleave(torch.arange(n, device=dev), mine, output_size=t)
        dest = offsets[:-1][pat] + before[pat] + (torch.arange(t, device=dev) - local.pos_offsets[:-1][pat])
        out[0, dest] = local.positions + text_base
        # sequence index -1 (no terminator beyond the position, EFMCollection.cpp:124-133) stays -1
        out[1, dest] = torch.where(local.seq < 0, local.seq, local.seq + first_seq)
        out[2, dest] = local.off
    if total:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return MergedResult(counts, offsets, out[0], out[1], out[2])


# ------------------------------------------------------------------ GPU side


class _DevArray:
    """zero-copy view of a device pointer for torch.as_tensor (CUDA array interface v2)"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def local_result_from_device(res, device) -> LocalResult:
    """Wrap an e2fm_b200.capi.Result (device-resident) as int64 tensors without leaving the GPU."""
    p = res.device_ptrs()
    n, t = res.n, res.total

    def view(ptr, count, typestr, dtype):
        if count == 0 or not ptr:
            return torch.zeros(0, dtype=dtype, device=device)
        return torch.as_tensor(_DevArray(ptr, count, typestr), device=device)

    counts = view(p["counts"], n, "<i8", torch.int64).clone()
    if not res.locate:
        z = torch.zeros(0, dtype=torch.int64, device=device)
        return LocalResult(counts, torch.zeros(n + 1, dtype=torch.int64, device=device), z, z, z)
    offs = view(p["pos_offsets"], n + 1, "<i8", torch.int64).clone()
    pos = view(p["positions"], t, "<i8", torch.int64).clone()
    seq = view(p["seq"], t, "<i4", torch.int32).to(torch.int64)      # 0xFFFFFFFF (no sequence) becomes -1
    off = view(p["off"], t, "<i8", torch.int64).clone()
    return LocalResult(counts, offs, pos, seq, off)


class ShardedCollection:
    """rank-local handle: this rank's shard index on its GPU + the collective merge"""

    def __init__(self, index_bytes: bytes, key64: bytes, device_index: int, first_seq: int, text_base: int, group=None):
        from . import capi
        self.capi = capi
        self.ix = capi.DeviceIndex(index_bytes, key64, device_index)
        self.device = torch.device("cuda", device_index)
        self.first_seq = first_seq
        self.text_base = text_base
        self.group = group
        self.last_phase_ms = {}

    def search_batch(self, pats_rank0, shape, locate: bool, pieces: int | None = None) -> MergedResult:
        """pats_rank0: (n, L) uint8 tensor on rank 0, ideally pinned (ignored elsewhere); shape = (n, L) known to every rank.

        The batch is cut into pieces: rank 0 queues every piece's host->device copy up front on a side stream, and each piece is
        broadcast, searched and merged as soon as it has arrived, so the PCIe transfer of the later pieces overlaps the NCCL,
        kernel and merge time of the earlier ones."""
        import time
        n, L = shape
        rank = dist.get_rank(self.group)
        if pieces is None:
            # worth it only when the PCIe transfer dominates (measured: 96 MB batches lose to the per-piece overheads)
            pieces = 4 if n * L >= (256 << 20) else 1
        bounds = [n * i // pieces for i in range(pieces + 1)]
        t0 = time.perf_counter()
        d_pats = torch.empty((n, L), dtype=torch.uint8, device=self.device)
        events = []
        if rank == 0:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for i in range(pieces):
                    a, b = bounds[i], bounds[i + 1]
                    d_pats[a:b].copy_(pats_rank0[a:b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    events.append(ev)
        ph = {"broadcast": 0.0, "search_device": 0.0, "search_wall": 0.0, "merge": 0.0}
        parts = []
        for i in range(pieces):
            a, b = bounds[i], bounds[i + 1]
            if b == a:
                continue
            t1 = time.perf_counter()
            if rank == 0:
                torch.cuda.current_stream(self.device).wait_event(events[i])
            dist.broadcast(d_pats[a:b], src=0, group=self.group)
            # (stream-level syncs only: a device-wide one would also wait for the copies of the later pieces)
            torch.cuda.current_stream(self.device).synchronize()   # the library searches on its own stream
            t2 = time.perf_counter()
            res = self.ix.search_device_fixed(d_pats.data_ptr() + a * L, b - a, L, locate)
            ph["search_device"] += self.ix.last_batch_ms()["total"]
            local = local_result_from_device(res, self.device)
            torch.cuda.current_stream(self.device).synchronize()
            res.free()
            t3 = time.perf_counter()
            parts.append(merge_results(local, self.first_seq, self.text_base, self.group, locate))
            torch.cuda.current_stream(self.device).synchronize()
            t4 = time.perf_counter()
            ph["broadcast"] += (t2 - t1) * 1e3
            ph["search_wall"] += (t3 - t2) * 1e3
            ph["merge"] += (t4 - t3) * 1e3
        ph["wall"] = (time.perf_counter() - t0) * 1e3
        self.last_phase_ms = ph
        if len(parts) == 1:
            return parts[0]
        if not parts:
            z = torch.zeros(0, dtype=torch.int64, device=self.device)
            return MergedResult(z, torch.zeros(1, dtype=torch.int64, device=self.device), z, z, z)
        # pieces are contiguous pattern ranges: concatenate, shifting the CSR offsets
        offs = [parts[0].pos_offsets]
        shift = int(parts[0].positions.numel())
        for p in parts[1:]:
            offs.append(p.pos_offsets[1:] + shift)
            shift += int(p.positions.numel())
        return MergedResult(torch.cat([p.counts for p in parts]), torch.cat(offs), torch.cat([p.positions for p in parts]),
                            torch.cat([p.seq for p in parts]), torch.cat([p.off for p in parts]))

    def close(self):
        self.ix.close()
