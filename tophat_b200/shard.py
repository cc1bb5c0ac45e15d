"""Read sharding across ranks and the merge of per-rank result sets (host logic of the multi-GPU path).

The reference partitions reads into contiguous read-id ranges, one per thread, and unions the
per-thread sets afterwards (segment_juncs.cpp:4776-4825, 4911-4922).  We do the same across GPUs:
rank r owns a contiguous bundle range of every batch; junction / deletion sets are unioned, and an
insertion keeps the sequence of the earliest bundle in the single-process order (order_base + bundle
index), which is what the std::set first-insert-wins rule of insertions.h:52-67 yields at -p1.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import synth


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of n items for `rank` of `world`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    q, r = divmod(n, world)
    b = rank * q + min(rank, r)
    return b, b + q + (1 if rank < r else 0)


def shard_batch(b: synth.PackedBatch, rank: int, world: int) -> synth.PackedBatch:
    """The bundles [begin, end) of `b` with re-based hit offsets; order_base keeps the global position."""
    lo, hi = shard_range(b.n_bundles, rank, world)
    nb = b.n_bundles
    h0 = int(b.bundles["hit_begin"][lo]) if lo < nb else int(b.hits.shape[0])
    h1 = int(b.bundles["hit_begin"][hi]) if hi < nb else int(b.hits.shape[0])
    p0 = int(b.bundles["partner_begin"][lo]) if lo < nb else int(b.partner_hits.shape[0])
    p1 = int(b.bundles["partner_begin"][hi]) if hi < nb else int(b.partner_hits.shape[0])
    bundles = b.bundles[lo:hi].copy()
    bundles["hit_begin"] -= h0
    bundles["partner_begin"] -= p0
    return synth.PackedBatch(b.n_segs, b.read_words, bundles, np.ascontiguousarray(b.seg_count[lo:hi]),
                             np.ascontiguousarray(b.reads[lo:hi]), np.ascontiguousarray(b.hits[h0:h1]),
                             np.ascontiguousarray(b.partner_hits[p0:p1]), b.order_base + lo)


def shard_join_batch(b: synth.PackedJoinBatch, rank: int, world: int) -> synth.PackedJoinBatch:
    """The reads [begin, end) of a join batch with re-based hit / CIGAR offsets.  long_spanning_reads needs no exchange: the
    junction / indel / fusion sets are inputs every rank holds, every read is joined on its own (long_spanning_reads.cpp:3051-3140)."""
    lo, hi = shard_range(b.n_bundles, rank, world)
    nb = b.n_bundles
    h0 = int(b.bundles["hit_begin"][lo]) if lo < nb else int(b.hits.shape[0])
    h1 = int(b.bundles["hit_begin"][hi]) if hi < nb else int(b.hits.shape[0])
    e0 = int(b.bundles["ops_begin"][lo]) if lo < nb else int(b.ops_ext.shape[0])
    e1 = int(b.bundles["ops_begin"][hi]) if hi < nb else int(b.ops_ext.shape[0])
    bundles = b.bundles[lo:hi].copy()
    bundles["hit_begin"] -= h0
    bundles["ops_begin"] -= e0
    return synth.PackedJoinBatch(b.n_segs, b.read_words, bundles, np.ascontiguousarray(b.seg_count[lo:hi]), np.ascontiguousarray(b.reads[lo:hi]),
                                 np.ascontiguousarray(b.hits[h0:h1]), np.ascontiguousarray(b.ops_ext[e0:e1]))


def _uniq_sorted(rec: np.ndarray, fields: Sequence[str]) -> np.ndarray:
    if rec.size == 0:
        return rec
    order = np.lexsort(tuple(rec[f] for f in reversed(fields)))
    rec = rec[order]
    keep = np.ones(rec.size, dtype=bool)
    same = np.ones(rec.size - 1, dtype=bool)
    for f in fields:
        same &= rec[f][1:] == rec[f][:-1]
    keep[1:] = ~same
    return rec[keep]


def merge_junction_sets(parts: List[np.ndarray], cap: int = 10_000_000) -> np.ndarray:
    """Union in Junction order (junctions.h:39-57), capped like the reference's set (segment_juncs.cpp:58)."""
    allj = np.concatenate(parts) if parts else np.zeros(0, dtype=synth.JUNCTION_DTYPE)
    return _uniq_sorted(allj, ("ref_id", "left", "right", "antisense"))[:cap]


def merge_insertion_sets(parts: List[np.ndarray], orders: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Union by (ref, left, length); the record with the smallest processing order wins."""
    rec = np.concatenate(parts) if parts else np.zeros(0, dtype=synth.INSERTION_DTYPE)
    od = np.concatenate(orders) if orders else np.zeros(0, dtype=np.uint64)
    if rec.size == 0:
        return rec, od
    idx = np.lexsort((od, rec["len"], rec["left"], rec["ref_id"]))
    rec, od = rec[idx], od[idx]
    keep = np.ones(rec.size, dtype=bool)
    keep[1:] = ~((rec["ref_id"][1:] == rec["ref_id"][:-1]) & (rec["left"][1:] == rec["left"][:-1]) &
                 (rec["len"][1:] == rec["len"][:-1]))
    return rec[keep], od[keep]
