"""world_size-2 gloo test of the multi-GPU host logic: read sharding + merge of per-rank sets equals the
single-process result (the reference's per-thread set union, segment_juncs.cpp:4911-4922)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from tophat_b200 import shard, synth
from oracle import pyoracle


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl, P, _ = helpers.load_golden(name)
        batches = [shard.shard_batch(b, rank, world) for b in helpers.pack_both(wl)]
        res, _ = pyoracle.segjuncs(P, wl.ref, batches)
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(j=res.junctions, d=res.deletions, i=res.insertions, o=res.insertion_order))
        j = shard.merge_junction_sets([g["j"] for g in gathered])
        d = shard.merge_junction_sets([g["d"] for g in gathered])
        i, _ = shard.merge_insertion_sets([g["i"] for g in gathered], [g["o"] for g in gathered])
        if rank == 0:
            q.put((j, d, i))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["indel_heavy", "splice_2contig"])
def test_two_rank_shard_equals_single(name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs: p.start()
    j, d, i = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    wl, P, want = helpers.load_golden(name)
    full, _ = pyoracle.segjuncs(P, wl.ref, helpers.pack_both(wl))
    assert (j == full.junctions).all() and j.shape == full.junctions.shape
    assert (d == full.deletions).all() and d.shape == full.deletions.shape
    assert (i == full.insertions).all() and i.shape == full.insertions.shape


def test_shard_range_partition():
    for n in (0, 1, 7, 100):
        for w in (1, 2, 3, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1


def test_insertion_first_wins_across_shards():
    """A right-mate insertion on rank 0 must lose against a left-mate insertion on rank 1 (left mates are
    processed first in the reference, segment_juncs.cpp:4752 / 4831)."""
    a = np.zeros(1, dtype=synth.INSERTION_DTYPE); a["ref_id"] = 1; a["left"] = 10; a["len"] = 2; a["seq"] = b"AA"
    b = a.copy(); b["seq"] = b"CC"
    rec, od = shard.merge_insertion_sets([a, b], [np.array([9000 << 12], dtype=np.uint64), np.array([5 << 12], dtype=np.uint64)])
    assert rec.size == 1 and rec["seq"][0] == b"CC"


def test_join_batch_shards_cover_the_batch():
    """shard_join_batch: contiguous read ranges with re-based hit / CIGAR offsets; together they are the whole batch."""
    import numpy as np
    from tophat_b200 import shard, synth
    wl = synth.generate(synth.SynthConfig(contig_lens=(200_000,), n_pairs=1500, seed=77, indel_prob=0.2, keep_candidates=True))
    j = np.zeros(0, dtype=synth.JUNCTION_DTYPE)
    jb = synth.pack_join_side(wl, wl.left, j)
    for world in (1, 2, 3, 8):
        parts = [shard.shard_join_batch(jb, r, world) for r in range(world)]
        assert sum(p.n_bundles for p in parts) == jb.n_bundles
        assert (np.concatenate([p.hits for p in parts]) == jb.hits).all() and (np.concatenate([p.ops_ext for p in parts]) == jb.ops_ext).all()
        assert (np.concatenate([p.reads for p in parts]) == jb.reads).all() and (np.concatenate([p.seg_count for p in parts]) == jb.seg_count).all()
        for p in parts:
            if p.n_bundles:
                assert int(p.bundles["hit_begin"][0]) == 0 and int(p.bundles["ops_begin"][0]) == 0
                cnt = p.seg_count.reshape(p.n_bundles, -1).sum(axis=1)
                assert (p.bundles["hit_begin"][1:] == np.cumsum(cnt)[:-1]).all()
