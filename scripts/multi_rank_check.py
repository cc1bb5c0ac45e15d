"""Multi-GPU parity check (one process per GPU, launched by torchrun): every rank runs segment_juncs -- with
--fusion-search -- on its contiguous shard of the bundles, the ranks exchange their junction / deletion / insertion /
fusion records with thb_segjuncs_allgather (NCCL), and every rank must end up with exactly the sets the CPU oracle
computes for the whole input (BASELINE configs[4]: fusion inter/intra set on 2 GPUs; configs[2]: read-shard + all-gather).
Then the join runs collective-free on each rank's shard of the reads and the per-rank record counts are summed.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_rank_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np
import torch
import torch.distributed as dist

from tophat_b200 import capi, shard, synth
from oracle import pyoracle
import helpers


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, kw, over in [
        ("fusion inter/intra", dict(contig_lens=(600_000, 250_000, 90_000), n_pairs=12_000, seed=611, indel_prob=0.2, fusion_frac=0.15),
         dict(fusion_search=1, fusion_min_dist=30000)),
        ("splice + indel", dict(contig_lens=(900_000, 300_000), n_pairs=20_000, seed=612, indel_prob=0.4), {}),
    ]:
        wl = synth.generate(synth.SynthConfig(keep_candidates=True, **kw))
        o = dict(inner_dist_mean=50, inner_dist_std_dev=20); o.update(over)
        P = capi.default_params(**o)
        batches = helpers.pack_both(wl, P)
        ctx = capi.Context(local); ctx.ref_upload(wl.ref)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        ctx.segjuncs_begin(P)
        for b in batches:
            ctx.segjuncs_submit(shard.shard_batch(b, rank, world))
        ctx.segjuncs_allgather()
        got = ctx.segjuncs_finish()
        want, _ = pyoracle.segjuncs(P, wl.ref, batches)
        try:
            helpers.assert_same_results(got, want, "%s, rank %d of %d" % (name, rank, world))
            print("[rank %d] %s: %d junctions, %d deletions, %d insertions, %d fusions == oracle" % (
                rank, name, len(got.junctions), len(got.deletions), len(got.insertions), len(got.fusions)), flush=True)
        except AssertionError as e:
            ok = False
            print("[rank %d] MISMATCH %s" % (rank, e), flush=True)
        # stage 2 on the same ranks (BASELINE configs[4]: closures + fusions path on 2 GPUs): every rank joins its shard of the reads
        # against the full sets; the shards' records together must be the records one context produces for all reads
        juncs, ins = capi.join_sets_from_results(got)
        ctx.join_begin(P, juncs, ins)
        if o.get("fusion_search"):
            ctx.join_set_fusions(got.fusions)
        mine, whole = [], []
        for side in (wl.left, wl.right):
            jb = synth.pack_join_side(wl, side, got.junctions)
            lo, _ = shard.shard_range(jb.n_bundles, rank, world)
            for r in ctx.join_submit(shard.shard_join_batch(jb, rank, world)):
                mine.append((int(jb.bundles["read_id"][lo + int(r["bundle"])]), int(r["ref_id"]), int(r["left"]), int(r["flags"]), int(r["mismatches"]), int(r["edit_dist"]),
                             tuple(int(x) for x in r["ops"][:r["n_ops"]]), int(r["ops"][26]) if (r["flags"] & 0x10) else 0))
            if rank == 0:
                for r in ctx.join_submit(jb):
                    whole.append((int(jb.bundles["read_id"][int(r["bundle"])]), int(r["ref_id"]), int(r["left"]), int(r["flags"]), int(r["mismatches"]), int(r["edit_dist"]),
                                  tuple(int(x) for x in r["ops"][:r["n_ops"]]), int(r["ops"][26]) if (r["flags"] & 0x10) else 0))
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            union = sorted(x for p_ in parts for x in p_)
            n_fused = sum(1 for x in union if x[3] & 0x10)
            if union == sorted(whole) and (not o.get("fusion_search") or n_fused > 20):
                print("[rank 0] %s: join on %d shards = single context: %d alignments, %d across a fusion" % (name, world, len(union), n_fused), flush=True)
            else:
                ok = False
                print("[rank 0] MISMATCH %s: sharded join %d records vs %d (%d fused)" % (name, len(union), len(whole), n_fused), flush=True)
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTI-RANK OK (world %d)" % world, flush=True)


if __name__ == "__main__":
    main()
