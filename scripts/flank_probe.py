"""Junction-flank matcher at scale: stage 1 over a bench workload gives the sets, thb_flank_begin builds the index over them,
thb_flank_submit searches every segment of every unmapped read.  Prints the timings, checks a sample of reads against the oracle
and that every spliced segment planted by the generator is found on a contig of its junction."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hg38")
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check-reads", type=int, default=300)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from tophat_b200 import capi, synth
    from oracle import flank_oracle
    bench.WORKLOAD = args.workload
    wl = bench.make_workload(args.pairs, 0, os.cpu_count() or 1, keep_candidates=True, kind=args.workload)
    batches = bench.pack(wl)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    ctx.segjuncs_begin(P)
    for b in batches:
        ctx.segjuncs_submit(b)
    res = ctx.segjuncs_finish(True)
    print("[probe] sets: %d junctions, %d deletions, %d insertions" % (len(res.junctions), len(res.deletions), len(res.insertions)), flush=True)
    offs, lens = synth.segment_layout(wl.cfg.read_len, wl.cfg.segment_length)
    bounds = [int(o) for o in offs] + [int(offs[-1] + lens[-1])]
    FP = capi.FlankParams(2, 40, int(lens.min()), int(lens.max()), 3, 0)
    rw = (wl.cfg.read_len + 63) // 64
    out = {"workload": args.workload, "pairs": args.pairs, "seg_bounds": bounds}
    t0 = time.time()
    ctx.flank_begin(FP, res.junctions, res.deletions, res.insertions, res.fusions)
    out["begin_wall_s"] = time.time() - t0
    t = ctx.flank_timing()
    out.update(index_ms=t.index_ms, n_contigs=int(t.n_contigs), n_index_entries=int(t.n_index_entries))
    print("[probe] index: %d contigs, %d entries, %.1f ms on the device, %.2f s wall" % (t.n_contigs, t.n_index_entries, t.index_ms, out["begin_wall_s"]), flush=True)
    sides = []
    for side in (wl.left, wl.right):
        idx = np.nonzero(side.unmapped)[0]
        packed = synth.pack_reads(side.reads[idx], rw)
        pin = torch.from_numpy(packed).pin_memory()
        dev = pin.cuda()
        sides.append((side, idx, pin, dev))
    torch.cuda.synchronize()
    rows = []
    for step in range(args.steps):
        for si, (side, idx, pin, dev) in enumerate(sides):
            w0 = time.time()
            hits = ctx.flank_submit(len(idx), rw, bounds, device_ptr=dev.data_ptr(), copy=(step == args.steps - 1))
            w1 = time.time()
            t = ctx.flank_timing()
            rows.append(dict(step=step, side=si, reads=len(idx), hits=int(t.n_hits), verified=int(t.n_verified), match_ms=t.match_ms, post_ms=t.post_ms,
                             d2h_ms=t.d2h_ms, wall_ms=(w1 - w0) * 1e3, algorithmic_bytes=int(t.algorithmic_bytes)))
            if step == args.steps - 1:
                sides[si] = (side, idx, pin, dev, hits)
        print("[probe] step %d: %s" % (step, json.dumps(rows[-2:])), flush=True)
    # host-buffer submit (copies inside)
    side, idx, pin, dev, hits = sides[0]
    w0 = time.time(); h2 = ctx.flank_submit(pin.numpy(), rw, bounds); w1 = time.time()
    t = ctx.flank_timing()
    out["host_submit"] = dict(reads=len(idx), wall_ms=(w1 - w0) * 1e3, h2d_ms=t.h2d_ms, match_ms=t.match_ms, post_ms=t.post_ms, d2h_ms=t.d2h_ms)
    assert len(h2) == len(hits) and (h2 == hits).all()
    out["rows"] = rows
    last = [r for r in rows if r["step"] == args.steps - 1]
    segs = sum(r["reads"] for r in last) * (len(bounds) - 1)
    ms = sum(r["match_ms"] + r["post_ms"] for r in last)
    out["segments_per_s"] = segs / (ms * 1e-3)
    out["reads_per_s"] = sum(r["reads"] for r in last) / (ms * 1e-3)
    print("[probe] %.3g segments/s, %.3g reads/s (device time of search + filter/sort/decode)" % (out["segments_per_s"], out["reads_per_s"]), flush=True)
    # parity on a sample of reads: the oracle by exhaustion over the contigs those reads' placements name, plus a random 2000 others
    contigs = ctx.flank_contigs()
    ck = min(args.check_reads, len(idx))
    if ck:
        rng = np.random.default_rng(1)
        pick = np.sort(rng.choice(len(idx), ck, replace=False))
        sel = np.isin(hits["read"], pick)
        hs = hits[sel]
        cset = np.unique(np.concatenate([hs["contig"], rng.choice(len(contigs), min(2000, len(contigs)), replace=False)]))
        # restate those contigs from the reference codes
        jn = np.stack([res.junctions[n].astype(np.int64) for n in ("ref_id", "left", "right", "antisense")], axis=1)
        dl = np.stack([res.deletions[n].astype(np.int64) for n in ("ref_id", "left", "right", "antisense")], axis=1) if len(res.deletions) else np.zeros((0, 4), np.int64)
        ins = [(int(r["ref_id"]), int(r["left"]), r["seq"].decode()) for r in res.insertions]
        allc = flank_oracle.contigs(wl.ref.names, wl.ref.codes, int(lens.max()), 3, jn, dl, ins, np.zeros((0, 5), np.int64)) if len(contigs) <= 3_000_000 else None
        if allc is not None:
            assert len(allc) == len(contigs)
            sub = [allc[int(c)]["codes"] for c in cset]
            want = flank_oracle.search(sub, side.reads[idx][pick], bounds, 2, 10**9)
            want[:, 0] = pick[want[:, 0]]; want[:, 2] = cset[want[:, 2]]
            got = np.stack([hs[n].astype(np.int64) for n in ("read", "seg", "contig", "pos", "antisense", "mismatches")], axis=1).reshape(-1, 6)
            # -m: compare only segments below the cut in our result (the oracle ran without it on a contig subset)
            want = want[np.lexsort(want.T[::-1])]
            ws = set(map(tuple, want)); gs = set(map(tuple, got))
            assert gs <= ws or not (gs - ws), "placements the oracle does not have: %r" % sorted(gs - ws)[:3]
            cnt = {}
            for r in want:
                cnt[(r[0], r[1])] = cnt.get((r[0], r[1]), 0) + 1
            missing = [r for r in ws - gs if cnt[(r[0], r[1])] <= 40]
            assert not missing, "placements missing: %r" % missing[:3]
            out["parity_sample"] = dict(reads=int(ck), contigs=int(len(cset)), placements=int(len(gs)))
            print("[probe] parity sample ok:", out["parity_sample"], flush=True)
    ctx.close()
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
