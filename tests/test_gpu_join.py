"""C-ABI level parity of the chain join: packed join batches (vectorised packer used by bench.py) through
thb_join_begin / thb_join_submit against the records the reference's long_spanning_reads binary writes for the same
reads, segment hits and junction sets."""
import os
import tempfile

import numpy as np
import pytest

import helpers
from tophat_b200 import capi, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu

OPCH = {1: "M", 3: "I", 5: "D", 11: "N", 13: "S"}


def joined_to_keys(joined, batch, P):
    """Applies the worker's sort/unique/filters (long_spanning_reads.cpp:2805-2813) and returns (read_id, ref_id, pos, cigar, flag, NM)."""
    out = set()
    for r in joined:
        ops = [(int(o) & 15, int(o) >> 4) for o in r["ops"][:r["n_ops"]]]
        gap = sum(l for c, l in ops if c in (3, 5))
        if r["mismatches"] > P.read_mismatches or gap > P.read_gap_length or r["edit_dist"] > P.read_edit_dist:
            continue
        cig = "".join("%d%s" % (l, OPCH[c]) for c, l in ops)
        out.add((int(batch.bundles["read_id"][r["bundle"]]), int(r["ref_id"]), int(r["left"]), cig, 16 if r["flags"] & 1 else 0,
                 int(r["mismatches"]) + gap))
    return out


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("kw,chunk", [
    (dict(contig_lens=(300_000, 100_000), n_pairs=3000, seed=501), None),
    (dict(contig_lens=(300_000,), n_pairs=3000, seed=502, indel_prob=0.5), None),
    (dict(contig_lens=(300_000, 50_000), n_pairs=3000, seed=503, indel_prob=0.3), 512),      # 6 pipeline chunks per submit
])
def test_join_capi_matches_reference_records(kw, chunk, monkeypatch):
    if chunk:
        monkeypatch.setenv("THB_JOIN_CHUNK_READS", str(chunk))
    wl = synth.generate(synth.SynthConfig(keep_truth=True, **kw))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    res, _ = helpers.gpu_segjuncs(P, wl.ref, helpers.pack_both(wl), ctx)
    juncs, ins = capi.join_sets_from_results(res)
    ctx.join_begin(P, juncs, ins)
    names = wl.ref.names
    with tempfile.TemporaryDirectory() as td:
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg)
        assert open(outs["juncs"]).read() == pyoracle.format_juncs(res.junctions, names)
        jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg)
        for sname, side in (("left", wl.left), ("right", wl.right)):
            batch = synth.pack_join_side(wl, side, res.junctions)
            joined = ctx.join_submit(batch)
            got = joined_to_keys(joined, batch, P)
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=sname, tag=".ref")
            refs, recs = pyoracle.read_bam(ref_bam)
            want = set((int(r[0]), names.index(r[2]) + 1, r[3], r[5], r[1], r[11]["NM"]) for r in recs)
            assert got == want, "%s: %d vs %d alignments; only ours %r; only reference %r" % (
                sname, len(got), len(want), sorted(got - want)[:3], sorted(want - got)[:3])
            assert len(want) > 200
            t = ctx.join_timing()
            assert t.n_joined >= len(got) and t.launches >= 1
    ctx.close()


def test_join_requires_begin_and_sorted_sets():
    wl, P, _ = helpers.load_golden("splice_2contig")
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    j = np.zeros(2, dtype=synth.JUNCTION_DTYPE); j["ref_id"] = 1; j["left"] = [50, 10]; j["right"] = [90, 40]
    with pytest.raises(capi.ThbError):
        ctx.join_begin(P, j, np.zeros(0, dtype=synth.INSERTION_DTYPE))          # not sorted
    ctx.join_begin(capi.default_params(fusion_search=1), j[::-1].copy(), np.zeros(0, dtype=synth.INSERTION_DTYPE))     # --fusion-search is on the GPU path
    f = np.zeros(2, dtype=synth.FUSION_DTYPE); f["ref_id1"] = 1; f["ref_id2"] = 1; f["left"] = [500, 100]; f["right"] = [900, 700]; f["dir"] = 7
    with pytest.raises(capi.ThbError):
        ctx.join_set_fusions(f)                                                                                       # not sorted
    ctx.join_set_fusions(f[::-1].copy())
    ctx.close()


def _recount_mismatches(wl, batch, joined):
    """Independent numpy re-read of every joined alignment: walks the CIGAR over the genome codes and the (oriented) read and
    counts differing bases the way check_editdist_consistency does (bwt_map.cpp:2349-2465)."""
    comp = np.array([3, 2, 1, 0, 4], dtype=np.uint8)
    # unpack the batch's read planes back to codes
    rw = batch.read_words
    planes = batch.reads.reshape(batch.n_bundles, 3, rw)
    out_mm, out_len, out_nn = np.zeros(len(joined), dtype=np.int64), np.zeros(len(joined), dtype=np.int64), np.zeros(len(joined), dtype=np.int64)
    for i, r in enumerate(joined):
        b = int(r["bundle"]); L = int(batch.bundles["read_len"][b])
        bits = lambda w: ((planes[b, w][:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).reshape(-1)[:L].astype(np.uint8)
        codes = bits(0) | (bits(1) << 1); codes[bits(2) == 1] = 4
        if r["flags"] & 1:
            codes = comp[codes[::-1]]
        ref = wl.ref.codes[int(r["ref_id"]) - 1]
        pos_ref, pos_seq, mm, nn = int(r["left"]), 0, 0, 0
        for o in r["ops"][:r["n_ops"]]:
            c, l = int(o) & 15, int(o) >> 4
            if c == 1:
                g = ref[pos_ref:pos_ref + l]; s = codes[pos_seq:pos_seq + l]
                mm += int(((g != s) & ~((g > 3) & (s > 3))).sum()) if len(g) == l else 10**6
                nn += int(((g > 3) & (s > 3)).sum()) if len(g) == l else 0
                pos_ref += l; pos_seq += l
            elif c == 3:
                pos_seq += l
            elif c in (5, 11):
                pos_ref += l
        out_mm[i] = mm; out_len[i] = pos_seq; out_nn[i] = nn
    return out_mm, out_len, out_nn


def test_join_properties_at_scale(monkeypatch):
    """100k pairs through both stages: size-independent properties of the joined alignments -- the CIGAR covers the whole read,
    starts and ends with a match, every REF_SKIP is a junction of the set handed to thb_join_begin (after the +-4 bp boundary
    adjustment the skip itself is an exact set member), the mismatch field equals an independent recount for a sample of the
    records, and the record set does not depend on how the batch is cut into pipeline chunks or shards."""
    wl = synth.generate(synth.SynthConfig(contig_lens=(6_000_000, 2_000_000), n_pairs=100_000, seed=521, indel_prob=0.2, keep_candidates=True), workers=4)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    ctx = capi.Context(0); ctx.ref_upload(wl.ref)
    res, _ = helpers.gpu_segjuncs(P, wl.ref, helpers.pack_both(wl), ctx)
    juncs, ins = capi.join_sets_from_results(res)
    ctx.join_begin(P, juncs, ins)
    jset = set((int(j["ref_id"]), int(j["left"]), int(j["right"])) for j in juncs)
    batch = synth.pack_join_side(wl, wl.left, res.junctions)
    joined = ctx.join_submit(batch)
    assert len(joined) > 40_000
    def key(recs, bundle_ids):
        return sorted((int(bundle_ids[r["bundle"]]), int(r["ref_id"]), int(r["left"]), int(r["flags"]), int(r["mismatches"]), int(r["edit_dist"]),
                       tuple(int(o) for o in r["ops"][:r["n_ops"]])) for r in recs)
    whole = key(joined, batch.bundles["read_id"])
    n_skip = 0
    for r in joined:
        ops = [(int(o) & 15, int(o) >> 4) for o in r["ops"][:r["n_ops"]]]
        assert ops[0][0] == 1 and ops[-1][0] == 1
        assert sum(l for c, l in ops if c in (1, 3, 13)) == int(batch.bundles["read_len"][r["bundle"]])
        pos = int(r["left"])
        for c, l in ops:
            if c == 11:
                assert (int(r["ref_id"]), pos - 1, pos + l) in jset, "REF_SKIP that is not in the junction set"
                n_skip += 1
            if c in (1, 5, 11):
                pos += l
    assert n_skip > 20_000
    sample = joined[:: max(1, len(joined) // 3000)]
    mm, ln, nn = _recount_mismatches(wl, batch, sample)          # positions where read and genome are both N may count or not
    assert ((mm == sample["mismatches"]) | (mm + nn == sample["mismatches"])).all() and (ln == batch.bundles["read_len"][sample["bundle"]]).all()
    # chunk pipeline: 13 chunks instead of 1
    monkeypatch.setenv("THB_JOIN_CHUNK_READS", "4096")
    assert key(ctx.join_submit(batch), batch.bundles["read_id"]) == whole
    monkeypatch.delenv("THB_JOIN_CHUNK_READS")
    # shards of the reads, submitted in reverse order (what the multi-GPU join does per rank)
    got = []
    nb = batch.n_bundles
    for lo, hi in [(2 * nb // 3, nb), (nb // 3, 2 * nb // 3), (0, nb // 3)]:
        h0 = int(batch.bundles["hit_begin"][lo]); h1 = int(batch.bundles["hit_begin"][hi]) if hi < nb else len(batch.hits)
        e0 = int(batch.bundles["ops_begin"][lo]); e1 = int(batch.bundles["ops_begin"][hi]) if hi < nb else len(batch.ops_ext)
        bu = batch.bundles[lo:hi].copy(); bu["hit_begin"] -= h0; bu["ops_begin"] -= e0
        part = synth.PackedJoinBatch(batch.n_segs, batch.read_words, bu, np.ascontiguousarray(batch.seg_count[lo:hi]),
                                     np.ascontiguousarray(batch.reads[lo:hi]), np.ascontiguousarray(batch.hits[h0:h1]), np.ascontiguousarray(batch.ops_ext[e0:e1]))
        got += key(ctx.join_submit(part), bu["read_id"])
    assert sorted(got) == whole
    ctx.close()


def _manual_join_batch(reads):
    """reads: list of dict(seq=ascii, hits=[[(ref_id, left, len, mism, anti)] per segment]) -> PackedJoinBatch (single-match hits)."""
    nseg = max(len(r["hits"]) for r in reads)
    L = max(len(r["seq"]) for r in reads); rw = (L + 63) // 64
    bundles = np.zeros(len(reads), dtype=synth.JBUNDLE_DTYPE); segc = np.zeros((len(reads), nseg), dtype="<u2")
    rd = np.zeros((len(reads), 3 * rw), dtype="<u8"); full = []
    for i, r in enumerate(reads):
        codes = synth.codes_from_ascii(r["seq"]); pad = np.zeros((1, L), dtype=np.uint8); pad[0, :len(codes)] = codes
        rd[i] = synth.pack_reads(pad, rw)[0]
        bundles[i]["read_id"] = i + 1; bundles[i]["hit_begin"] = len(full); bundles[i]["read_len"] = len(codes); bundles[i]["n_segs"] = len(r["hits"])
        for s, hs in enumerate(r["hits"]):
            segc[i, s] = len(hs)
            for (rid, left, ln, mm, anti) in hs:
                full.append((rid, left, 1, (1 if anti else 0) | (2 if s == len(r["hits"]) - 1 else 0), mm, 0, [(ln << 4) | 1] + [0] * 8))
    fa = np.array([(a, b, c, d, e, f, tuple(g)) for (a, b, c, d, e, f, g) in full], dtype=synth.JHIT_FULL_DTYPE) if full else np.zeros(0, dtype=synth.JHIT_FULL_DTYPE)
    heads, ext, ops_begin = synth.pack_join_hits(fa, bundles["hit_begin"].astype(np.int64))
    bundles["ops_begin"] = ops_begin
    return synth.PackedJoinBatch(nseg, rw, bundles, np.ascontiguousarray(segc), np.ascontiguousarray(rd), heads, ext)


def test_join_guards_and_budget():
    """join_segments_for_read's multihit guard (long_spanning_reads.cpp:2624-2632: a read with more than max_seg_multihits hits in
    a segment is skipped under bowtie2) and dfs_seg_hits' budget of 10,000 complete chains per first-segment hit (2647, 2236):
    checked through the chain counter of the C ABI; plus the exact abutting read, an empty batch and a one-segment read."""
    rng = np.random.default_rng(12)
    ref = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 400_000)])
    refimg = synth.build_ref_image(["c1"], [synth.codes_from_ascii(ref)])
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    ctx = capi.Context(0); ctx.ref_upload(refimg)
    ctx.join_begin(P, np.zeros(0, dtype=synth.JUNCTION_DTYPE), np.zeros(0, dtype=synth.INSERTION_DTYPE))
    read = ref[1000:1101]
    exact = [[(1, 1000, 25, 0, 0)], [(1, 1025, 25, 0, 0)], [(1, 1050, 25, 0, 0)], [(1, 1075, 26, 0, 0)]]
    # (a) exact read -> one alignment 101M; (b) 41 hits in segment 1 -> skipped; (c) the same with 40 -> processed
    many = lambda k: [(1, 1025, 25, 0, 0)] + [(1, 5000 + 40 * j, 25, 0, 0) for j in range(k - 1)]
    b = _manual_join_batch([dict(seq=read, hits=exact),
                            dict(seq=read, hits=[exact[0], many(41), exact[2], exact[3]]),
                            dict(seq=read, hits=[exact[0], many(40), exact[2], exact[3]])])
    out = ctx.join_submit(b)
    got = sorted((int(r["bundle"]), int(r["left"]), int(r["n_ops"]), int(r["ops"][0])) for r in out)
    assert got == [(0, 1000, 1, (101 << 4) | 1), (2, 1000, 1, (101 << 4) | 1)]
    # (d) budget: 2 first-segment hits x 25^3 mutually compatible chains each -> 10,000 leaves per first-segment hit
    spaced = lambda base, ln: [(1, base + 200 * j, ln, 0, 0) for j in range(25)]
    b = _manual_join_batch([dict(seq=read, hits=[[(1, 1000, 25, 0, 0), (1, 1200, 25, 0, 0)], spaced(10_000, 25), spaced(50_000, 25), spaced(100_000, 26)])])
    assert len(ctx.join_submit(b)) == 0
    t = ctx.join_timing()
    assert t.n_chains == 1 + 1 + 20_000, t.n_chains          # counters accumulate since thb_join_begin: (a) + (c) + the budgeted read
    # (e) empty batch, one-segment read
    e = _manual_join_batch([dict(seq=read, hits=exact)])
    empty = synth.PackedJoinBatch(e.n_segs, e.read_words, e.bundles[:0], e.seg_count[:0], e.reads[:0], e.hits[:0], e.ops_ext[:0])
    assert len(ctx.join_submit(empty)) == 0
    one = _manual_join_batch([dict(seq=ref[2000:2030], hits=[[(1, 2000, 30, 0, 0)]])])
    out = ctx.join_submit(one)
    assert len(out) == 1 and int(out[0]["left"]) == 2000 and int(out[0]["ops"][0]) == (30 << 4) | 1
    ctx.close()


@pytest.mark.parametrize("name", helpers.join_golden_cases())
def test_join_matches_reference_golden(name):
    helpers.check_join_golden(name)


def resident_handoff_check(n_pairs=1500, seed=353):
    """thb_segjuncs_finish_resident + thb_join_begin_resident (sets never leave the device) against thb_segjuncs_finish +
    thb_join_begin with the host arrays: same sets after thb_segjuncs_fetch, same joined records; with deletions and insertions
    in the sets (the junction / deletion merge) and without."""
    total = 0
    for indel_prob in (0.4, 0.0):
        wl = synth.generate(synth.SynthConfig(keep_candidates=True, contig_lens=(200_000, 80_000), n_pairs=n_pairs, seed=seed, indel_prob=indel_prob))
        P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
        ctx = capi.Context(0); ctx.ref_upload(wl.ref)
        batches = helpers.pack_both(wl)
        ctx.segjuncs_begin(P)
        for b in batches:
            ctx.segjuncs_submit(b)
        want = ctx.segjuncs_finish(True)
        juncs, ins = capi.join_sets_from_results(want)
        jb = [synth.pack_join_side(wl, s, want.junctions) for s in (wl.left, wl.right)]
        ctx.join_begin(P, juncs, ins)
        ref = [joined_to_keys(ctx.join_submit(b), b, P) for b in jb]
        if indel_prob:
            assert len(want.deletions) > 20 and len(want.insertions) > 20
        # second pass over the same batches, resident hand-off
        ctx.segjuncs_begin(P)
        for b in batches:
            ctx.segjuncs_submit(b)
        raw = ctx.segjuncs_finish_resident()
        assert (raw.n_junctions, raw.n_deletions, raw.n_insertions) == (len(want.junctions), len(want.deletions), len(want.insertions))
        ctx.join_begin_resident(P)
        got = [joined_to_keys(ctx.join_submit(b), b, P) for b in jb]
        ctx.segjuncs_fetch()
        res = capi.SegJuncsResults.from_c(raw, copy=True)
        helpers.assert_same_results(res, want, "resident finish")
        for g, r in zip(got, ref):
            assert g == r
            total += len(g)
        with pytest.raises(capi.ThbError):
            capi.Context(0).join_begin_resident(P)          # no finished pass in that context
        ctx.close()
    return total


@pytest.mark.gpu
def test_resident_handoff_equals_host_handoff():
    assert resident_handoff_check() > 1000
