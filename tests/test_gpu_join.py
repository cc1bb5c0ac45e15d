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
    with pytest.raises(capi.ThbError):
        ctx.join_begin(capi.default_params(fusion_search=1), j[::-1].copy(), np.zeros(0, dtype=synth.INSERTION_DTYPE))
    ctx.close()
