"""CPU tests: the oracle restatement (oracle/segjuncs_oracle.c) is pinned against the outputs of the
reference's own segment_juncs binary -- committed golden files and, when oracle/_ref is present, a
live run on identical BAM/FASTA inputs."""
import os
import tempfile

import numpy as np
import pytest

import helpers
import kat
from tophat_b200 import capi, synth
from oracle import pyoracle


@pytest.mark.parametrize("name", helpers.golden_cases())
def test_oracle_matches_golden(name):
    wl, P, want = helpers.load_golden(name)
    res, _ = pyoracle.segjuncs(P, wl.ref, helpers.pack_both(wl, P))
    got = helpers.as_text(res, wl.ref.names)
    for k in want:
        assert got[k] == want[k], "%s: segment.%s differs from the reference binary's output" % (name, k)


def test_golden_nonempty():
    tot = {"juncs": 0, "insertions": 0, "deletions": 0}
    for name in helpers.golden_cases():
        for k in tot:
            tot[k] += sum(1 for _ in open(os.path.join(helpers.GOLDEN, name, "segment." + k)))
    assert tot["juncs"] > 1000 and tot["insertions"] > 100 and tot["deletions"] > 100


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,indel,threads", [(31, 0.0, 1), (32, 0.4, 1), (33, 0.2, 2)])
def test_oracle_matches_live_reference(seed, indel, threads):
    wl = synth.generate(synth.SynthConfig(contig_lens=(250_000, 90_000), n_pairs=2500, seed=seed, indel_prob=indel))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    res, _ = pyoracle.segjuncs(P, wl.ref, helpers.pack_both(wl))
    got = helpers.as_text(res, wl.ref.names)
    with tempfile.TemporaryDirectory() as td:
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg,
                                          threads=threads)
        for k in ("juncs", "insertions", "deletions"):
            assert got[k] == open(outs[k]).read(), "segment.%s differs (seed %d)" % (k, seed)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,threads,min_dist", [(41, 1, 20000), (43, 2, 10000000)])
def test_oracle_fusions_match_live_reference(seed, threads, min_dist):
    """find_fusions / detect_fusion (segment_juncs.cpp:2976-3291, 2629-2805) + the fusion writer (5096-5180)."""
    wl = synth.generate(synth.SynthConfig(contig_lens=(250_000, 90_000, 60_000), n_pairs=2500, seed=seed, indel_prob=0.1, fusion_frac=0.2))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20, fusion_search=1, fusion_min_dist=min_dist)
    res, cnt = pyoracle.segjuncs(P, wl.ref, helpers.pack_both(wl, P))
    got = helpers.as_text(res, wl.ref.names)
    assert cnt.n_fusion_tasks > 500 and len(set(int(d) for d in res.fusions["dir"])) == 4      # ff, fr, rf, rr all occur
    with tempfile.TemporaryDirectory() as td:
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        opts = pyoracle.tophat_common_opts(extra=["--fusion-search", "--fusion-min-dist", str(min_dist)])
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts, threads=threads)
        for k in ("juncs", "insertions", "deletions", "fusions"):
            assert got[k] == open(outs[k]).read(), "segment.%s differs (seed %d)" % (k, seed)


def _run_kat(case):
    contigs, reads, expected = case
    ref, batch = helpers.manual_workload(contigs, reads)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    res, _ = pyoracle.segjuncs(P, ref, [batch])
    return res, ref, expected


def test_kat_junction():
    for case in (kat.kat_junction(), kat.kat_junction_seg1_unmapped()):
        res, ref, exp = _run_kat(case)
        lines = pyoracle.format_juncs(res.junctions, ref.names).splitlines()
        assert "%s\t%d\t%d\t%s" % exp in lines


def test_kat_deletion():
    res, ref, exp = _run_kat(kat.kat_deletion())
    assert "%s\t%d\t%d" % exp in pyoracle.format_deletions(res.deletions, ref.names).splitlines()


def test_kat_insertion():
    res, ref, exp = _run_kat(kat.kat_insertion())
    assert "%s\t%d\t%d\t%s" % exp in pyoracle.format_insertions(res.insertions, ref.names).splitlines()


def test_kat_q0_quirk():
    res, ref, exp = _run_kat(kat.kat_q0_quirk())
    lines = pyoracle.format_juncs(res.junctions, ref.names).splitlines()
    for e in exp:
        assert "%s\t%d\t%d\t%s" % e in lines


def test_junction_order_axioms():
    """Junction::operator< axioms (tests/unit_tests/testjunctions.cpp:447-660 of the reference): the result
    arrays are strictly increasing in (refid, left, right, antisense)."""
    wl, P, _ = helpers.load_golden("indel_heavy")
    res, _ = pyoracle.segjuncs(P, wl.ref, helpers.pack_both(wl))
    for arr in (res.junctions, res.deletions):
        key = [(int(r["ref_id"]), int(r["left"]), int(r["right"]), int(r["antisense"])) for r in arr]
        assert all(a < b for a, b in zip(key, key[1:]))
    ik = [(int(r["ref_id"]), int(r["left"]), int(r["len"])) for r in res.insertions]
    assert all(a < b for a, b in zip(ik, ik[1:]))


def test_empty_batch():
    wl, P, _ = helpers.load_golden("splice_2contig")
    b = helpers.pack_both(wl)[0]
    empty = synth.PackedBatch(b.n_segs, b.read_words, b.bundles[:0], b.seg_count[:0], b.reads[:0], b.hits[:0], b.partner_hits[:0])
    res, cnt = pyoracle.segjuncs(P, wl.ref, [empty])
    assert len(res.junctions) == 0 and cnt.n_windows == 0


@pytest.mark.parametrize("name", helpers.reference_input_cases())
def test_oracle_on_the_references_own_test_inputs(name):
    """fusion_test/ read sets of the reference (BASELINE configs[0] / configs[4]) through the oracle: single-end, --bowtie1,
    --fusion-search --fusion-min-dist 500 --max-intron-length 500; expected = outputs of the reference binary on the same files."""
    wl, P, want = helpers.load_reference_input_case(name)
    batch = synth.pack_side(wl.left, None, False, True)
    res, _ = pyoracle.segjuncs(P, wl.ref, [batch])
    got = helpers.reference_input_texts(res, wl.ref.names)
    for k in want:
        assert got[k] == want[k], "%s: segment.%s differs from the reference binary's output" % (name, k)


def _opts_with(over):
    o = pyoracle.tophat_common_opts(50, 20)
    for k, v in over.items():
        o[o.index(k) + 1] = str(v)
    return o


OPTION_VARIANTS = [
    ({"--segment-mismatches": 1}, dict(segment_mismatches=1)),
    ({"--segment-mismatches": 3}, dict(segment_mismatches=3)),
    ({"--max-insertion-length": 5, "--max-deletion-length": 6}, dict(max_insertion_length=5, max_deletion_length=6)),
    ({"--min-segment-intron": 200, "--max-segment-intron": 3000}, dict(min_segment_intron_length=200, max_segment_intron_length=3000)),
    ({"--max-seg-multihits": 2}, dict(max_seg_multihits=2)),
]


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("opts,params", OPTION_VARIANTS)
def test_oracle_matches_live_reference_under_option_variants(opts, params):
    """Non-default values of the options the path reads (segment mismatches -> rescue seed length, indel lengths, segment intron
    bounds, multihit guard): oracle vs the reference binary on identical files."""
    wl = synth.generate(synth.SynthConfig(contig_lens=(250_000, 90_000), n_pairs=2500, seed=61, indel_prob=0.4, decoy_rate=1.5))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20, **params)
    res, _ = pyoracle.segjuncs(P, wl.ref, helpers.pack_both(wl))
    got = helpers.as_text(res, wl.ref.names)
    with tempfile.TemporaryDirectory() as td:
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=_opts_with(opts))
        for k in ("juncs", "insertions", "deletions"):
            assert got[k] == open(outs[k]).read(), "segment.%s differs under %r" % (k, opts)
