"""Kernel-logic checks WITHOUT a GPU: tophat_b200/csrc compiled for the host against tests/emu/shim (every warp = 32
cooperative fibers, warp collectives = barriers; see tests/emu/shim/cuda_runtime.h) and compared with the oracle.
This is a development harness for catching logic and warp-divergence bugs before GPU time is spent; the parity gate
remains tests/test_gpu_*.py (-m gpu), which run the real sm_100a build."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import helpers
import kat
from tophat_b200 import capi, synth
from oracle import pyoracle


@pytest.mark.parametrize("kw,over", [
    (dict(contig_lens=(300_000, 100_000), n_pairs=2500, seed=301, indel_prob=0.3), {}),
    (dict(contig_lens=(200_000,), n_pairs=1500, seed=302, decoy_rate=3.0, n_rate=0.01), dict(library_type=2)),
    (dict(contig_lens=(200_000,), n_pairs=1500, seed=303, read_len=150, indel_prob=0.2), dict(inner_dist_mean=10, inner_dist_std_dev=40)),
])
def test_emu_segjuncs_matches_oracle(emu_lib, kw, over):
    wl = synth.generate(synth.SynthConfig(**kw))
    o = dict(inner_dist_mean=50, inner_dist_std_dev=20); o.update(over)
    P = capi.default_params(**o)
    batches = helpers.pack_both(wl)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
    want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, str(kw))
    assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted) == \
        (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted)


@pytest.mark.parametrize("case", ["kat_junction", "kat_junction_seg1_unmapped", "kat_deletion", "kat_insertion", "kat_q0_quirk"])
def test_emu_known_answers(emu_lib, case):
    contigs, reads, exp = getattr(kat, case)()
    ref, batch = helpers.manual_workload(contigs, reads)
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    got, _ = helpers.gpu_segjuncs(P, ref, [batch])
    want, _ = pyoracle.segjuncs(P, ref, [batch])
    helpers.assert_same_results(got, want, case)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("chunk", [None, 96])
def test_emu_join_matches_reference_records(emu_lib, monkeypatch, chunk):
    from test_gpu_join import joined_to_keys
    if chunk:                                                       # many pipeline chunks per submit
        monkeypatch.setenv("THB_JOIN_CHUNK_READS", str(chunk))
    wl = synth.generate(synth.SynthConfig(keep_truth=True, contig_lens=(200_000, 80_000), n_pairs=1200, seed=511, indel_prob=0.4))
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
            got = joined_to_keys(ctx.join_submit(batch), batch, P)
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg,
                                                       side=sname, tag=".ref")
            _, recs = pyoracle.read_bam(ref_bam)
            want = set((int(r[0]), names.index(r[2]) + 1, r[3], r[5], r[1], r[11]["NM"]) for r in recs)
            assert got == want, "%s: only ours %r; only reference %r" % (sname, sorted(got - want)[:3], sorted(want - got)[:3])
            assert len(want) > 50
    ctx.close()


# ---- --fusion-search ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kw,over", [
    (dict(contig_lens=(250_000, 90_000, 60_000), n_pairs=2500, seed=321, indel_prob=0.1, fusion_frac=0.15), dict(fusion_min_dist=20000)),
    (dict(contig_lens=(200_000,), n_pairs=1500, seed=322, fusion_frac=0.3, decoy_rate=2.0), {}),
    (dict(contig_lens=(150_000, 50_000), n_pairs=1200, seed=323, read_len=150, fusion_frac=0.2, n_rate=0.005), dict(fusion_anchor_length=30, fusion_min_dist=100)),
])
def test_emu_fusions_match_oracle(emu_lib, kw, over):
    wl = synth.generate(synth.SynthConfig(**kw))
    o = dict(inner_dist_mean=50, inner_dist_std_dev=20, fusion_search=1); o.update(over)
    P = capi.default_params(**o)
    batches = helpers.pack_both(wl, P)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
    want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, str(kw))
    assert len(want.fusions) > 50
    assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted, t.n_fusion_tasks) == \
        (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted, cnt.n_fusion_tasks)


@pytest.mark.parametrize("name", [n for n in helpers.golden_cases() if n.startswith("fusion")])
def test_emu_fusion_goldens(emu_lib, name):
    wl, P, want = helpers.load_golden(name)
    got, _ = helpers.gpu_segjuncs(P, wl.ref, helpers.pack_both(wl, P))
    txt = helpers.as_text(got, wl.ref.names)
    for k in want:
        assert txt[k] == want[k], "%s: segment.%s differs from the reference binary's output" % (name, k)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("extra", [["--fusion-search", "--fusion-min-dist", "20000"],
                                   ["--fusion-search", "--fusion-ignore-chromosomes", "chrS2", "--fusion-do-not-resolve-conflicts"]])
def test_emu_cli_fusions_match_reference_binary(extra):
    """The segment_juncs host executable (BAM ingestion, bundle rules incl. the fusion-only reads, fusion writer) linked
    against the emulated library, against the reference binary: all four output files byte for byte."""
    import sys
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "emu"))
    import build_emu
    exe = build_emu.build_cli("segment_juncs")
    with tempfile.TemporaryDirectory() as td:
        wl = synth.generate(synth.SynthConfig(contig_lens=(200_000, 90_000, 50_000), n_pairs=2000, seed=331, indel_prob=0.1, fusion_frac=0.2))
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        opts = pyoracle.tophat_common_opts(50, 20, extra)
        ref = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts, tag=".ref")
        ours = pyoracle.run_segment_juncs(exe, files, bams, td, nseg, opts=opts, tag=".emu")
        for k in ("juncs", "insertions", "deletions", "fusions"):
            a, b = open(ours[k]).read(), open(ref[k]).read()
            assert a == b, "segment.%s differs from the reference binary (%d vs %d lines)" % (k, a.count("\n"), b.count("\n"))
        assert open(ref["fusions"]).read().count("\n") > 50


def test_emu_allgather_world1():
    """thb_segjuncs_allgather's staging / padding / union of all four record kinds, in a fresh process with a stand-in for libnccl
    (tests/emu/fake_nccl.c): world size 1, and a replicating "world size 3" (sort + unique union, first-wins insertions, fusion
    counts, the refused late submit, the resident hand-off of the gathered union)."""
    import subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tests", "emu", "allgather_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "allgather ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name", helpers.join_golden_cases())
def test_emu_join_golden(emu_lib, name):
    helpers.check_join_golden(name)


def test_emu_growth_paths_tiny_capacities():
    """Overflow -> grow -> repeat paths of every device structure (tests/emu/tiny_caps_check.py, fresh process with THB_TINY_CAPS=1)."""
    import subprocess, sys
    env = dict(os.environ, THB_TINY_CAPS="1")
    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tests", "emu", "tiny_caps_check.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "tiny caps ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", helpers.reference_input_cases())
def test_emu_on_the_references_own_test_inputs(emu_lib, name):
    """The reference's own fusion_test/ read sets (BASELINE configs[0] / configs[4]): single-end, --bowtie1, --fusion-search
    --fusion-min-dist 500 --max-intron-length 500; all four outputs equal the reference binary's (tests/golden/reference_*)."""
    wl, P, want = helpers.load_reference_input_case(name)
    batch = synth.pack_side(wl.left, None, False, True)
    got, _ = helpers.gpu_segjuncs(P, wl.ref, [batch])
    txt = helpers.reference_input_texts(got, wl.ref.names)
    for k in want:
        assert txt[k] == want[k], "%s: segment.%s differs from the reference binary's output" % (name, k)
    oracle, _ = pyoracle.segjuncs(P, wl.ref, [batch])
    helpers.assert_same_results(got, oracle, name)


@pytest.mark.parametrize("params", [
    dict(segment_mismatches=1), dict(segment_mismatches=3), dict(max_insertion_length=5, max_deletion_length=6),
    dict(min_segment_intron_length=200, max_segment_intron_length=3000), dict(max_seg_multihits=2),
    dict(fusion_search=1, fusion_anchor_length=10, fusion_min_dist=1000, segment_mismatches=1),
    dict(bowtie2=0, fusion_search=1, fusion_min_dist=20000, max_seg_multihits=2),
])
def test_emu_option_variants_match_oracle(emu_lib, params):
    """Non-default values of every option the path reads (the oracle is pinned on the reference binary under the same variants:
    tests/test_oracle.py::test_oracle_matches_live_reference_under_option_variants)."""
    wl = synth.generate(synth.SynthConfig(contig_lens=(250_000, 90_000), n_pairs=2000, seed=61, indel_prob=0.4, decoy_rate=1.5,
                                          fusion_frac=0.15 if params.get("fusion_search") else 0.0))
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20, **params)
    batches = helpers.pack_both(wl, P)
    got, t = helpers.gpu_segjuncs(P, wl.ref, batches)
    want, cnt = pyoracle.segjuncs(P, wl.ref, batches)
    helpers.assert_same_results(got, want, str(params))
    assert (t.n_windows, t.n_indel_tasks, t.n_rescue_tasks, t.n_juncs_emitted, t.n_fusion_tasks) == \
        (cnt.n_windows, cnt.n_indel_tasks, cnt.n_rescue_tasks, cnt.n_juncs_emitted, cnt.n_fusion_tasks)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("over", [
    {"--max-insertion-length": 5, "--max-deletion-length": 6},
    {"--min-report-intron": 150, "--max-report-intron": 3000},
    {"--read-mismatches": 1, "--read-edit-dist": 1, "--read-gap-length": 1},
    {"--min-anchor": 5, "--max-seg-multihits": 3},
])
def test_emu_cli_join_option_variants_match_reference(over):
    """long_spanning_reads executable (emulated library) under non-default options of the join (indel lengths, reported intron bounds,
    read-level filters, multihit guard) against the reference binary: records identical, in order."""
    import sys
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "emu"))
    import build_emu
    exe = build_emu.build_cli("long_spanning_reads")
    opts = pyoracle.tophat_common_opts(50, 20)
    for k, v in over.items():
        opts[opts.index(k) + 1] = str(v)
    with tempfile.TemporaryDirectory() as td:
        wl = synth.generate(synth.SynthConfig(keep_truth=True, contig_lens=(250_000, 80_000), n_pairs=2000, seed=71, indel_prob=0.5, decoy_rate=1.0))
        files = synth.write_pipeline_files(wl, td)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, td, nseg)
        outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts)
        jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg)
        n = 0
        for side in ("left", "right"):
            ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg, side=side, tag=".ref", opts=opts)
            our_bam = pyoracle.run_long_spanning_reads(exe, files, bams, jin, outs, td, nseg, side=side, tag=".emu", opts=opts, threads=3)
            a, b = pyoracle.read_bam(our_bam)[1], pyoracle.read_bam(ref_bam)[1]
            assert a == b, "%s: records differ under %r (%d vs %d)" % (side, over, len(a), len(b))
            n += len(b)
        assert n > 100


def test_emu_join_tile_kernel_equals_queue_kernels():
    """Multi-hit-heavy reads (more than 8 hits in a segment, more than 4 chains per read): the tile kernel's generic walk and parking
    overflow give the records of round 1's queue kernels (tests/emu/join_variant_check.py, one process per variant)."""
    outs = []
    for legacy in (False, True):
        env = dict(os.environ); env.pop("THB_JOIN_LEGACY", None); env.pop("THB_JOIN_TILE", None); env.pop("THB_CHECK_GPU", None)
        if not legacy:
            env["THB_JOIN_TILE"] = "1"
        r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tests", "emu", "join_variant_check.py")], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("JOIN_DIGEST")][-1])
    assert outs[0] == outs[1] and int(outs[0].split()[-1]) > 3000


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("name", ["reference_test_total_inter", "reference_test_indel_intra2"])
def test_emu_fusion_test_sets_join_matches_reference(name):
    """--fusion-search in the join (fusion_join_kernel.cuh, emulated): the reference's own fusion_test sets, junction index with fusion
    contigs from the reference's juncs_db; records (incl. the two-record XF form) equal the reference binary's."""
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "emu"))
    import build_emu
    exe = build_emu.build_cli("long_spanning_reads")
    with tempfile.TemporaryDirectory() as td:
        wl, files, bams, jin, outs, nseg, opts = helpers.fusion_test_pipeline(name, td)
        ref_bam = pyoracle.run_long_spanning_reads(os.path.join(pyoracle.REF_DIR, "long_spanning_reads"), files, bams, jin, outs, td, nseg, side="left",
                                                   tag=".ref", opts=opts, fusions=outs["fusions"])
        our_bam = pyoracle.run_long_spanning_reads(exe, files, bams, jin, outs, td, nseg, side="left", tag=".emu", opts=opts, fusions=outs["fusions"])
        _, a = pyoracle.read_bam(our_bam); _, b = pyoracle.read_bam(ref_bam)
        assert a == b and sum(1 for r in b if "XF" in r[11]) > 200
        # and with the junction index (fusion contigs included) built and searched inside the executable
        flank_bam = pyoracle.run_long_spanning_reads(exe, files, bams, jin, outs, td, nseg, side="left", tag=".emuflank", opts=opts, fusions=outs["fusions"],
                                                     with_spliced=False, env=dict(os.environ, TOPHAT_GPU_FLANK_SEARCH="1", TOPHAT_GPU_FLANK_LENGTH="26"))
        _, c = pyoracle.read_bam(flank_bam)
        assert c == b


@pytest.mark.parametrize("name", ["v2_101bp", "v2_m2_suppression", "v3_two_word_contigs", "v3_direct_buckets", "v0_exact", "v2_long_last_segment"])
def test_emu_flank_matcher_equals_oracle(emu_lib, name):
    """junction-flank matcher kernels (flank_kernel.cuh) under emulation against oracle/flank_oracle.py"""
    import numpy as np
    import test_flank
    seed, v, bounds, n_reads, max_hits, npol, kw = test_flank.FLANK_CASES[name]
    case = test_flank.flank_case(seed, v, np.asarray(bounds), n_reads, **kw)
    n, _ = test_flank.check_flank(case, v, max_hits, npol)
    assert n > 20


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_emu_flank_pipeline_matches_reference_binaries(emu_lib):
    """stage 1 -> junction-flank matcher -> join, against segment_juncs -> juncs_db -> long_spanning_reads of the reference"""
    import test_flank
    assert test_flank.flank_pipeline_check() > 500


def test_emu_resident_handoff_equals_host_handoff(emu_lib):
    import test_gpu_join
    assert test_gpu_join.resident_handoff_check() > 1000


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_emu_cli_reads_as_fastq_text(monkeypatch):
    """long_spanning_reads with the reads argument as FASTQ text instead of BAM (tests/test_cli_long_spanning_reads.py), emulated library"""
    import test_cli_long_spanning_reads as t
    monkeypatch.setenv("THB_TEST_EMU", "1")
    t.test_cli_reads_as_fastq_text(False)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_emu_cli_long_spanning_reads_searches_the_junction_index_itself(emu_lib, monkeypatch):
    import test_flank
    monkeypatch.setenv("THB_TEST_EMU", "1")
    assert test_flank.flank_pipeline_check(cli_bin=helpers.our_bin("long_spanning_reads")) > 500


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_emu_cli_in_process_junction_index_with_two_read_lengths(emu_lib, monkeypatch):
    """101-bp and 75-bp reads in one run: per-length segment layouts in the in-process junction index search (tests/test_flank.py)"""
    import test_flank
    monkeypatch.setenv("THB_TEST_EMU", "1")
    assert test_flank.mixed_length_cli_check(helpers.our_bin("long_spanning_reads")) > 400


@pytest.mark.parametrize("name", ["flank_v2_101bp", "flank_v3_two_word", "flank_v2_m2"])
def test_emu_flank_matcher_equals_golden(emu_lib, name):
    """the matcher against the committed golden vectors (FASTA of the reference's juncs_db, placements over it)"""
    import test_flank
    assert test_flank.flank_golden_check(name) > 40
