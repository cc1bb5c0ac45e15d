"""Drop-in boundary test: our segment_juncs executable (C++ host over libtophat_b200.so) against the reference's
own segment_juncs binary on identical BAM / FASTA inputs and the argv tophat.py builds (tophat.py:3097-3112).
Outputs are compared byte for byte."""
import os
import subprocess
import tempfile

import pytest

import helpers
from tophat_b200 import build, synth
from oracle import pyoracle

OUR_BIN = os.path.join(build.BIN_DIR, "segment_juncs")


def _prepare(td, cfg):
    wl = synth.generate(cfg)
    files = synth.write_pipeline_files(wl, td)
    nseg = len(wl.left.seg_hits)
    bams = pyoracle.make_bams(files, td, nseg)
    return wl, files, bams, nseg


def test_host_binary_builds():
    build.build_library()
    outs = build.build_host_binaries()
    assert OUR_BIN in outs and os.access(OUR_BIN, os.X_OK)


def test_usage_and_empty_segment_list_exit_codes():
    build.build_all()
    r = subprocess.run([OUR_BIN], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("segment_juncs v")
    r = subprocess.run([OUR_BIN, "--no-such-option"], capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("kw,paired,extra", [
    (dict(contig_lens=(300_000, 120_000), n_pairs=3000, seed=301, indel_prob=0.3), True, []),
    (dict(contig_lens=(300_000,), n_pairs=3000, seed=302, indel_prob=0.2, n_rate=0.003), False, []),
    (dict(contig_lens=(250_000, 80_000), n_pairs=2500, seed=303), True, ["--library-type", "fr-secondstrand"]),
    (dict(contig_lens=(250_000, 80_000, 60_000), n_pairs=3000, seed=305, indel_prob=0.1, fusion_frac=0.2), True,
     ["--fusion-search", "--fusion-min-dist", "20000"]),
    (dict(contig_lens=(250_000, 80_000), n_pairs=2500, seed=306, fusion_frac=0.2), True,
     ["--fusion-search", "--fusion-ignore-chromosomes", "chrS2", "--fusion-do-not-resolve-conflicts"]),
])
def test_cli_matches_reference_binary(kw, paired, extra):
    OUR_BIN = helpers.our_bin("segment_juncs")
    with tempfile.TemporaryDirectory() as td:
        wl, files, bams, nseg = _prepare(td, synth.SynthConfig(**kw))
        opts = pyoracle.tophat_common_opts(50, 20, extra)
        ref = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts,
                                         paired=paired, tag=".ref")
        ours = pyoracle.run_segment_juncs(OUR_BIN, files, bams, td, nseg, opts=opts, paired=paired, tag=".b200")
        for k in ("juncs", "insertions", "deletions", "fusions"):
            a, b = open(ours[k]).read(), open(ref[k]).read()
            assert a == b, "segment.%s differs from the reference binary (%d vs %d lines)" % (k, a.count("\n"), b.count("\n"))
        assert open(ref["juncs"]).read().count("\n") > 50
        log = open(ours["log"]).read()
        assert "found" in log and "potential split-segment junctions" in log


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
def test_cli_no_hits_exit_zero():
    OUR_BIN = helpers.our_bin("segment_juncs")
    with tempfile.TemporaryDirectory() as td:
        wl, files, bams, nseg = _prepare(td, synth.SynthConfig(contig_lens=(100_000,), n_pairs=200, seed=304))
        outs = [os.path.join(td, "o.%s" % k) for k in ("j", "i", "d", "f")]
        cmd = [OUR_BIN] + pyoracle.tophat_common_opts() + ["--sam-header", files["header"], files["fasta"]] + outs + \
              [bams["left_reads"], bams["left_mapped"], ""]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0 and "No hits to process, exiting" in r.stderr
        assert all(os.path.exists(o) for o in outs)


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")
@pytest.mark.parametrize("threads", [4, 9])
def test_cli_read_id_ranges_equal_reference_p1(threads):
    """-p N > 1: our executable works through (mate side) x (read-id range) tasks on N threads, every stream positioned by its own
    .index side file; all four outputs (insertions depend on the first-wins order over sides and ranges) must equal the reference's
    -p1 output -- the reference's own -p N output can lack the hit groups at its range boundaries."""
    OUR_BIN = helpers.our_bin("segment_juncs")
    with tempfile.TemporaryDirectory() as td:
        wl, files, bams, nseg = _prepare(td, synth.SynthConfig(contig_lens=(300_000, 100_000), n_pairs=9000, seed=307, indel_prob=0.5, fusion_frac=0.1))
        opts = pyoracle.tophat_common_opts(50, 20, ["--fusion-search", "--fusion-min-dist", "20000"])
        ref = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=opts, tag=".ref", threads=1)
        ours = pyoracle.run_segment_juncs(OUR_BIN, files, bams, td, nseg, opts=opts, tag=".b200p%d" % threads, threads=threads)
        for k in ("juncs", "insertions", "deletions", "fusions"):
            assert open(ours[k]).read() == open(ref[k]).read(), "segment.%s differs from the reference's -p1 output" % k
        assert open(ref["insertions"]).read().count("\n") > 100
        assert len(open(files["left_fq"]).read().split("\n")) > 4 * 8000          # enough reads for several index entries per file
