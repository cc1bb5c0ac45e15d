"""CPU tests of the C++ host I/O layer (no GPU): BGZF/BAM reader and writer, FASTA loader + genome image packer, and the
BAMHitFactory / SplicedBAMHitFactory record extraction, on files produced by the reference's own prep_reads / fix_map_ordering
/ juncs_db binaries."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tophat_b200 import build, synth
from oracle import pyoracle

SELFTEST = os.path.join(build.BIN_DIR, "thb_host_selftest")
pytestmark = pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref (reference binaries) not present")


@pytest.fixture(scope="module")
def pipeline():
    build.build_all()
    td = tempfile.mkdtemp(prefix="thb_hostio_")
    wl = synth.generate(synth.SynthConfig(contig_lens=(200_000, 60_000), n_pairs=1500, seed=601, indel_prob=0.3, keep_truth=True, ref_n_frac=0.02))
    files = synth.write_pipeline_files(wl, td)
    bams = pyoracle.make_bams(files, td, len(wl.left.seg_hits))
    yield wl, files, bams, td
    subprocess.run(["rm", "-rf", td])


def test_bam_roundtrip(pipeline):
    wl, files, bams, td = pipeline
    for key in ("left_seg1", "left_reads", "right_mapped"):
        out = os.path.join(td, key + ".copy.bam")
        r = subprocess.run([SELFTEST, "bam", bams[key], out, files["header"]], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        refs_a, a = pyoracle.read_bam(bams[key])
        refs_b, b = pyoracle.read_bam(out)
        assert int(r.stdout) == len(a) == len(b) and len(a) > 100
        if key != "left_reads":
            assert refs_a == refs_b
        assert a == b


def test_fasta_image_matches_python_packer(pipeline):
    wl, files, bams, td = pipeline
    out = os.path.join(td, "image.bin")
    r = subprocess.run([SELFTEST, "fasta", files["fasta"], files["header"], out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == wl.ref.names
    raw = open(out, "rb").read()
    nc, nb = np.frombuffer(raw, dtype="<u8", count=2)
    off = 16
    starts = np.frombuffer(raw, dtype="<u8", count=int(nc), offset=off); off += 8 * int(nc)
    lens = np.frombuffer(raw, dtype="<u4", count=int(nc), offset=off); off += 4 * int(nc)
    planes = np.frombuffer(raw, dtype="<u8", count=2 * int(nb), offset=off); off += 16 * int(nb)
    nmask = np.frombuffer(raw, dtype="<u8", count=int(nb), offset=off)
    assert (starts == wl.ref.contig_start).all() and (lens == wl.ref.contig_len).all()
    assert int(nb) == wl.ref.n_blocks and (planes == wl.ref.planes).all() and (nmask == wl.ref.nmask).all()
    assert int(nmask.astype(np.uint64).sum()) > 0          # the reference contains N runs


def test_hit_stream_records(pipeline):
    wl, files, bams, td = pipeline
    for k in range(len(wl.left.seg_hits)):
        r = subprocess.run([SELFTEST, "hits", bams["left_seg%d" % (k + 1)], files["header"]], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = [tuple(int(x) for x in l.split()) for l in r.stdout.splitlines()]
        h = wl.left.seg_hits[k]
        want = sorted((int(x["read_idx"]) + 1, int(x["ref_id"]), int(x["left"]), int(x["right"]), int(x["read_len"]), int(x["edit_dist"]), int(x["flags"])) for x in h)
        assert sorted(got) == want and len(want) > 100
        assert [g[0] for g in got] == sorted(g[0] for g in got)          # id-sorted stream, grouped


def test_spliced_hit_stream_records(pipeline):
    """SplicedBAMHitFactory: contig coordinates -> genomic left + M/N/M CIGAR (bwt_map.cpp:1469-1770, 681-883)."""
    wl, files, bams, td = pipeline
    nseg = len(wl.left.seg_hits)
    outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg)
    jin = pyoracle.make_join_inputs(wl, files, outs, td, nseg, fast=True)
    junctions = pyoracle.parse_juncs(outs["juncs"], wl.ref.names)
    spl = synth.spliced_placements(wl, wl.left, junctions)
    total = 0
    for k in range(nseg):
        r = subprocess.run([SELFTEST, "jhits", jin["left_spl%d" % (k + 1)], files["header"], "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = sorted(tuple(l.split()) for l in r.stdout.splitlines())
        d = spl[k]
        want = []
        for i in range(d["read_idx"].shape[0]):
            x = int(d["x"][i]); ln = d["ln"]
            flags = (1 if d["anti"][i] else 0) | (2 if k == nseg - 1 else 0) | (4 if d["asplice"][i] else 0)
            want.append((str(int(d["read_idx"][i]) + 1), wl.ref.names[int(d["ref_id"][i]) - 1], str(int(d["jl"][i]) - x + 1), str(flags),
                         str(int(d["nm"][i])), str(int(d["smm"][i])), "%d:1" % x, "%d:11" % (int(d["jr"][i]) - int(d["jl"][i]) - 1), "%d:1" % (ln - x)))
        assert got == sorted(want)
        total += len(want)
    assert total > 200
