"""Golden vectors from the reference's OWN test inputs (fusion_test/: BASELINE configs[0] and configs[4]).

The reference's fusion_test/run_test.sh feeds 100-bp FASTA reads over a 3-contig, 21-kb genome (testcases/test.fa) through
`tophat --fusion-search --bowtie1 --max-intron-length 500 --fusion-min-dist 500 --fusion-do-not-resolve-conflicts` and stores no
outputs.  This script (run in the build container, where /root/reference and oracle/_ref exist):
  1. stands in for bowtie with a brute-force un-gapped mapper (<= 2 mismatches, both strands) for the whole reads and for their
     25-bp segments -- what `bowtie -v 2` reports on a genome this small;
  2. writes the BAM / FASTA files of the stage boundary with the reference's own prep_reads / fix_map_ordering;
  3. runs the reference's segment_juncs with the options tophat.py derives from that command line;
  4. commits, per read set: the inputs as one compressed .npz (genome, reads, hits) and the four segment.* outputs.
tests/ rebuild the packed batches from the .npz and compare the oracle and the GPU path with the committed outputs.
"""
import json, os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tophat_b200 import synth
from oracle import pyoracle

SRC = "/root/reference/fusion_test"
SETS = ["test_junction_intra", "test_total_intra", "test_total_inter", "test_fusion_inter2", "test_indel_intra2"]
EXTRA = ["--bowtie1", "--fusion-search", "--fusion-min-dist", "500", "--fusion-do-not-resolve-conflicts"]
OPT_OVERRIDES = {"--max-report-intron": "500", "--max-segment-intron": "500", "--max-coverage-intron": "500", "--max-closure-intron": "500"}


def read_fasta(path):
    names, seqs = [], []
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            names.append(line[1:].split()[0]); seqs.append([])
        else:
            seqs[-1].append(line.upper())
    return names, ["".join(s) for s in seqs]


def brute_hits(ref_codes, seg, max_mm=2):
    """all (ref_id, left, mismatches, antisense) with <= max_mm mismatches, in (strand, contig, position) order"""
    out = []
    L = seg.shape[0]
    for anti, q in ((0, seg), (1, synth.revcomp_codes(seg))):
        for rid, codes in enumerate(ref_codes):
            if codes.shape[0] < L:
                continue
            win = np.lib.stride_tricks.sliding_window_view(codes, L)
            mm = ((win != q[None, :]) | (win > 3) | (q[None, :] > 3)).sum(axis=1)
            for pos in np.nonzero(mm <= max_mm)[0]:
                out.append((rid + 1, int(pos), int(mm[pos]), anti))
    return out


def build_workload(set_name):
    names, seqs = read_fasta(os.path.join(SRC, "testcases", "test.fa"))
    ref = synth.build_ref_image(names, [synth.codes_from_ascii(s.encode()) for s in seqs])
    _, reads = read_fasta(os.path.join(SRC, set_name + ".fasta"))
    L = max(set(len(r) for r in reads), key=[len(r) for r in reads].count)
    n_all = len(reads); reads = [r for r in reads if len(r) == L]          # the packed workload holds one read length; 2 reads of 102 bp
    if len(reads) != n_all:                                                  # in the *_intra indel sets are left out
        print("  %s: %d of %d reads are not %d bp long and are left out" % (set_name, n_all - len(reads), n_all, L))
    cfg = synth.SynthConfig(contig_lens=tuple(len(s) for s in seqs), n_pairs=len(reads), read_len=L, segment_length=25)
    codes = np.stack([synth.codes_from_ascii(r.encode()) for r in reads])
    offs, lens = synth.segment_layout(L, 25)
    nseg = len(offs)
    mapped, seg = [], [[] for _ in range(nseg)]
    unm = np.ones(len(reads), dtype=bool)
    for i in range(len(reads)):
        wh = brute_hits(ref.codes, codes[i])
        if wh:
            unm[i] = False
            for (rid, pos, mm, anti) in wh:
                mapped.append((i, rid, pos, pos + L, L, mm, (synth.HIT_ANTISENSE if anti else 0) | synth.HIT_END, 0))
            continue
        for k in range(nseg):
            for (rid, pos, mm, anti) in brute_hits(ref.codes, codes[i, offs[k]:offs[k] + lens[k]]):
                seg[k].append((i, rid, pos, pos + int(lens[k]), int(lens[k]), mm, (synth.HIT_ANTISENSE if anti else 0) | (synth.HIT_END if k == nseg - 1 else 0), 0))
    arr = lambda x: np.array(x, dtype=synth.SEGHIT_DTYPE) if x else np.zeros(0, dtype=synth.SEGHIT_DTYPE)
    left = synth.SideData(codes, np.arange(1, len(reads) + 1, dtype="<u4"), [arr(s) for s in seg], arr(mapped), unm)
    empty = synth.SideData(np.zeros((0, L), dtype=np.uint8), np.zeros(0, dtype="<u4"), [arr([]) for _ in range(nseg)], arr([]), np.zeros(0, dtype=bool))
    return synth.Workload(cfg, ref, left, empty, np.zeros((0, 4), dtype=np.int64))


def options():
    o = pyoracle.tophat_common_opts(50, 20, EXTRA)
    for k, v in OPT_OVERRIDES.items():
        o[o.index(k) + 1] = v
    return o


def main():
    assert pyoracle.build_reference()
    for name in SETS:
        wl = build_workload(name)
        out = os.path.join(ROOT, "tests", "golden", "reference_" + name)
        os.makedirs(out, exist_ok=True)
        with tempfile.TemporaryDirectory() as td:
            files = {"fasta": os.path.join(td, "ref.fa"), "header": os.path.join(td, "hdr.sam"), "left_fq": os.path.join(td, "left.fq")}
            synth.write_fasta(files["fasta"], wl.ref); synth.write_sam_header(files["header"], wl.ref); synth.write_fastq(files["left_fq"], wl.left)
            nseg = len(wl.left.seg_hits)
            sams = {"mapped": os.path.join(td, "left_mapped.sam")}
            synth.write_hits_sam(sams["mapped"], wl, wl.left, wl.left.mapped_hits, None)
            for k in range(nseg):
                sams["seg%d" % (k + 1)] = os.path.join(td, "left_seg%d.sam" % (k + 1))
                synth.write_hits_sam(sams["seg%d" % (k + 1)], wl, wl.left, wl.left.seg_hits[k], k)
            bams = {"left_reads": os.path.join(td, "left_kept_reads.bam")}
            subprocess.run([os.path.join(pyoracle.REF_DIR, "prep_reads"), "--sam-header", files["header"], "--outfile", bams["left_reads"],
                            "--index-outfile", bams["left_reads"] + ".index", "--aux-outfile", os.path.join(td, "left.info"), files["left_fq"]],
                           check=True, stderr=subprocess.DEVNULL)
            for key, sam in sams.items():
                bam = os.path.join(td, "left_kept_reads_%s.bam" % key)
                subprocess.run([os.path.join(pyoracle.REF_DIR, "fix_map_ordering"), "--sam-header", files["header"], "--index-outfile", bam + ".index", sam, bam],
                               check=True, stderr=subprocess.DEVNULL)
                bams["left_" + key] = bam
            outs = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, td, nseg, opts=options(), paired=False)
            for k in ("juncs", "insertions", "deletions", "fusions"):
                shutil.copy(outs[k], os.path.join(out, "segment." + k))
        np.savez_compressed(os.path.join(out, "inputs.npz"), contig_names=np.array(wl.ref.names), contig_codes=np.concatenate(wl.ref.codes),
                            contig_lens=np.array([c.shape[0] for c in wl.ref.codes]), reads=wl.left.reads, unmapped=wl.left.unmapped,
                            mapped_hits=wl.left.mapped_hits, **{"seg%d" % k: wl.left.seg_hits[k] for k in range(len(wl.left.seg_hits))})
        with open(os.path.join(out, "config.json"), "w") as f:
            json.dump(dict(source="/root/reference/fusion_test/%s.fasta + testcases/test.fa" % name, options=options(), generator="scripts/make_fusion_test_golden.py",
                           binary="oracle/_ref/segment_juncs (TopHat 2.1.2, -p1), segment hits from a brute-force <= 2-mismatch mapper"), f, indent=1)
        print(name, "reads %d, unmapped %d, segment hits %s" % (wl.left.reads.shape[0], int(wl.left.unmapped.sum()), [len(s) for s in wl.left.seg_hits]),
              {k: sum(1 for _ in open(os.path.join(out, "segment." + k))) for k in ("juncs", "insertions", "deletions", "fusions")})


if __name__ == "__main__":
    main()
