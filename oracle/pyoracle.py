"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/segjuncs_oracle.c and runner for the
reference's own CPU binaries built into oracle/_ref/ by oracle/Makefile.ref.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under tophat_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tophat_b200 import capi, synth  # noqa: E402  (struct mirrors only; no product compute)

BUILD_DIR = os.path.join(_HERE, "_build")
ORACLE_SO = os.path.join(BUILD_DIR, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "segjuncs_oracle.c")]
    join_src = os.path.join(_HERE, "join_oracle.c")
    if os.path.exists(join_src):
        srcs.append(join_src)
    os.makedirs(BUILD_DIR, exist_ok=True)
    newest = max(os.path.getmtime(s) for s in srcs + [os.path.join(_ROOT, "include", "tophat_b200.h")])
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < newest:
        cmd = ["gcc", "-O2", "-std=gnu11", "-shared", "-fPIC", "-Wall", "-Wno-unused-function",
               "-o", ORACLE_SO] + srcs
        subprocess.run(cmd, check=True)
    return ORACLE_SO


def build_reference() -> bool:
    """Builds oracle/_ref when the reference tree is present (this container only)."""
    if not os.path.isdir("/root/reference/src"):
        return have_reference()
    subprocess.run(["make", "-s", "-f", os.path.join("oracle", "Makefile.ref"), "-j8"], cwd=_ROOT, check=True)
    return have_reference()


def have_reference() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, p)) for p in
               ("segment_juncs", "long_spanning_reads", "prep_reads", "fix_map_ordering", "juncs_db"))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_results_new.restype = C.c_void_p
        L.orc_results_free.argtypes = [C.c_void_p]
        L.orc_segjuncs_batch.argtypes = [C.POINTER(capi.Params), C.POINTER(capi.RefImageC), C.POINTER(capi.BatchC), C.c_void_p]
        L.orc_segjuncs_finish.argtypes = [C.c_void_p]
        for n in ("orc_n_juncs", "orc_n_dels", "orc_n_ins", "orc_n_fus"):
            getattr(L, n).argtypes = [C.c_void_p]
            getattr(L, n).restype = C.c_size_t
        for n in ("orc_get_juncs", "orc_get_dels", "orc_get_ins", "orc_get_fus", "orc_get_counters", "orc_get_ins_order"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, n).restype = None
        _lib = L
    return _lib


class Counters:
    def __init__(self, a):
        self.n_windows, self.n_indel_tasks, self.n_rescue_tasks, self.n_fusion_tasks, self.n_juncs_emitted = [int(x) for x in a]


def segjuncs(params: capi.Params, ref: synth.RefImage, batches: List[synth.PackedBatch]):
    """Runs the CPU restatement over the batches in order; returns (SegJuncsResults, Counters)."""
    L = lib()
    res = L.orc_results_new()
    try:
        img = capi.ref_image_c(ref)
        for b in batches:
            bc = capi.batch_c(b)
            rc = L.orc_segjuncs_batch(C.byref(params), C.byref(img), C.byref(bc), res)
            if rc != 0:
                raise RuntimeError("oracle failed: %d" % rc)
        L.orc_segjuncs_finish(res)
        j = np.zeros(L.orc_n_juncs(res), dtype=synth.JUNCTION_DTYPE)
        d = np.zeros(L.orc_n_dels(res), dtype=synth.JUNCTION_DTYPE)
        i = np.zeros(L.orc_n_ins(res), dtype=synth.INSERTION_DTYPE)
        f = np.zeros(L.orc_n_fus(res), dtype=synth.FUSION_DTYPE)
        if j.size: L.orc_get_juncs(res, j.ctypes.data)
        if d.size: L.orc_get_dels(res, d.ctypes.data)
        if i.size: L.orc_get_ins(res, i.ctypes.data)
        if f.size: L.orc_get_fus(res, f.ctypes.data)
        cnt = np.zeros(5, dtype=np.uint64)
        L.orc_get_counters(res, cnt.ctypes.data)
        out = capi.SegJuncsResults(j, d, i, f)
        out.insertion_order = np.zeros(i.size, dtype=np.uint64)
        if i.size: L.orc_get_ins_order(res, out.insertion_order.ctypes.data)
        return out, Counters(cnt)
    finally:
        L.orc_results_free(res)


# ---------------------------------------------------------------------------------------------
# the reference's own binaries


def tophat_common_opts(inner_mean: int = 50, inner_sd: int = 20, extra: Optional[List[str]] = None) -> List[str]:
    """The option block tophat.py passes to every stage binary (TopHatParams.cmd, tophat.py:824-900)."""
    o = ["--min-anchor", "8", "--splice-mismatches", "0", "--min-report-intron", "50",
         "--max-report-intron", "500000", "--min-isoform-fraction", "0.15", "--output-dir", "./",
         "--max-multihits", "20", "--max-seg-multihits", "40", "--segment-length", "25",
         "--segment-mismatches", "2", "--min-closure-exon", "100", "--min-closure-intron", "50",
         "--max-closure-intron", "5000", "--min-coverage-intron", "50", "--max-coverage-intron", "20000",
         "--min-segment-intron", "50", "--max-segment-intron", "500000", "--read-mismatches", "2",
         "--read-gap-length", "2", "--read-edit-dist", "2", "--read-realign-edit-dist", "3",
         "--max-insertion-length", "3", "--max-deletion-length", "3", "-z", "gzip",
         "--inner-dist-mean", str(inner_mean), "--inner-dist-std-dev", str(inner_sd),
         "--no-closure-search", "--no-coverage-search", "--no-microexon-search"]
    return o + (extra or [])


def make_bams(files: Dict[str, str], outdir: str, nseg: int) -> Dict[str, str]:
    """FASTQ/SAM text -> BAM via the reference's prep_reads / fix_map_ordering (oracle/_ref); the conversions are independent
    processes and run side by side."""
    from concurrent.futures import ThreadPoolExecutor
    hdr = files["header"]
    out = {}
    cmds = []
    for side in ("left", "right"):
        kept = os.path.join(outdir, side + "_kept_reads.bam")
        cmds.append([os.path.join(REF_DIR, "prep_reads"), "--sam-header", hdr, "--outfile", kept,
                     "--index-outfile", kept + ".index", "--aux-outfile", os.path.join(outdir, side + ".info"), files[side + "_fq"]])
        out[side + "_reads"] = kept
        for key, name in [("mapped", side + "_kept_reads.mapped.bam")] + \
                         [("seg%d" % (k + 1), "%s_kept_reads_seg%d.bam" % (side, k + 1)) for k in range(nseg)]:
            bam = os.path.join(outdir, name)
            cmds.append([os.path.join(REF_DIR, "fix_map_ordering"), "--sam-header", hdr, "--index-outfile",
                         bam + ".index", files["%s_%s_sam" % (side, key)], bam])
            out["%s_%s" % (side, key)] = bam
    with ThreadPoolExecutor(max_workers=max(1, min(len(cmds), os.cpu_count() or 1))) as ex:
        for r in ex.map(lambda c: subprocess.run(c, check=True, stderr=subprocess.DEVNULL), cmds):
            pass
    return out


def run_segment_juncs(binary: str, files: Dict[str, str], bams: Dict[str, str], outdir: str, nseg: int,
                      opts: Optional[List[str]] = None, paired: bool = True, threads: int = 1,
                      env: Optional[dict] = None, tag: str = "") -> Dict[str, str]:
    outs = {k: os.path.join(outdir, "segment%s.%s" % (tag, k)) for k in ("juncs", "insertions", "deletions", "fusions")}
    cmd = [binary] + (opts if opts is not None else tophat_common_opts()) + \
          ["-p%d" % threads, "--sam-header", files["header"], "--ium-reads", "", files["fasta"],
           outs["juncs"], outs["insertions"], outs["deletions"], outs["fusions"],
           bams["left_reads"], bams["left_mapped"], ",".join(bams["left_seg%d" % (k + 1)] for k in range(nseg))]
    if paired:
        cmd += [bams["right_reads"], bams["right_mapped"], ",".join(bams["right_seg%d" % (k + 1)] for k in range(nseg))]
    log = os.path.join(outdir, "segment_juncs%s.log" % tag)
    with open(log, "w") as lf:
        subprocess.run(cmd, check=True, stderr=lf, env=env)
    outs["log"] = log
    return outs


def parse_juncs(path: str, names: List[str]) -> np.ndarray:
    idx = {n: i + 1 for i, n in enumerate(names)}
    rows = []
    with open(path) as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            rows.append((idx[t[0]], int(t[1]) & 0xFFFFFFFF, int(t[2]) & 0xFFFFFFFF, 1 if t[3] == "-" else 0))
    return np.array(rows, dtype=synth.JUNCTION_DTYPE) if rows else np.zeros(0, dtype=synth.JUNCTION_DTYPE)


def format_juncs(j: np.ndarray, names: List[str]) -> str:
    """segment.juncs text exactly as the driver writes it (segment_juncs.cpp:5041-5046)."""
    return "".join("%s\t%d\t%d\t%c\n" % (names[int(r["ref_id"]) - 1], np.int32(r["left"]), np.int32(r["right"]),
                                         "-" if r["antisense"] else "+") for r in j)


def format_deletions(d: np.ndarray, names: List[str]) -> str:
    """segment.deletions text (segment_juncs.cpp:5070-5074)."""
    return "".join("%s\t%d\t%d\n" % (names[int(r["ref_id"]) - 1], np.int32(r["left"]) + 1, np.int32(r["right"])) for r in d)


def format_insertions(i: np.ndarray, names: List[str]) -> str:
    """segment.insertions text (segment_juncs.cpp:5085-5090)."""
    return "".join("%s\t%d\t%d\t%s\n" % (names[int(r["ref_id"]) - 1], np.int32(r["left"]), np.int32(r["left"]),
                                         r["seq"].decode()) for r in i)


def format_fusions(f: np.ndarray, juncs: np.ndarray, names: List[str], resolve_conflicts: bool = True) -> str:
    """segment.fusions text: the driver's conflict resolution between neighbouring fusions (same contigs, left
    coordinates < 10 apart, same direction, same offset on both sides: keep the better supported one, ties broken by
    coincidence with a splice-junction coordinate) and its print-out (segment_juncs.cpp:5096-5180).  `f` in Fusion order."""
    coords = set()
    for j in juncs:                                                   # 5048-5052
        coords.add((int(j["ref_id"]), int(np.int32(j["left"])))); coords.add((int(j["ref_id"]), int(np.int32(j["right"]))))
    n = len(f)
    r1 = [int(x) for x in f["ref_id1"]]; r2 = [int(x) for x in f["ref_id2"]]
    left = [int(np.int32(x)) for x in f["left"]]; right = [int(np.int32(x)) for x in f["right"]]
    dr = [int(x) for x in f["dir"]]; count = [int(x) for x in f["count"]]
    lc = [(r1[i], left[i]) in coords for i in range(n)]               # 5104-5115
    rc = [(r2[i], right[i]) in coords for i in range(n)]
    skip = [False] * n
    out = []
    for i in range(n):
        k = i + 1
        while k < n:                                                  # 5124-5155
            ld = abs(left[i] - left[k])
            if r1[i] == r1[k] and r2[i] == r2[k] and ld < 10:
                if dr[i] == dr[k] and ld == abs(right[i] - right[k]):
                    if count[k] > count[i]:
                        skip[i] = True
                    elif count[k] == count[i]:
                        if int(lc[i]) + int(rc[i]) < int(lc[k]) + int(rc[k]):
                            skip[i] = True
                        else:
                            skip[k] = True
                    else:
                        skip[k] = True
                k += 1
            else:
                break
        if skip[i] and resolve_conflicts:                             # 5157
            continue
        d = {8: "fr", 9: "rf", 10: "rr"}.get(dr[i], "ff")             # 5163-5171
        out.append("%s\t%d\t%s\t%d\t%s\n" % (names[r1[i] - 1], left[i], names[r2[i] - 1], right[i], d))
    return "".join(out)


# ---------------------------------------------------------------------------------------------
# long_spanning_reads: junction index + spliced segment hits + the reference binary


def make_join_inputs(wl, files: Dict[str, str], outs: Dict[str, str], outdir: str, nseg: int, max_seg_len: int = 26, fast: bool = False) -> Dict[str, str]:
    """juncs_db (reference binary) on the segment_juncs outputs, then synthetic segment hits against its contigs,
    turned into id-sorted BAMs by the reference's fix_map_ordering (tophat.py:3686-3741)."""
    fa = os.path.join(outdir, "segment_juncs.fa")
    with open(fa, "w") as f:
        subprocess.run([os.path.join(REF_DIR, "juncs_db"), "3", str(max_seg_len), outs["juncs"], outs["insertions"], outs["deletions"],
                        "/dev/null", files["fasta"]], check=True, stdout=f, stderr=subprocess.DEVNULL)
    contigs = synth.parse_juncs_db_fasta(fa)
    hdr = os.path.join(outdir, "segment_juncs.hdr.sam")
    synth.write_contig_header(hdr, contigs)
    j = {"juncs_fa": fa, "juncs_header": hdr, "n_contigs": len(contigs)}
    junctions = parse_juncs(outs["juncs"], wl.ref.names)
    for sname, side in (("left", wl.left), ("right", wl.right)):
        per_seg = synth.spliced_hits_for_sam(wl, side, junctions, contigs) if fast else synth.spliced_segment_hits(wl, side, contigs)
        for k in range(nseg):
            sam = os.path.join(outdir, "%s_seg%d.to_spliced.sam" % (sname, k + 1))
            synth.write_spliced_sam(sam, wl.cfg, k, per_seg[k])
            bam = os.path.join(outdir, "%s_kept_reads_seg%d.to_spliced.bam" % (sname, k + 1))
            subprocess.run([os.path.join(REF_DIR, "fix_map_ordering"), "--sam-header", hdr, "--index-outfile", bam + ".index", sam, bam],
                           check=True, stderr=subprocess.DEVNULL)
            j["%s_spl%d" % (sname, k + 1)] = bam
        j["%s_n_spliced" % sname] = sum(len(x) for x in per_seg)
    return j


def run_long_spanning_reads(binary: str, files: Dict[str, str], bams: Dict[str, str], jin: Dict[str, str], outs: Dict[str, str], outdir: str,
                            nseg: int, side: str = "left", opts: Optional[List[str]] = None, tag: str = "", env: Optional[dict] = None,
                            with_spliced: bool = True, threads: int = 1, fusions: str = "/dev/null") -> str:
    out_bam = os.path.join(outdir, "%s_candidates%s.bam" % (side, tag))
    cmd = [binary] + (opts if opts is not None else tophat_common_opts()) + \
          ["-p%d" % threads, "--sam-header", files["header"], "--bowtie2-max-penalty", "6", "--bowtie2-min-penalty", "2", "--bowtie2-penalty-for-N", "1",
           "--bowtie2-read-gap-open", "5", "--bowtie2-read-gap-cont", "3", "--bowtie2-ref-gap-open", "5", "--bowtie2-ref-gap-cont", "3",
           files["fasta"], bams[side + "_reads"], outs["juncs"], outs["insertions"], outs["deletions"], fusions, out_bam,
           ",".join(bams["%s_seg%d" % (side, k + 1)] for k in range(nseg))]
    if with_spliced:
        cmd.append(",".join(jin["%s_spl%d" % (side, k + 1)] for k in range(nseg)))
    with open(os.path.join(outdir, "long_spanning_reads%s.%s.log" % (tag, side)), "w") as lf:
        subprocess.run(cmd, check=True, stderr=lf, env=env)
    return out_bam


def read_bam(path: str):
    """Decoded BAM records (BGZF is multi-member gzip): (qname, flag, tid, pos, mapq, cigar, seq, qual, mtid, mpos, tlen, aux dict).
    A stage run with -p N leaves <path minus .bam>0.bam .. N-1.bam instead of <path> (long_spanning_reads.cpp:3056-3064): those are
    read in order and concatenated."""
    import gzip
    import struct
    if not os.path.exists(path) and os.path.exists(path[:-4] + "0.bam"):
        refs, recs, i = None, [], 0
        while os.path.exists(path[:-4] + "%d.bam" % i):
            r, rec = read_bam(path[:-4] + "%d.bam" % i)
            refs = refs or r; recs += rec; i += 1
        return refs, recs
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", data, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", data, off)[0]; off += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", data, off)[0]; off += 4
        refs.append(data[off:off + ln - 1].decode()); off += ln + 4
    recs = []
    while off < len(data):
        bs = struct.unpack_from("<i", data, off)[0]; p = off + 4; off = p + bs
        tid, pos, l_qn, mapq, _bin, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", data, p)
        q = p + 32
        qname = data[q:q + l_qn - 1].decode(); q += l_qn
        cig = "".join("%d%s" % (c >> 4, "MIDNSHP=X"[c & 15]) for c in struct.unpack_from("<%dI" % n_cig, data, q)); q += 4 * n_cig
        sb = data[q:q + (l_seq + 1) // 2]; q += (l_seq + 1) // 2
        seq = "".join("=ACMGRSVTWYHKDBN"[(sb[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        qual = bytes(data[q:q + l_seq]); q += l_seq
        aux = {}
        while q < off:
            tag = data[q:q + 2].decode(); t = chr(data[q + 2]); q += 3
            if t == "A": aux[tag] = chr(data[q]); q += 1
            elif t in "cCsSiI":
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[t]
                aux[tag] = struct.unpack_from(fmt, data, q)[0]; q += struct.calcsize(fmt)
            elif t == "Z":
                e = data.index(b"\x00", q); aux[tag] = data[q:e].decode(); q = e + 1
            else:
                raise ValueError("aux type %s" % t)
        recs.append((qname, flag, refs[tid] if tid >= 0 else "*", pos, mapq, cig, seq, qual, mtid, mpos, tlen, aux))
    return refs, recs
