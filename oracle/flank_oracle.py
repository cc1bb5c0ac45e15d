"""CPU restatement of the step BETWEEN the two hot binaries: the junction index and the segment search against it.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and bench.py's checker), never by the product path.

The reference does this with three programs (tophat.py:2546-2600, 3686-3741):
  juncs_db <min_anchor=3> <max_seg_len> juncs insertions deletions fusions ref.fa  >  segment_juncs.fa     (src/juncs_db.cpp)
  bowtie-build segment_juncs.fa ; bowtie -v <segment_mismatches> -k <max_seg_multihits> -m <max_seg_multihits> <segments>

* contigs(): the FASTA records of juncs_db, restated from juncs_db.cpp:72-229 (print_insertion, print_splice, print_fusion) and the
  driver's order 481-528 (junctions, deletions, insertions, fusions; each in its std::set order).  PINNED: tests/test_flank.py
  compares names and sequences with the output of oracle/_ref/juncs_db (the reference's own source, compiled here) on seeded sets.
* search(): bowtie is a third-party program that is not part of /root/reference (tophat.py shells out to whatever `bowtie` is on
  PATH; pinned version: none -- tophat 2.1.x asks for bowtie >= 0.12.9 / bowtie2 >= 2.0.5).  Its published -v/-k/-m contract
  (Bowtie 1 manual, "The -v alignment mode", "-k", "-m"): every end-to-end un-gapped placement with at most v mismatches is valid, an N
  in the read is a mismatch, a placement over an ambiguous reference character is invalid; a read with more than m valid
  placements reports none, otherwise with k = m all of them.  search() is that definition by exhaustion.  With bowtie2
  (TopHat's default) the segment search is heuristic (seed -N 1 -L 20, tophat.py:2299-2304) and has no closed-form result set;
  ref_n_is_mismatch=True gives the bowtie2-style treatment of reference N (a mismatch instead of an invalid placement).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

KIND_JUNC, KIND_DEL, KIND_INS, KIND_FUS = 0, 1, 2, 3
FUSION_FF, FUSION_FR, FUSION_RF, FUSION_RR = 7, 8, 9, 10       # CigarOpCode values, bwt_map.h:36-50
_C2A = np.frombuffer(b"ACGTN", dtype=np.uint8)


def _rc(codes: np.ndarray) -> np.ndarray:
    r = codes[::-1]
    return np.where(r > 3, 4, 3 - r).astype(np.uint8)


def contigs(ref_names: Sequence[str], ref_codes: Sequence[Optional[np.ndarray]], max_seg_len: int, min_anchor: int,
            junctions: np.ndarray, deletions: np.ndarray, insertions: Sequence[Tuple[int, int, str]], fusions: np.ndarray):
    """-> list of dict(name, codes, kind, ref_id, ref_id2, left_start, left, right, right_end, aux) in juncs_db's output order.

    junctions / deletions: (n, 4) [ref_id (1-based), left, right, antisense] sorted in Junction order (junctions.h:27-80);
    insertions: (ref_id, left, sequence) sorted in Insertion order; fusions: (n, 5) [ref1, ref2, left, right, dir] in Fusion order."""
    out = []

    def splice(ref_id, left, right, tag, kind, anti):
        # print_splice, juncs_db.cpp:109-149 (half_splice_len = read_len, 118-119)
        seq = ref_codes[ref_id - 1]
        if seq is None:
            return
        n = len(seq); half = max_seg_len
        if not (left <= n and right <= n):
            return
        ls = left - half + 1 if left - half + 1 >= 0 else 0
        le = ls + half
        rs = right
        re = rs + half if rs + half < n else n
        if ls < le and le <= n and rs < re and re <= n:
            out.append(dict(name="%s|%d|%d-%d|%d|%s" % (ref_names[ref_id - 1], ls, left, right, re, tag),
                            codes=np.concatenate([seq[ls:le], seq[rs:re]]).astype(np.uint8), kind=kind, ref_id=ref_id, ref_id2=ref_id,
                            left_start=ls, left=left, right=right, right_end=re, aux=anti))

    for j in np.asarray(junctions, dtype=np.int64).reshape(-1, 4):
        splice(int(j[0]), int(j[1]), int(j[2]), "GTAG|rev" if j[3] else "GTAG|fwd", KIND_JUNC, int(j[3]))
    for d in np.asarray(deletions, dtype=np.int64).reshape(-1, 4):
        splice(int(d[0]), int(d[1]), int(d[2]), "del|fwd", KIND_DEL, 0)      # read back with antisense=false (juncs_db.cpp:372)
    for (ref_id, left, s) in insertions:
        # print_insertion, juncs_db.cpp:72-104
        seq = ref_codes[ref_id - 1]
        if seq is None or "N" in s:                                          # 418-430: no ambiguity in the inserted bases
            continue
        n = len(seq); half = max_seg_len - min_anchor
        if not left <= n:
            continue
        ls = left - half + 1 if left - half + 1 >= 0 else 0
        le = ls + half
        rs = le
        re = rs + half if rs + half < n else n
        if ls < le and le <= n and rs < re and re <= n:
            ins = np.frombuffer(s.encode(), dtype=np.uint8)
            ic = np.select([ins == 65, ins == 67, ins == 71, ins == 84], [0, 1, 2, 3], 4).astype(np.uint8)
            out.append(dict(name="%s|%d|%d-%s|%d|ins|fwd" % (ref_names[ref_id - 1], ls, left, s, re),
                            codes=np.concatenate([seq[ls:le], ic, seq[rs:re]]).astype(np.uint8), kind=KIND_INS, ref_id=ref_id, ref_id2=ref_id,
                            left_start=ls, left=left, right=0, right_end=re, aux=len(s), ins=s))
    for f in np.asarray(fusions, dtype=np.int64).reshape(-1, 5):
        # print_fusion, juncs_db.cpp:151-229
        r1, r2, left, right, d = (int(x) for x in f)
        s1, s2 = ref_codes[r1 - 1], ref_codes[r2 - 1]
        if s1 is None or s2 is None:
            continue
        n1, n2 = len(s1), len(s2); half = max_seg_len - min_anchor
        if not (left < n1 and right < n2):
            continue
        if d in (FUSION_FF, FUSION_FR):
            ls = left - half + 1 if left + 1 >= half else 0
            le = ls + half
        else:
            ls = left
            le = ls + half if ls + half < n1 else n1
        if d in (FUSION_FF, FUSION_RF):
            rs = right
            re = rs + half if rs + half < n2 else n2
        else:
            re = right + 1
            rs = re - half if re >= half else 0
        if ls < le and le <= n1 and rs < re and re <= n2:
            a, b = s1[ls:le], s2[rs:re]
            name_ls, name_re = ls, re
            if d in (FUSION_RF, FUSION_RR):
                a = _rc(a); name_ls = le - 1
            if d in (FUSION_FR, FUSION_RR):
                b = _rc(b); name_re = (rs - 1) & 0xFFFFFFFFFFFFFFFF      # size_t arithmetic: 0 - 1 prints as 2^64-1 (juncs_db.cpp:214)
            out.append(dict(name="%s-%s|%d|%d-%d|%d|fus|%s" % (ref_names[r1 - 1], ref_names[r2 - 1], name_ls, left, right, name_re,
                                                              {FUSION_FF: "ff", FUSION_FR: "fr", FUSION_RF: "rf", FUSION_RR: "rr"}[d]),
                            codes=np.concatenate([a, b]).astype(np.uint8), kind=KIND_FUS, ref_id=r1, ref_id2=r2,
                            left_start=name_ls, left=left, right=right, right_end=name_re, aux=d))
    return out


def fasta(cs) -> str:
    return "".join(">%s\n%s\n" % (c["name"], _C2A[np.minimum(c["codes"], 4)].tobytes().decode()) for c in cs)


def search(contig_codes: Sequence[np.ndarray], reads: np.ndarray, seg_bounds: Sequence[int], max_mismatches: int, max_multihits: int,
           ref_n_is_mismatch: bool = False) -> np.ndarray:
    """Every placement bowtie -v/-k/-m reports: (n, 6) int64 rows [read, seg, contig, pos0, antisense, mismatches], sorted.

    reads: (n_reads, L) uint8 codes (4 = N); seg_bounds: [o_0, o_1, .., o_nseg] base offsets of the segments in the read
    (a segment is taken from the read as sequenced; antisense = its reverse complement matched)."""
    rows = []
    nseg = len(seg_bounds) - 1
    # contigs grouped by length so that the windows of one group form one array
    by_len = {}
    for ci, c in enumerate(contig_codes):
        by_len.setdefault(len(c), []).append(ci)
    groups = [(n, np.asarray(ix), np.stack([contig_codes[i] for i in ix])) for n, ix in sorted(by_len.items())]
    for k in range(nseg):
        a, b = int(seg_bounds[k]), int(seg_bounds[k + 1]); s = b - a
        for ri in range(reads.shape[0]):
            seg = reads[ri, a:b]
            found = []
            for anti, q in ((0, seg), (1, _rc(seg))):
                qn = q > 3
                for n, ix, mat in groups:
                    if n < s:
                        continue
                    win = np.lib.stride_tricks.sliding_window_view(mat, s, axis=1)        # (contigs, n-s+1, s)
                    cn = win > 3
                    mm = ((win != q) | qn | cn).sum(axis=2)
                    ok = mm <= max_mismatches
                    if not ref_n_is_mismatch:
                        ok &= ~cn.any(axis=2)
                    for c, p in zip(*np.nonzero(ok)):
                        found.append((ri, k, int(ix[c]), int(p), anti, int(mm[c, p])))
            if len(found) <= max_multihits:
                rows.extend(found)
    out = np.asarray(sorted(rows), dtype=np.int64).reshape(-1, 6)
    return out


# CigarOpCode values (bwt_map.h:36-55) and the wire flags of include/tophat_b200.h
OP_MATCH, OP_mATCH, OP_INS, OP_DEL, OP_REF_SKIP = 1, 2, 3, 5, 11
HIT_ANTISENSE, HIT_END, JHIT_ANTISENSE_SPLICE, JHIT_SEQ_FLIPPED = 1, 2, 4, 0x10


def spliced_hit(c: dict, pos: int, anti: int, mism: np.ndarray, min_anchor_len: int, last_segment: bool):
    """An un-gapped placement (len(mism) matched bases at contig offset pos; mism[o] = base o mismatches) as the BowtieHit that
    SplicedBAMHitFactory::get_hit_from_buf makes of it (bwt_map.cpp:1469-1770 with spliceCigar 678-883 for a single MATCH):
    dict(ref_id, ref_id2, left, ops [(code, len)..], flags, mismatches, splice_mms) or None where the reference discards the hit.
    PINNED end to end: tests/test_flank.py feeds the same placements to oracle/_ref/long_spanning_reads and to the join."""
    s = len(mism); nm = int(np.sum(mism))
    end = HIT_END if last_segment else 0
    if c["kind"] == KIND_INS:
        left = c["left_start"] + pos                                     # 1640-1651
        if left > c["left"]:
            return None
        at = c["left"] + 1 - left; ln = c["aux"]; ev_end = at + ln
        if s <= ev_end:                                                  # the hit ends inside or in front of the insertion: < 3 ops (874)
            return None
        smm = int(np.sum(mism[at:ev_end]))
        return dict(ref_id=c["ref_id"], ref_id2=c["ref_id"], left=left, ops=[(OP_MATCH, at), (OP_INS, ln), (OP_MATCH, s - ev_end)],
                    flags=(HIT_ANTISENSE if anti else 0) | end, mismatches=nm - smm, splice_mms=0)
    fusion = c["kind"] == KIND_FUS
    leftwards = fusion and c["aux"] in (FUSION_RF, FUSION_RR)            # 1690-1742
    left = c["left_start"] - pos if leftwards else c["left_start"] + pos
    lsp = c["left"] - 1 if leftwards else c["left"] + 1
    if (left <= lsp) if leftwards else (left >= lsp):
        return None
    at = abs(lsp - left)
    gap = c["right"] if fusion else c["right"] - c["left"] - 1
    if at >= s or gap <= 0:
        return None
    smm = int(sum(1 for o in range(s) if mism[o] and abs(at - o) < min_anchor_len))
    code = c["aux"] if fusion else (OP_DEL if c["kind"] == KIND_DEL else OP_REF_SKIP)
    before = OP_mATCH if fusion and c["aux"] in (FUSION_RF, FUSION_RR) else OP_MATCH
    after = OP_mATCH if fusion and c["aux"] in (FUSION_FR, FUSION_RR) else OP_MATCH
    anti_out = (not anti) if leftwards else bool(anti)
    flags = (HIT_ANTISENSE if anti_out else 0) | end | (JHIT_ANTISENSE_SPLICE if c["kind"] == KIND_JUNC and c["aux"] else 0) | \
            (JHIT_SEQ_FLIPPED if leftwards else 0)
    return dict(ref_id=c["ref_id"], ref_id2=c["ref_id2"], left=left, ops=[(before, at), (code, gap), (after, s - at)], flags=flags,
                mismatches=nm, splice_mms=smm)
