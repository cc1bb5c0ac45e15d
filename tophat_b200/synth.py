"""Synthetic RNA-seq workload generator for the segment_juncs / long_spanning_reads hot path.

Implements the generator spec of SURVEY.md section 8(d): random reference contigs with planted
GT-AG / GC-AG / AT-AC introns, 2x101 bp reads sampled from the spliced (exonic) coordinate space,
substitutions / Ns / optional 1-3 bp indels, and *analytic segment placement*: every 25/25/25/26 bp
segment is placed ungapped at the genomic position of its first and of its last base and kept as a
segment hit when it has <= 2 mismatches there -- the rule an ungapped `-v 2`-style mapper applies
(reference behaviour studied in src/tophat.py:2878-2990 split_reads and
src/bwt_map.cpp:1101-1452 BAMHitFactory).

Everything is numpy-vectorised so that the 10 M-pair configuration is generated in chunks.  Two
consumers:
  * `pack_side()` builds the packed C-ABI batch (include/tophat_b200.h) directly, reproducing the
    bundle rules of look_for_hit_group / process_next_hit_group (segment_juncs.cpp:3823-4123);
  * `write_pipeline_files()` writes FASTA / FASTQ / SAM text so that the reference's own
    prep_reads + fix_map_ordering binaries (oracle/_ref) turn them into the BAM inputs both the
    reference binaries and our host binaries consume.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------
# C-ABI record dtypes (include/tophat_b200.h)

HIT_DTYPE = np.dtype([("ref_id", "<u4"), ("left", "<i4"), ("right", "<i4"), ("read_len", "u1"),
                      ("edit_dist", "u1"), ("flags", "u1"), ("reserved", "u1")])
BUNDLE_DTYPE = np.dtype([("read_id", "<u4"), ("hit_begin", "<u4"), ("partner_begin", "<u4"),
                         ("n_partner", "<u2"), ("read_len", "u1"), ("flags", "u1")])
JUNCTION_DTYPE = np.dtype([("ref_id", "<u4"), ("left", "<u4"), ("right", "<u4"), ("antisense", "<u4")])
INSERTION_DTYPE = np.dtype([("ref_id", "<u4"), ("left", "<u4"), ("len", "<u4"), ("seq", "S20")])
FUSION_DTYPE = np.dtype([("ref_id1", "<u4"), ("ref_id2", "<u4"), ("left", "<u4"), ("right", "<u4"),
                         ("dir", "<u4"), ("count", "<u4"), ("edit_dist", "<u4"), ("reserved", "<u4")])
assert HIT_DTYPE.itemsize == 16 and BUNDLE_DTYPE.itemsize == 16

HIT_ANTISENSE, HIT_END = 1, 2
B_INDELS, B_GAPS, B_FUSIONS, B_FUSIONS_LAST, B_RIGHT_MATE = 1, 2, 4, 8, 16

CODE2CHAR = np.frombuffer(b"ACGTN", dtype=np.uint8)
_CHAR2CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CHAR2CODE[_c] = _i
    _CHAR2CODE[ord(chr(_c).lower())] = _i
_CHAR2CODE[ord("U")] = 3
_CHAR2CODE[ord("u")] = 3


def codes_from_ascii(b: bytes) -> np.ndarray:
    return _CHAR2CODE[np.frombuffer(b, dtype=np.uint8)]


def revcomp_codes(a: np.ndarray) -> np.ndarray:
    """Reverse-complement along the last axis; code 4 (N) stays N (reads.cpp:191-207)."""
    r = a[..., ::-1]
    return np.where(r < 4, 3 - r, 4).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# reference image (bit planes)


@dataclasses.dataclass
class RefImage:
    names: List[str]
    contig_len: np.ndarray      # u4 [n]
    contig_start: np.ndarray    # u8 [n], multiples of 64
    planes: np.ndarray          # u8 [2*n_blocks]
    nmask: np.ndarray           # u8 [n_blocks]
    codes: Optional[List[np.ndarray]] = None   # per-contig uint8 codes (0..3, 4 = N)

    @property
    def n_blocks(self) -> int:
        return int(self.nmask.shape[0])


def pack_planes(codes: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """codes (uint8, 0..4) -> (plane0, plane1, planeN) little-bit-order uint64 words."""
    n = codes.shape[0]
    nw = (n + 63) // 64
    pad = nw * 64 - n
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]) if pad else codes
    isn = c > 3
    c = np.where(isn, 0, c)
    p0 = np.packbits((c & 1).astype(np.uint8), bitorder="little").view("<u8")
    p1 = np.packbits(((c >> 1) & 1).astype(np.uint8), bitorder="little").view("<u8")
    pn = np.packbits(isn.astype(np.uint8), bitorder="little").view("<u8")
    return p0, p1, pn


def build_ref_image(names: List[str], contigs: List[Optional[np.ndarray]]) -> RefImage:
    starts, lens = [], []
    g = 0
    for c in contigs:
        starts.append(g)
        n = 0 if c is None else int(c.shape[0])
        lens.append(n)
        g += ((n + 63) // 64 + 1) * 64      # >= 64 bases of zero padding after every contig
    nb = g // 64 + 1
    planes = np.zeros(2 * nb, dtype="<u8")
    nmask = np.zeros(nb, dtype="<u8")
    for s, c in zip(starts, contigs):
        if c is None or c.shape[0] == 0:
            continue
        p0, p1, pn = pack_planes(c)
        b0 = s // 64
        planes[2 * b0: 2 * (b0 + p0.shape[0]): 2] = p0
        planes[2 * b0 + 1: 2 * (b0 + p0.shape[0]) + 1: 2] = p1
        nmask[b0: b0 + pn.shape[0]] = pn
    return RefImage(list(names), np.asarray(lens, dtype="<u4"), np.asarray(starts, dtype="<u8"),
                    planes, nmask, [None if c is None else c for c in contigs])


def pack_reads(codes: np.ndarray, read_words: int) -> np.ndarray:
    """codes (n, L) uint8 -> (n, 3*read_words) uint64: plane0 | plane1 | planeN."""
    n, L = codes.shape
    W = read_words * 64
    buf = np.zeros((n, W), dtype=np.uint8)
    isn = codes > 3
    c = np.where(isn, 0, codes)
    out = np.empty((n, 3 * read_words), dtype="<u8")
    buf[:, :L] = c & 1
    out[:, 0:read_words] = np.packbits(buf, axis=1, bitorder="little").view("<u8")
    buf[:, :L] = (c >> 1) & 1
    out[:, read_words:2 * read_words] = np.packbits(buf, axis=1, bitorder="little").view("<u8")
    buf[:, :L] = isn
    out[:, 2 * read_words:] = np.packbits(buf, axis=1, bitorder="little").view("<u8")
    return out


# ---------------------------------------------------------------------------------------------
# workload description


@dataclasses.dataclass
class SynthConfig:
    contig_lens: Tuple[int, ...] = (2_000_000,)
    n_pairs: int = 20_000
    seed: int = 20240611
    read_len: int = 101
    segment_length: int = 25
    sub_rate: float = 0.005
    n_rate: float = 0.0005
    ref_n_frac: float = 0.005          # fraction of reference bases inside N runs
    indel_prob: float = 0.0            # per mate (config 4: 0.5)
    decoy_rate: float = 0.3            # Poisson mean of decoy multihits per segment
    inner_mean: float = 50.0
    inner_sd: float = 20.0
    exon_mu: float = 5.0
    exon_sigma: float = 0.6
    intron_mu: float = 7.3
    intron_sigma: float = 1.3
    intron_max: int = 400_000
    fusion_frac: float = 0.0           # fraction of fragments that are chimeric (config 5 style)
    chunk: int = 500_000
    chunk_seed_base: int = 0           # added to the chunk index in the RNG stream id (per-rank shards)
    keep_truth: bool = False           # keep the per-base genomic origin of every read (needed to place spliced segment hits)
    keep_candidates: bool = False      # keep compact junction-spanning segment placements instead (large workloads)


@dataclasses.dataclass
class SideData:
    """Everything the pipeline knows about one mate side (left = mate 1, right = mate 2)."""
    reads: np.ndarray                   # (n, L) codes, read orientation
    ids: np.ndarray                     # (n,) u4 read ids (1-based, shared by mates)
    seg_hits: List[np.ndarray]          # per segment: structured array (read_idx + HIT fields)
    mapped_hits: np.ndarray             # full-read hits (the *.mapped.bam stream)
    unmapped: np.ndarray                # bool (n,): read went to segment mapping
    cand: Optional[list] = None         # keep_candidates: per segment dict of arrays (see _placement_candidates)
    truth: Optional[dict] = None        # keep_truth: fwd (n,L) codes in genome orientation, gpos (n,L) genomic position of
                                        # every base (-1 = inserted), ref_id (n,), rev (n,) read is the reverse complement of fwd


SEGHIT_DTYPE = np.dtype([("read_idx", "<u4"), ("ref_id", "<u4"), ("left", "<i4"), ("right", "<i4"),
                         ("read_len", "u1"), ("edit_dist", "u1"), ("flags", "u1"), ("pad", "u1")])


@dataclasses.dataclass
class Workload:
    cfg: SynthConfig
    ref: RefImage
    left: SideData
    right: SideData
    introns: np.ndarray                 # planted introns: (ref_id, start, end, minus_strand)


def segment_layout(read_len: int, seglen: int) -> Tuple[np.ndarray, np.ndarray]:
    """Offsets/lengths of split_reads (tophat.py:2975-2984): last segment absorbs the remainder."""
    nseg = max(1, read_len // seglen)
    offs = np.arange(nseg) * seglen
    lens = np.full(nseg, seglen)
    lens[-1] = read_len - offs[-1]
    return offs, lens


# ---------------------------------------------------------------------------------------------
# reference + annotation


def _make_contig(rng: np.random.Generator, n: int, cfg: SynthConfig, ref_id: int):
    seq = rng.integers(0, 4, size=n, dtype=np.uint8)
    # alternating exon / intron layout over the whole contig
    est = max(16, int(n / 400) + 16)
    ex = np.clip(rng.lognormal(cfg.exon_mu, cfg.exon_sigma, est * 4), 30, 2000).astype(np.int64)
    it = np.clip(rng.lognormal(cfg.intron_mu, cfg.intron_sigma, est * 4), 70, cfg.intron_max).astype(np.int64)
    # vectorised layout: exon_k = [p_k, p_k + ex_k), intron_k = [p_k + ex_k, p_{k+1})
    step = ex + it
    p = 600 + np.concatenate([[0], np.cumsum(step)[:-1]])      # keep clear of the contig start
    nex = int(np.searchsorted(p + ex + 800, n, side="left"))    # exons that end >= 800 before the end
    nex = max(nex, 1)
    exons = np.stack([p[:nex], p[:nex] + ex[:nex]], axis=1)
    introns = np.stack([p[:nex - 1] + ex[:nex - 1], p[1:nex]], axis=1) if nex > 1 else np.zeros((0, 2), np.int64)
    ni = introns.shape[0]
    minus = rng.random(ni) < 0.5
    u = rng.random(ni)
    motif = np.where(u < 0.987, 0, np.where(u < 0.997, 1, 2))       # GT-AG, GC-AG, AT-AC
    don = np.array([[2, 3], [2, 1], [0, 3]], dtype=np.uint8)          # GT GC AT
    acc = np.array([[0, 2], [0, 2], [0, 1]], dtype=np.uint8)          # AG AG AC
    if ni:
        s0, e0 = introns[:, 0], introns[:, 1]
        d, a = don[motif], acc[motif]                # (ni, 2)
        first = np.where(minus[:, None], (3 - a)[:, ::-1], d)      # reverse strand: rc(acceptor) ... rc(donor)
        lastp = np.where(minus[:, None], (3 - d)[:, ::-1], a)
        seq[s0] = first[:, 0]; seq[s0 + 1] = first[:, 1]
        seq[e0 - 2] = lastp[:, 0]; seq[e0 - 1] = lastp[:, 1]
    # N runs (never over a planted motif's 2 bases is not required; reads simply inherit them)
    n_target = int(cfg.ref_n_frac * n)
    filled = 0
    while filled < n_target:
        ln = int(rng.integers(50, 500))
        p = int(rng.integers(0, max(1, n - ln)))
        seq[p:p + ln] = 4
        filled += ln
    ann = np.zeros(ni, dtype=[("ref_id", "<u4"), ("start", "<i8"), ("end", "<i8"), ("minus", "?")])
    ann["ref_id"] = ref_id
    ann["start"] = introns[:, 0] if ni else 0
    ann["end"] = introns[:, 1] if ni else 0
    ann["minus"] = minus
    return seq, exons, ann


class _ExonSpace:
    """Concatenated exonic coordinate space of all contigs."""

    def __init__(self, exons_per_contig: List[np.ndarray]):
        gs, ln, rid = [], [], []
        self.contig_first = []
        for ci, ex in enumerate(exons_per_contig):
            self.contig_first.append(sum(len(x) for x in gs))
            gs.append(ex[:, 0])
            ln.append(ex[:, 1] - ex[:, 0])
            rid.append(np.full(ex.shape[0], ci + 1, dtype=np.int64))
        self.gstart = np.concatenate(gs)
        self.len = np.concatenate(ln)
        self.ref_id = np.concatenate(rid)
        self.cum = np.concatenate([[0], np.cumsum(self.len)])       # exonic offset of each exon
        # per-contig exonic ranges
        self.contig_lo = np.array([self.cum[f] for f in self.contig_first], dtype=np.int64)
        nxt = self.contig_first[1:] + [self.gstart.shape[0]]
        self.contig_hi = np.array([self.cum[f] for f in nxt], dtype=np.int64)

    def exon_of(self, x: np.ndarray) -> np.ndarray:
        return np.searchsorted(self.cum, x, side="right") - 1

    def to_genome(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        e = self.exon_of(x)
        return self.ref_id[e], self.gstart[e] + (x - self.cum[e]), e


def _gather_ref(ref_codes: List[np.ndarray], ref_id: np.ndarray, pos: np.ndarray, n: int) -> np.ndarray:
    """ref bases [pos, pos+n) for every row (rows may sit on different contigs); out-of-range -> 5."""
    out = np.full((pos.shape[0], n), 5, dtype=np.uint8)
    for ci, codes in enumerate(ref_codes):
        sel = np.nonzero(ref_id == ci + 1)[0]
        if sel.size == 0:
            continue
        idx = pos[sel, None] + np.arange(n)[None, :]
        ok = (idx >= 0) & (idx < codes.shape[0])
        vals = codes[np.clip(idx, 0, codes.shape[0] - 1)]
        out[sel] = np.where(ok, vals, 5)
    return out


# ---------------------------------------------------------------------------------------------
# reads + analytic segment placement


def _sample_mates(rng, cfg: SynthConfig, space: _ExonSpace, ref_codes, n: int):
    """Returns forward-genome-orientation data for n fragments: for both mates the exonic start,
    forward read codes (after errors / indels), strand, and per-base exonic->read maps needed to
    place segments."""
    L = cfg.read_len
    # choose a contig proportional to its exonic length, then a fragment inside it
    w = (space.contig_hi - space.contig_lo).astype(np.float64)
    ci = rng.choice(len(w), size=n, p=w / w.sum())
    inner = np.maximum(-L + 10, np.rint(rng.normal(cfg.inner_mean, cfg.inner_sd, n))).astype(np.int64)
    flen = 2 * L + inner
    span = space.contig_hi[ci] - space.contig_lo[ci] - flen - 8
    span = np.maximum(span, 1)
    f0 = space.contig_lo[ci] + (rng.random(n) * span).astype(np.int64)
    plus = rng.random(n) < 0.5                  # fragment strand
    a0 = f0                                     # exonic start of the upstream (forward) read
    b0 = f0 + flen - L                          # exonic start of the downstream (reverse) read
    return a0, b0, plus


def _make_read_fwd(rng, cfg: SynthConfig, space: _ExonSpace, exonic: np.ndarray, x0: np.ndarray):
    """Forward-orientation read of length L starting at exonic offset x0, with substitutions, Ns and
    (optionally) one 1-3 bp indel.  Returns (codes (n,L), gmap (n,L) exonic coordinate each read
    base came from, or -1 for inserted bases)."""
    n = x0.shape[0]
    L = cfg.read_len
    ar = np.arange(L)[None, :]
    emap = x0[:, None] + ar                      # exonic coordinate of each read base
    if cfg.indel_prob > 0:
        has = rng.random(n) < cfg.indel_prob
        is_del = rng.random(n) < 0.5
        ln = rng.integers(1, 4, size=n)
        near = rng.random(n) < 0.6
        seg = cfg.segment_length
        bnd = rng.choice(np.array([seg, 2 * seg]), size=n) + rng.integers(-2, 3, size=n)
        anyw = rng.integers(5, L - 8, size=n)
        p = np.where(near, bnd, anyw)            # read offset of the first base after the event
        # deletion: read bases >= p come from exonic + ln ; insertion: ln inserted bases at p
        dsel = has & is_del
        isel = has & ~is_del
        shift = np.zeros((n, L), dtype=np.int64)
        shift[dsel] = np.where(ar >= p[dsel, None], ln[dsel, None], 0)
        ins_zone = isel[:, None] & (ar >= p[:, None]) & (ar < (p + ln)[:, None])
        shift[isel] = np.where(ar >= (p + ln)[isel, None], -ln[isel, None], 0)
        emap = emap + shift
        emap[ins_zone] = -1
    inv = None
    if cfg.fusion_frac > 0:
        # chimeric fragments (config 5): the read's tail, from breakpoint p on, comes from a second exonic locus,
        # in the same orientation (ff / rr fusions) or reverse-complemented (fr / rf)
        chim = rng.random(n) < cfg.fusion_frac
        bp = rng.integers(22, L - 21, size=n)
        x1 = (rng.random(n) * (exonic.shape[0] - 2 * L)).astype(np.int64) + L
        kind = rng.integers(0, 4, size=n)                  # 0, 1: same orientation; 2: tail inverted; 3: head inverted
        tail = chim[:, None] & (ar >= bp[:, None])
        head = chim[:, None] & (ar < bp[:, None]) & (kind == 3)[:, None]
        fwd_tail = x1[:, None] + (ar - bp[:, None])
        rev_tail = x1[:, None] + (L - 1 - ar)
        rev_head = x1[:, None] + (bp[:, None] - 1 - ar)
        emap = np.where(tail & (kind != 3)[:, None], np.where((kind == 2)[:, None], rev_tail, fwd_tail), emap)
        emap = np.where(head, rev_head, emap)
        inv = (tail & (kind == 2)[:, None]) | head
    codes = exonic[np.clip(emap, 0, exonic.shape[0] - 1)]
    if inv is not None:
        codes = np.where(inv & (codes < 4), 3 - codes, codes)
    inserted = emap < 0
    if inserted.any():
        codes = np.where(inserted, rng.integers(0, 4, size=codes.shape, dtype=np.uint8), codes)
    sub = rng.random(codes.shape) < cfg.sub_rate
    codes = np.where(sub & (codes < 4), (codes + rng.integers(1, 4, size=codes.shape, dtype=np.uint8)) % 4, codes)
    nn = rng.random(codes.shape) < cfg.n_rate
    codes = np.where(nn, 4, codes).astype(np.uint8)
    if inv is not None:
        return codes, emap, inv, chim
    return codes, emap


def _place(ref_codes, space: _ExonSpace, codes_fwd: np.ndarray, emap: np.ndarray, lo: int, ln: int,
           max_mm: int):
    """Ungapped placement of read_fwd[lo:lo+ln] anchored at its first and at its last base.
    Returns list of (row_idx, ref_id, left, mismatches) arrays (first-anchor, last-anchor)."""
    seg = codes_fwd[:, lo:lo + ln]
    res = []
    for anchor in (0, ln - 1):
        ex = emap[:, lo + anchor]
        ok = ex >= 0
        rid, gpos, _ = space.to_genome(np.where(ok, ex, 0))
        left = gpos - anchor
        refb = _gather_ref(ref_codes, rid, left, ln)
        mm = ((seg != refb) | (seg > 3) | (refb > 3)).sum(axis=1)
        good = ok & (mm <= max_mm) & (refb != 5).all(axis=1)
        res.append((good, rid, left, mm))
    g1, r1, l1, m1 = res[0]
    g2, r2, l2, m2 = res[1]
    same = g1 & g2 & (r1 == r2) & (l1 == l2)
    g2 = g2 & ~same
    return (g1, r1, l1, m1), (g2, r2, l2, m2)


_GEN_STATE = {}


def _gen_chunk(ci: int):
    """One chunk of fragments (independent RNG stream per chunk so chunks can run in any process)."""
    st = _GEN_STATE
    cfg, contigs, space, exonic = st["cfg"], st["contigs"], st["space"], st["exonic"]
    rng = np.random.default_rng(np.random.SeedSequence([cfg.seed, cfg.chunk_seed_base + ci + 1]))
    L = cfg.read_len
    offs, lens = segment_layout(L, cfg.segment_length)
    nseg = offs.shape[0]
    done = ci * cfg.chunk
    n = min(cfg.chunk, cfg.n_pairs - done)
    acc = {s: dict(reads=[], seg=[[] for _ in range(nseg)], mapped=[], unm=[], truth=[], cand=[]) for s in ("left", "right")}
    a0, b0, plus = _sample_mates(rng, cfg, space, contigs, n)
    for which, x0 in (("A", a0), ("B", b0)):
        made = _make_read_fwd(rng, cfg, space, exonic, x0)
        codes_fwd, emap = made[0], made[1]
        inv = made[2] if len(made) > 2 else None
        chim = made[3] if len(made) > 3 else None
        # A is sequenced forward, B reverse-complemented (fr library); fragment strand decides
        # which one is mate 1.
        rev = which == "B"
        # per-row side assignment: plus fragments: A=left,B=right ; minus: B=left, A=right
        is_left = plus if which == "A" else ~plus
        # whole-read contiguous mapping (the *.mapped.bam stream): <= 2 mismatches end to end
        (g1, r1, l1, m1), _ = _place(contigs, space, codes_fwd, emap, 0, L, 2)
        mapped = g1 & (emap[:, L - 1] - emap[:, 0] == L - 1)
        # re-check contiguity on the genome: first-anchor placement must also fit the last base
        rid_last, g_last, _ = space.to_genome(np.maximum(emap[:, L - 1], 0))
        mapped &= (rid_last == r1) & (g_last == l1 + L - 1)
        read_codes = revcomp_codes(codes_fwd) if rev else codes_fwd
        for side, selmask in (("left", is_left), ("right", ~is_left)):
            sel = np.nonzero(selmask)[0]
            if sel.size == 0:
                continue
            base = done  # read index within side == pair index
            A = acc[side]
            A["reads"].append((sel + base, read_codes[sel]))
            mh = np.zeros(int(mapped[sel].sum()), dtype=SEGHIT_DTYPE)
            ms = sel[mapped[sel]]
            mh["read_idx"] = ms + base
            mh["ref_id"] = r1[ms]; mh["left"] = l1[ms]; mh["right"] = l1[ms] + L
            mh["read_len"] = L; mh["edit_dist"] = m1[ms]
            mh["flags"] = (HIT_ANTISENSE if rev else 0) | HIT_END
            A["mapped"].append(mh)
            A["unm"].append((sel + base, ~mapped[sel]))
            if cfg.keep_candidates:
                us = sel[~mapped[sel]] if chim is None else sel[~mapped[sel] & ~chim[sel]]     # chimeric reads span two loci: no single-locus candidates
                if us.size:
                    oku = emap[us] >= 0
                    rid_u, gpos_u, _ = space.to_genome(np.where(oku, emap[us], 0))
                    A["cand"].append(_placement_candidates(cfg, contigs, codes_fwd[us], np.where(oku, gpos_u, -1), rid_u[:, 0],
                                                           np.full(us.size, rev), us + base))
            if cfg.keep_truth:
                ok = emap[sel] >= 0
                rid_t, gpos_t, _ = space.to_genome(np.where(ok, emap[sel], 0))
                gpos_t = np.where(ok, gpos_t, -1)
                A["truth"].append((sel + base, codes_fwd[sel], gpos_t, rid_t[:, 0], np.full(sel.size, rev)))
        # segment hits for reads that did not map end to end
        for k in range(nseg):
            # segment k of the *sequenced* read covers, in forward orientation:
            lo = int(L - offs[k] - lens[k]) if rev else int(offs[k])
            ln = int(lens[k])
            for (g, r, l, m) in _place(contigs, space, codes_fwd, emap, lo, ln, 2):
                keep = g & ~mapped
                rows = np.nonzero(keep)[0]
                if rows.size == 0:
                    continue
                sh = np.zeros(rows.size, dtype=SEGHIT_DTYPE)
                sh["read_idx"] = rows + done
                sh["ref_id"] = r[rows]; sh["left"] = l[rows]; sh["right"] = l[rows] + ln
                sh["read_len"] = ln; sh["edit_dist"] = m[rows]
                sh["flags"] = (HIT_ANTISENSE if rev else 0) | (HIT_END if k == nseg - 1 else 0)
                for side, selmask in (("left", is_left), ("right", ~is_left)):
                    part = sh[selmask[rows]]
                    if part.size:
                        acc[side]["seg"][k].append(part)
            if inv is not None:
                # segments lying entirely inside a reverse-complemented tail map to the opposite strand
                rows = np.nonzero(inv[:, lo] & inv[:, lo + ln - 1] & ~mapped)[0]
                if rows.size:
                    rid_a, g_a, _ = space.to_genome(emap[rows, lo + ln - 1])
                    rid_b, g_b, _ = space.to_genome(emap[rows, lo])
                    refb = _gather_ref(contigs, rid_a, g_a, ln)
                    seg_rc = revcomp_codes(codes_fwd[rows, lo:lo + ln])
                    mm = ((seg_rc != refb) | (seg_rc > 3) | (refb > 3)).sum(axis=1)
                    good = (rid_a == rid_b) & (g_b == g_a + ln - 1) & (mm <= 2) & (refb != 5).all(axis=1)
                    rows, rid_a, g_a, mm = rows[good], rid_a[good], g_a[good], mm[good]
                    sh = np.zeros(rows.size, dtype=SEGHIT_DTYPE)
                    sh["read_idx"] = rows + done
                    sh["ref_id"] = rid_a; sh["left"] = g_a; sh["right"] = g_a + ln
                    sh["read_len"] = ln; sh["edit_dist"] = mm
                    sh["flags"] = (0 if rev else HIT_ANTISENSE) | (HIT_END if k == nseg - 1 else 0)
                    for side, selmask in (("left", is_left), ("right", ~is_left)):
                        part = sh[selmask[rows]]
                        if part.size:
                            acc[side]["seg"][k].append(part)
            # decoy multihits at random positions
            if cfg.decoy_rate > 0:
                nd = rng.poisson(cfg.decoy_rate, size=n)
                nd[mapped] = 0
                rows = np.repeat(np.arange(n), nd)
                if rows.size:
                    dh = np.zeros(rows.size, dtype=SEGHIT_DTYPE)
                    cidx = rng.integers(0, len(contigs), size=rows.size)
                    clen = np.asarray([c.shape[0] for c in contigs])[cidx]
                    dh["read_idx"] = rows + done
                    dh["ref_id"] = cidx + 1
                    dl = (rng.random(rows.size) * (clen - 1400)).astype(np.int64) + 700
                    dh["left"] = dl; dh["right"] = dl + int(lens[k])
                    dh["read_len"] = int(lens[k]); dh["edit_dist"] = rng.integers(0, 3, size=rows.size)
                    dh["flags"] = (rng.integers(0, 2, size=rows.size) * HIT_ANTISENSE).astype(np.uint8) | \
                        (HIT_END if k == nseg - 1 else 0)
                    for side, selmask in (("left", is_left), ("right", ~is_left)):
                        part = dh[selmask[rows]]
                        if part.size:
                            acc[side]["seg"][k].append(part)
    return acc


@dataclasses.dataclass
class RefData:
    """The synthetic genome + its annotation: what generate() needs besides the read parameters.  Depends only on
    (contig_lens, seed, ref_n_frac, exon / intron distribution), so ranks of one box can share one copy (save_refdata /
    load_refdata: arrays memory-mapped from /dev/shm)."""
    names: List[str]
    contigs: List[np.ndarray]
    exons_pc: List[np.ndarray]
    introns: np.ndarray
    ref: RefImage


def make_reference(cfg: SynthConfig) -> RefData:
    rng = np.random.default_rng(np.random.PCG64(cfg.seed))
    names, contigs, exons_pc, anns = [], [], [], []
    for ci, n in enumerate(cfg.contig_lens):
        seq, ex, ann = _make_contig(rng, int(n), cfg, ci + 1)
        names.append("chrS%d" % (ci + 1))
        contigs.append(seq)
        exons_pc.append(ex)
        anns.append(ann)
    ref = build_ref_image(names, contigs)
    introns = np.concatenate(anns) if anns else np.zeros(0)
    return RefData(names, contigs, exons_pc, introns, ref)


def save_refdata(rd: RefData, dirname: str) -> None:
    os.makedirs(dirname, exist_ok=True)
    for i, (c, e) in enumerate(zip(rd.contigs, rd.exons_pc)):
        np.save(os.path.join(dirname, "contig%d.npy" % i), c); np.save(os.path.join(dirname, "exons%d.npy" % i), e)
    np.save(os.path.join(dirname, "introns.npy"), rd.introns)
    for k in ("contig_start", "contig_len", "planes", "nmask"):
        np.save(os.path.join(dirname, "ref_%s.npy" % k), getattr(rd.ref, k))
    with open(os.path.join(dirname, "names.txt"), "w") as f:
        f.write("\n".join(rd.names) + "\n")


def load_refdata(dirname: str) -> RefData:
    names = open(os.path.join(dirname, "names.txt")).read().split()
    mm = lambda n: np.load(os.path.join(dirname, n), mmap_mode="r")
    contigs = [mm("contig%d.npy" % i) for i in range(len(names))]
    exons = [np.load(os.path.join(dirname, "exons%d.npy" % i)) for i in range(len(names))]
    ref = RefImage(names, np.load(os.path.join(dirname, "ref_contig_len.npy")), np.load(os.path.join(dirname, "ref_contig_start.npy")),
                   mm("ref_planes.npy"), mm("ref_nmask.npy"), contigs)
    return RefData(names, contigs, exons, np.load(os.path.join(dirname, "introns.npy")), ref)


def generate(cfg: SynthConfig, workers: int = 1, refdata: Optional[RefData] = None) -> Workload:
    """Reference + reads + analytic hits.  Chunks of cfg.chunk fragments carry their own RNG stream
    (seed, chunk index), so the result does not depend on `workers` (fork-based process pool)."""
    if refdata is None:
        refdata = make_reference(cfg)
    contigs, exons_pc, ref = refdata.contigs, refdata.exons_pc, refdata.ref
    space = _ExonSpace(exons_pc)
    ex_parts = []
    for ci, ex in enumerate(exons_pc):
        ln = ex[:, 1] - ex[:, 0]
        cum = np.concatenate([[0], np.cumsum(ln)])
        idx = np.arange(int(cum[-1])) - np.repeat(cum[:-1], ln) + np.repeat(ex[:, 0], ln)
        ex_parts.append(np.asarray(contigs[ci][idx]))
    exonic = np.concatenate(ex_parts)
    L = cfg.read_len
    offs, lens = segment_layout(L, cfg.segment_length)
    nseg = offs.shape[0]

    _GEN_STATE.update(cfg=cfg, contigs=contigs, space=space, exonic=exonic)
    nchunks = (cfg.n_pairs + cfg.chunk - 1) // cfg.chunk
    if workers > 1 and nchunks > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, nchunks)) as pool:
            chunk_accs = pool.map(_gen_chunk, range(nchunks), chunksize=1)
    else:
        chunk_accs = [_gen_chunk(c) for c in range(nchunks)]
    _GEN_STATE.clear()
    acc = {s: dict(reads=[], seg=[[] for _ in range(nseg)], mapped=[], unm=[], truth=[], cand=[]) for s in ("left", "right")}
    for ca in chunk_accs:
        for s in acc:
            acc[s]["reads"] += ca[s]["reads"]; acc[s]["mapped"] += ca[s]["mapped"]; acc[s]["unm"] += ca[s]["unm"]
            acc[s]["truth"] += ca[s]["truth"]; acc[s]["cand"] += ca[s]["cand"]
            for k in range(nseg):
                acc[s]["seg"][k] += ca[s]["seg"][k]
    del chunk_accs

    def finish(side: str) -> SideData:
        A = acc[side]
        reads = np.zeros((cfg.n_pairs, L), dtype=np.uint8)
        for idx, rc in A["reads"]:
            reads[idx] = rc
        unm = np.zeros(cfg.n_pairs, dtype=bool)
        for idx, u in A["unm"]:
            unm[idx] = u
        # prep_reads drops reads with >= 10% N or > 90% of one base (prep_reads.cpp:255-269): such
        # reads are never mapped, so they contribute no hits to any stream.
        qc_fail = np.zeros(cfg.n_pairs, dtype=bool)
        for lo in range(0, cfg.n_pairs, 1 << 20):
            blk = reads[lo:lo + (1 << 20)]
            frac = np.stack([(blk == c).sum(axis=1) for c in range(5)], axis=1) / float(L)
            qc_fail[lo:lo + (1 << 20)] = (frac[:, :4] > 0.9).any(axis=1) | (frac[:, 4] >= 0.1)
        seg_hits = []
        for k in range(nseg):
            h = np.concatenate(A["seg"][k]) if A["seg"][k] else np.zeros(0, dtype=SEGHIT_DTYPE)
            h = h[~qc_fail[h["read_idx"]]]
            h = h[np.argsort(h["read_idx"], kind="stable")]
            seg_hits.append(h)
        mh = np.concatenate(A["mapped"]) if A["mapped"] else np.zeros(0, dtype=SEGHIT_DTYPE)
        mh = mh[~qc_fail[mh["read_idx"]]]
        mh = mh[np.argsort(mh["read_idx"], kind="stable")]
        cand = None
        if cfg.keep_candidates:
            cand = []
            for k in range(nseg):
                ds = [c[k] for c in A["cand"]]
                if ds:
                    d = {f: np.concatenate([x[f] for x in ds]) for f in ds[0] if f != "ln"}
                    keepc = ~qc_fail[d["read_idx"]]
                    d = {f: v[keepc] for f, v in d.items()}
                    order = np.argsort(d["read_idx"], kind="stable")
                    d = {f: v[order] for f, v in d.items()}
                    d["ln"] = ds[0]["ln"]
                else:
                    d = _placement_candidates(cfg, contigs, np.zeros((0, L), np.uint8), np.zeros((0, L), np.int64), np.zeros(0, np.int64),
                                              np.zeros(0, bool), np.zeros(0, np.int64))[k]
                cand.append(d)
        truth = None
        if cfg.keep_truth:
            truth = dict(fwd=np.zeros((cfg.n_pairs, L), np.uint8), gpos=np.full((cfg.n_pairs, L), -1, np.int64),
                         ref_id=np.zeros(cfg.n_pairs, np.int64), rev=np.zeros(cfg.n_pairs, bool), qc_fail=qc_fail)
            for idx, f, gp, rid, rv in A["truth"]:
                truth["fwd"][idx] = f; truth["gpos"][idx] = gp; truth["ref_id"][idx] = rid; truth["rev"][idx] = rv
        return SideData(reads, np.arange(1, cfg.n_pairs + 1, dtype="<u4"), seg_hits, mh, unm, cand, truth)

    return Workload(cfg, ref, finish("left"), finish("right"), refdata.introns)


def subset(wl: Workload, n_pairs: int) -> Workload:
    """The first n_pairs fragments of a workload (same reference): a bounded sample for the CPU baseline."""
    def cut(sd: SideData) -> SideData:
        tr = None if sd.truth is None else {k: v[:n_pairs] for k, v in sd.truth.items()}
        cd = None
        if sd.cand is not None:
            cd = []
            for d in sd.cand:
                m = d["read_idx"] < n_pairs
                e = {f: v[m] for f, v in d.items() if f != "ln"}; e["ln"] = d["ln"]; cd.append(e)
        return SideData(sd.reads[:n_pairs], sd.ids[:n_pairs], [h[h["read_idx"] < n_pairs] for h in sd.seg_hits],
                        sd.mapped_hits[sd.mapped_hits["read_idx"] < n_pairs], sd.unmapped[:n_pairs], cd, tr)
    cfg = dataclasses.replace(wl.cfg, n_pairs=n_pairs)
    return Workload(cfg, wl.ref, cut(wl.left), cut(wl.right), wl.introns)


# ---------------------------------------------------------------------------------------------
# packed batch assembly (host-side bundle rules)


@dataclasses.dataclass
class PackedBatch:
    n_segs: int
    read_words: int
    bundles: np.ndarray
    seg_count: np.ndarray
    reads: np.ndarray
    hits: np.ndarray
    partner_hits: np.ndarray
    order_base: int = 0

    @property
    def n_bundles(self) -> int:
        return int(self.bundles.shape[0])

    def nbytes(self) -> int:
        return int(self.bundles.nbytes + self.seg_count.nbytes + self.reads.nbytes + self.hits.nbytes +
                   self.partner_hits.nbytes)


def _to_hits(a: np.ndarray) -> np.ndarray:
    h = np.zeros(a.shape[0], dtype=HIT_DTYPE)
    for f in ("ref_id", "left", "right", "read_len", "edit_dist", "flags"):
        h[f] = a[f]
    return h


def pack_side(side: SideData, partner: Optional[SideData], right_mate: bool, fusion_search: bool = False,
              order_base: int = 0) -> PackedBatch:
    """Bundles in the reference's processing order with its per-bundle call rules.

    Rules derived from look_for_hit_group / process_next_hit_group (segment_juncs.cpp:3823-4123):
      * every read with at least one segment hit is visited exactly once, in increasing id order:
        when the LAST segment stream is exhausted the final call runs with insert_id 0, whose
        observation_order is VMAXINT32 (bwt_map.h:555-560), which flushes every remaining group of
        the lower streams through the different-group branch (3975-4035);
      * t = highest segment index with hits: t == N-1 -> indels, [fusions], gaps (4092-4117);
        0 < t < N-1 -> indels, gaps, then [fusions] on the mutated bundle (4005-4033);
        t == 0 -> only [fusions] (3981, 4022-4033);
      * partner group = the mate's *.mapped.bam group with the same id, else the mate's
        last-segment group (3322-3344).
    """
    nseg = len(side.seg_hits)
    n = side.reads.shape[0]
    L = side.reads.shape[1]
    counts = np.zeros((n, nseg), dtype=np.int64)
    for k in range(nseg):
        np.add.at(counts[:, k], side.seg_hits[k]["read_idx"], 1)
    has = counts > 0
    any_hit = has.any(axis=1)
    t = np.where(any_hit, nseg - 1 - np.argmax(has[:, ::-1], axis=1), -1)
    visit = any_hit.copy()
    if not fusion_search:
        visit &= t > 0
    sel = np.nonzero(visit)[0]
    nb = sel.shape[0]
    flags = np.zeros(nb, dtype=np.uint8)
    ts = t[sel]
    flags[ts > 0] |= B_INDELS | B_GAPS
    if fusion_search:
        flags |= B_FUSIONS
        flags[(ts > 0) & (ts < nseg - 1)] |= B_FUSIONS_LAST
    if right_mate:
        flags |= B_RIGHT_MATE
    # hits in (bundle, segment, file order)
    remap = np.full(n, -1, dtype=np.int64)
    remap[sel] = np.arange(nb)
    parts, keys = [], []
    for k in range(nseg):
        h = side.seg_hits[k]
        b = remap[h["read_idx"]]
        m = b >= 0
        parts.append(_to_hits(h[m]))
        keys.append(b[m] * nseg + k)
    allh = np.concatenate(parts) if parts else np.zeros(0, dtype=HIT_DTYPE)
    key = np.concatenate(keys) if keys else np.zeros(0, dtype=np.int64)
    order = np.argsort(key, kind="stable")
    allh = allh[order]
    seg_count = counts[sel].astype("<u2")
    hit_begin = np.concatenate([[0], np.cumsum(counts[sel].sum(axis=1))])[:-1]
    # partner groups
    n_partner = np.zeros(nb, dtype=np.int64)
    pparts = np.zeros(0, dtype=HIT_DTYPE)
    pbegin = np.zeros(nb, dtype=np.int64)
    if partner is not None:
        pm = partner.mapped_hits
        pl = partner.seg_hits[len(partner.seg_hits) - 1]
        cm = np.bincount(pm["read_idx"], minlength=n)[sel]
        cl = np.bincount(pl["read_idx"], minlength=n)[sel]
        use_m = cm > 0
        n_partner = np.where(use_m, cm, cl)
        pbegin = np.concatenate([[0], np.cumsum(n_partner)])[:-1]
        # gather partner hits in bundle order
        bm = remap[pm["read_idx"]]
        bl = remap[pl["read_idx"]]
        keep_m = bm >= 0
        keep_l = (bl >= 0)
        keep_l[keep_l] &= ~use_m[bl[keep_l]]
        cand = np.concatenate([_to_hits(pm[keep_m]), _to_hits(pl[keep_l])])
        ckey = np.concatenate([bm[keep_m], bl[keep_l]])
        pparts = cand[np.argsort(ckey, kind="stable")]
    bundles = np.zeros(nb, dtype=BUNDLE_DTYPE)
    bundles["read_id"] = side.ids[sel]
    bundles["hit_begin"] = hit_begin
    bundles["partner_begin"] = pbegin
    bundles["n_partner"] = n_partner
    bundles["read_len"] = L
    bundles["flags"] = flags
    read_words = (L + 63) // 64
    return PackedBatch(nseg, read_words, bundles, np.ascontiguousarray(seg_count),
                       pack_reads(side.reads[sel], read_words), allh, pparts, order_base)


# ---------------------------------------------------------------------------------------------
# text files for the reference's own prep_reads / fix_map_ordering


def _md_and_nm(read_fwd: np.ndarray, refb: np.ndarray) -> Tuple[str, int]:
    mm = (read_fwd != refb) | (read_fwd > 3) | (refb > 3)
    out, run = [], 0
    for j in range(read_fwd.shape[0]):
        if mm[j]:
            out.append(str(run)); out.append(chr(CODE2CHAR[min(int(refb[j]), 4)])); run = 0
        else:
            run += 1
    out.append(str(run))
    return "".join(out), int(mm.sum())


def write_fasta(path: str, ref: RefImage, width: int = 60) -> None:
    with open(path, "wb") as f:
        for name, codes in zip(ref.names, ref.codes):
            f.write(b">" + name.encode() + b"\n")
            if codes is None:
                continue
            s = CODE2CHAR[codes].tobytes()
            for i in range(0, len(s), width):
                f.write(s[i:i + width]); f.write(b"\n")


def write_sam_header(path: str, ref: RefImage) -> None:
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for name, ln in zip(ref.names, ref.contig_len):
            f.write("@SQ\tSN:%s\tLN:%d\n" % (name, int(ln)))
        f.write("@PG\tID:TopHat\tVN:2.1.2\n")


def write_fastq(path: str, side: SideData) -> None:
    with open(path, "wb") as f:
        q = b"I" * side.reads.shape[1]
        for i in range(side.reads.shape[0]):
            f.write(b"@r%d\n" % (i + 1)); f.write(CODE2CHAR[side.reads[i]].tobytes()); f.write(b"\n+\n"); f.write(q); f.write(b"\n")


def write_hits_sam(path: str, wl: Workload, side: SideData, hits: np.ndarray, seg_index: Optional[int]) -> None:
    """One SAM record per hit, bowtie2-style tags.  seg_index None = whole
    read (mapped.bam stream), else qname carries `id|offset:seg:nsegs`."""
    cfg = wl.cfg
    L = cfg.read_len
    offs, lens = segment_layout(L, cfg.segment_length)
    nseg = offs.shape[0]
    with open(path, "w") as f:
        for h in hits:
            ri = int(h["read_idx"]); rid = int(h["ref_id"])
            if seg_index is None:
                sub = side.reads[ri]; qn = "%d" % (ri + 1)
            else:
                o, l = int(offs[seg_index]), int(lens[seg_index])
                sub = side.reads[ri, o:o + l]; qn = "%d|%d:%d:%d" % (ri + 1, o, seg_index, nseg)
            anti = bool(h["flags"] & HIT_ANTISENSE)
            fwd = revcomp_codes(sub) if anti else sub
            codes = wl.ref.codes[rid - 1]
            left = int(h["left"])
            nm = int(h["edit_dist"])     # decoys carry a synthetic edit distance
            if nm == 0:
                md = str(fwd.shape[0])
            else:
                md, _ = _md_and_nm(fwd, codes[left:left + fwd.shape[0]])
            f.write("%s\t%d\t%s\t%d\t255\t%dM\t*\t0\t0\t%s\t%s\tAS:i:%d\tXN:i:0\tXM:i:%d\tXO:i:0\tXG:i:0\tNM:i:%d\tMD:Z:%s\tYT:Z:UU\n" % (
                qn, 16 if anti else 0, wl.ref.names[rid - 1], left + 1, fwd.shape[0],
                CODE2CHAR[fwd].tobytes().decode(), "I" * fwd.shape[0], -6 * nm, nm, nm, md))


def write_pipeline_files(wl: Workload, outdir: str) -> Dict[str, str]:
    """FASTA, SAM header, FASTQ and the SAM text of every hit stream; large workloads write the files in forked workers."""
    os.makedirs(outdir, exist_ok=True)
    p = {}
    p["fasta"] = os.path.join(outdir, "ref.fa"); write_fasta(p["fasta"], wl.ref)
    p["header"] = os.path.join(outdir, "hdr.sam"); write_sam_header(p["header"], wl.ref)
    nseg = len(wl.left.seg_hits)
    jobs = []
    for sname, side in (("left", wl.left), ("right", wl.right)):
        p[sname + "_fq"] = os.path.join(outdir, sname + ".fq"); jobs.append(("fq", p[sname + "_fq"], sname, None))
        p[sname + "_mapped_sam"] = os.path.join(outdir, sname + "_mapped.sam"); jobs.append(("sam", p[sname + "_mapped_sam"], sname, None))
        for k in range(nseg):
            key = "%s_seg%d_sam" % (sname, k + 1)
            p[key] = os.path.join(outdir, "%s_seg%d.sam" % (sname, k + 1)); jobs.append(("sam", p[key], sname, k))

    def run(job):
        kind, path, sname, k = job
        side = wl.left if sname == "left" else wl.right
        if kind == "fq":
            write_fastq(path, side)
        else:
            write_hits_sam(path, wl, side, side.mapped_hits if k is None else side.seg_hits[k], k)

    if wl.cfg.n_pairs < 50_000 or (os.cpu_count() or 1) < 2:
        for j in jobs:
            run(j)
        return p
    pids = []
    limit = max(1, min(len(jobs), os.cpu_count() or 1))
    for j in jobs:                                           # fork per file: the workload arrays are shared copy-on-write
        while len(pids) >= limit:
            _, st = os.wait(); pids.pop()
            if st != 0:
                raise RuntimeError("a writer process failed")
        pid = os.fork()
        if pid == 0:
            try:
                run(j); os._exit(0)
            except BaseException:
                os._exit(1)
        pids.append(pid)
    while pids:
        _, st = os.wait(); pids.pop()
        if st != 0:
            raise RuntimeError("a writer process failed")
    return p


# ---------------------------------------------------------------------------------------------
# long_spanning_reads inputs: segment hits against the juncs_db contigs


def parse_juncs_db_fasta(path: str):
    """juncs_db output (juncs_db.cpp:109-149): >ref|left_start|L-R|right_end|<type>|<strand>  ->  list of
    (name, ref, left_start, L, R, type, seq)."""
    out = []
    name = None
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                name = line[1:]
            elif name is not None:
                t = name.split("|")
                if len(t) >= 6 and "-" in t[-4]:
                    lr = t[-4].split("-")
                    try:
                        out.append((name, "|".join(t[:-5]), int(t[-5]), int(lr[0]), int(lr[1]) if lr[1].isdigit() else lr[1], t[-2], line))
                    except ValueError:
                        pass
                name = None
    return out


def spliced_segment_hits(wl: Workload, side: SideData, contigs, max_mm: int = 2):
    """For every segment of every unmapped read that crosses exactly one junction for which juncs_db made an intron
    contig, the ungapped placement on that contig (what bowtie reports against the junction index), kept when it has
    <= max_mm mismatches.  Returns per segment a list of dict(read_idx, contig, pos0, anti, fwd codes, nm, md)."""
    assert side.truth is not None, "generate the workload with keep_truth=True"
    cfg = wl.cfg
    L = cfg.read_len
    offs, lens = segment_layout(L, cfg.segment_length)
    nseg = offs.shape[0]
    key = {}
    for (name, ref, left_start, jl, jr, typ, seq) in contigs:
        if typ in ("ins", "del", "fus") or not isinstance(jr, int):
            continue
        key.setdefault((ref, jl, jr), []).append((name, left_start, seq))
    names = wl.ref.names
    tr = side.truth
    out = [[] for _ in range(nseg)]
    cand = np.nonzero(side.unmapped & ~tr["qc_fail"])[0]
    gpos = tr["gpos"]; fwd = tr["fwd"]
    for ri in cand:
        g = gpos[ri]
        jumps = np.nonzero((g[1:] - g[:-1] > 1) & (g[1:] >= 0) & (g[:-1] >= 0))[0]
        if jumps.size == 0:
            continue
        rev = bool(tr["rev"][ri])
        ref = names[int(tr["ref_id"][ri]) - 1]
        for k in range(nseg):
            lo = int(L - offs[k] - lens[k]) if rev else int(offs[k]); ln = int(lens[k])
            inside = jumps[(jumps >= lo) & (jumps < lo + ln - 1)]
            if inside.size != 1:
                continue
            j = int(inside[0])
            if (g[lo:lo + ln] < 0).any():
                continue
            for (name, left_start, cseq) in key.get((ref, int(g[j]), int(g[j + 1])), []):
              x = j - lo + 1                                   # bases on the left exon
              if True:
                pos0 = (int(g[j]) - left_start + 1) - x
                if pos0 < 0 or pos0 + ln > len(cseq):
                    continue
                cc = codes_from_ascii(cseq[pos0:pos0 + ln].encode())
                seg = fwd[ri, lo:lo + ln]
                mm = (seg != cc) | (seg > 3) | (cc > 3)
                nm = int(mm.sum())
                if nm > max_mm:
                    continue
                md, run = [], 0
                for q in range(ln):
                    if mm[q]:
                        md.append(str(run)); md.append(chr(CODE2CHAR[min(int(cc[q]), 4)])); run = 0
                    else:
                        run += 1
                md.append(str(run))
                out[k].append(dict(read_idx=int(ri), contig=name, pos0=pos0, anti=rev, fwd=seg, nm=nm, md="".join(md)))
    return out


def write_spliced_sam(path: str, cfg: SynthConfig, seg_index: int, hits) -> None:
    offs, lens = segment_layout(cfg.read_len, cfg.segment_length)
    nseg = offs.shape[0]
    with open(path, "w") as f:
        for h in sorted(hits, key=lambda d: d["read_idx"]):
            o = int(offs[seg_index])
            qn = "%d|%d:%d:%d" % (h["read_idx"] + 1, o, seg_index, nseg)
            ln = h["fwd"].shape[0]
            f.write("%s\t%d\t%s\t%d\t255\t%dM\t*\t0\t0\t%s\t%s\tAS:i:%d\tXN:i:0\tXM:i:%d\tXO:i:0\tXG:i:0\tNM:i:%d\tMD:Z:%s\tYT:Z:UU\n" % (
                qn, 16 if h["anti"] else 0, h["contig"], h["pos0"] + 1, ln, CODE2CHAR[h["fwd"]].tobytes().decode(), "I" * ln,
                -6 * h["nm"], h["nm"], h["nm"], h["md"]))


def write_contig_header(path: str, contigs) -> None:
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for c in contigs:
            f.write("@SQ\tSN:%s\tLN:%d\n" % (c[0], len(c[6])))
        f.write("@PG\tID:TopHat\tVN:2.1.2\n")


# ---------------------------------------------------------------------------------------------
# vectorised junction-spanning segment placement + packed join batches (bench / large tests)

# thb_jhit_full (unpacked, CIGAR inline) and the wire form thb_jhit (16 bytes) + thb_jops (include/tophat_b200.h)
JHIT_FULL_DTYPE = np.dtype([("ref_id", "<u4"), ("left", "<i4"), ("n_ops", "u1"), ("flags", "u1"), ("mismatches", "u1"),
                            ("splice_mms", "u1"), ("ops", "<u4", (9,))])
JHIT_DTYPE = np.dtype([("ref_id", "<u4"), ("left", "<i4"), ("right", "<i4"), ("flags_nops", "u1"), ("ops_index", "u1"),
                       ("mismatches", "u1"), ("splice_mms", "u1")])
JOPS_DTYPE = np.dtype([("ops", "<u4", (12,))])
JHIT_ONE_MATCH = 8
JBUNDLE_DTYPE = np.dtype([("read_id", "<u4"), ("hit_begin", "<u4"), ("read_len", "<u2"), ("n_segs", "u1"), ("reserved", "u1"),
                          ("ops_begin", "<u4")])
JOINED_DTYPE = np.dtype([("bundle", "<u4"), ("ref_id", "<u4"), ("left", "<i4"), ("n_ops", "u1"), ("flags", "u1"), ("mismatches", "u1"),
                         ("edit_dist", "u1"), ("splice_mms", "u1"), ("reserved8", "u1", (3,)), ("ops", "<u4", (27,))])
assert JHIT_FULL_DTYPE.itemsize == 48 and JHIT_DTYPE.itemsize == 16 and JOPS_DTYPE.itemsize == 48
assert JBUNDLE_DTYPE.itemsize == 16 and JOINED_DTYPE.itemsize == 128
JHIT_ANTISENSE_SPLICE = 4
OP_MATCH, OP_INS, OP_DEL, OP_REF_SKIP = 1, 3, 5, 11


def _junc_key(ref_id, left, right):
    return (np.asarray(ref_id, dtype=np.uint64) << np.uint64(56)) | (np.asarray(left, dtype=np.uint64) << np.uint64(28)) | \
        np.asarray(right, dtype=np.uint64)


def _placement_candidates(cfg: SynthConfig, ref_codes, fwd, gpos, rid, rev, read_idx, max_mm: int = 2, min_anchor_len: int = 8,
                          want_mm: bool = False):
    """Segments that cross exactly one junction-like jump of the read's genomic origin (no inserted base inside), placed
    ungapped over the two flanks.  fwd/gpos: (m, L) genome-orientation codes / positions, rid (m,), rev (m,) bool, read_idx (m,).
    Returns per segment k a dict of arrays: read_idx, ref_id, jl, jr, x (bases left of the splice), anti, nm, smm
    [+ mm (rows, ln) bool and fwd (rows, ln) codes when want_mm]."""
    L = cfg.read_len
    offs, lens = segment_layout(L, cfg.segment_length)
    nseg = offs.shape[0]
    out = []
    for k in range(nseg):
        ln = int(lens[k])
        parts = []
        for r in (False, True):
            sel = np.nonzero(rev == r)[0]
            if sel.size == 0:
                continue
            lo = int(L - offs[k] - lens[k]) if r else int(offs[k])
            G = gpos[sel, lo:lo + ln]
            jump = (G[:, 1:] - G[:, :-1] > 1) & (G[:, 1:] >= 0) & (G[:, :-1] >= 0)
            ok = (jump.sum(axis=1) == 1) & (G >= 0).all(axis=1)
            sel, G, jump = sel[ok], G[ok], jump[ok]
            if sel.size == 0:
                continue
            j = np.argmax(jump, axis=1)
            ar = np.arange(sel.size)
            jl, jr = G[ar, j], G[ar, j + 1]
            seg = fwd[sel, lo:lo + ln]
            refb = np.empty_like(seg)
            rsel = rid[sel]
            for ci, codes in enumerate(ref_codes):
                m = rsel == ci + 1
                if m.any():
                    refb[m] = codes[G[m]]
            mm = (seg != refb) | (seg > 3) | (refb > 3)
            nm = mm.sum(axis=1)
            x = j + 1
            near = np.abs(x[:, None] - np.arange(ln)[None, :]) < min_anchor_len
            smm = (mm & near).sum(axis=1)
            keep = (nm <= max_mm) & (jl >= 26)
            d = dict(read_idx=read_idx[sel][keep], ref_id=rsel[keep].astype(np.int64), jl=jl[keep].astype(np.int64), jr=jr[keep].astype(np.int64),
                     x=x[keep].astype(np.int16), anti=np.full(int(keep.sum()), r), nm=nm[keep].astype(np.uint8), smm=smm[keep].astype(np.uint8))
            if want_mm:
                d["mm"] = mm[keep]; d["fwd"] = seg[keep]
            parts.append(d)
        if parts:
            d = {f: np.concatenate([p[f] for p in parts]) for f in parts[0]}
        else:
            d = dict(read_idx=np.zeros(0, np.int64), ref_id=np.zeros(0, np.int64), jl=np.zeros(0, np.int64), jr=np.zeros(0, np.int64),
                     x=np.zeros(0, np.int16), anti=np.zeros(0, bool), nm=np.zeros(0, np.uint8), smm=np.zeros(0, np.uint8))
            if want_mm:
                d["mm"] = np.zeros((0, ln), bool); d["fwd"] = np.zeros((0, ln), np.uint8)
        d["ln"] = ln
        out.append(d)
    return out


def _filter_by_junctions(cands, junctions: np.ndarray):
    """Keeps the placements whose (ref, left, right) is a junction record; one output row per strand variant present
    (juncs_db writes one contig per record).  Adds `asplice`; rows sorted by read index."""
    jk = {}
    for anti in (0, 1):
        sel = junctions[junctions["antisense"] == anti]
        jk[anti] = np.sort(_junc_key(sel["ref_id"], sel["left"], sel["right"]))
    out = []
    for d in cands:
        key = _junc_key(d["ref_id"], d["jl"], d["jr"])
        parts = []
        for anti in (0, 1):
            pos = np.searchsorted(jk[anti], key)
            has = pos < jk[anti].shape[0]
            has[has] &= jk[anti][pos[has]] == key[has]
            if has.any():
                p = {f: v[has] for f, v in d.items() if f != "ln"}
                p["asplice"] = np.full(int(has.sum()), anti, np.uint8)
                parts.append(p)
        if parts:
            o = {f: np.concatenate([p[f] for p in parts]) for f in parts[0]}
            order = np.argsort(o["read_idx"], kind="stable")
            o = {f: v[order] for f, v in o.items()}
        else:
            o = {f: v[:0] for f, v in d.items() if f != "ln"}
            o["asplice"] = np.zeros(0, np.uint8)
        o["ln"] = d["ln"]
        out.append(o)
    return out


def spliced_placements(wl: Workload, side: SideData, junctions: np.ndarray, max_mm: int = 2, min_anchor_len: int = 8, want_mm: bool = False):
    """Junction-index segment hits of one side for the given junction set (segment_juncs output)."""
    if side.cand is not None and not want_mm:
        return _filter_by_junctions(side.cand, junctions)
    assert side.truth is not None, "generate the workload with keep_truth=True or keep_candidates=True"
    tr = side.truth
    rows = np.nonzero(side.unmapped & ~tr["qc_fail"])[0]
    cands = _placement_candidates(wl.cfg, wl.ref.codes, tr["fwd"][rows], tr["gpos"][rows], tr["ref_id"][rows], tr["rev"][rows], rows,
                                  max_mm, min_anchor_len, want_mm)
    return _filter_by_junctions(cands, junctions)


@dataclasses.dataclass
class PackedJoinBatch:
    n_segs: int
    read_words: int
    bundles: np.ndarray
    seg_count: np.ndarray
    reads: np.ndarray
    hits: np.ndarray                   # JHIT_DTYPE (wire form)
    ops_ext: np.ndarray                # JOPS_DTYPE, one per hit without JHIT_ONE_MATCH

    @property
    def n_bundles(self) -> int:
        return int(self.bundles.shape[0])

    def nbytes(self) -> int:
        return int(self.bundles.nbytes + self.seg_count.nbytes + self.reads.nbytes + self.hits.nbytes + self.ops_ext.nbytes)


def pack_join_hits(full: np.ndarray, hit_begin: np.ndarray):
    """Vectorised thb_join_pack_hits: unpacked hits in batch order + the first hit of every read -> (heads, ops_ext, ops_begin)."""
    n = full.shape[0]
    ops = full["ops"].astype(np.int64)
    code = ops & 15
    k = np.arange(9)[None, :] < full["n_ops"][:, None].astype(np.int64)
    adv = np.where(k & ((code == OP_MATCH) | (code == OP_DEL) | (code == OP_REF_SKIP)), ops >> 4, 0).sum(axis=1)
    one = (full["n_ops"] == 1) & ((full["ops"][:, 0] & 15) == OP_MATCH)
    heads = np.zeros(n, dtype=JHIT_DTYPE)
    heads["ref_id"] = full["ref_id"]; heads["left"] = full["left"]; heads["right"] = full["left"].astype(np.int64) + adv
    heads["flags_nops"] = (full["flags"] & 7) | np.where(one, JHIT_ONE_MATCH, 0) | (full["n_ops"].astype(np.uint8) << 4)
    heads["mismatches"] = full["mismatches"]; heads["splice_mms"] = full["splice_mms"]
    multi = ~one
    before = np.concatenate([[0], np.cumsum(multi)])                 # multi-op hits before hit i
    ops_begin = before[hit_begin] if hit_begin.size else np.zeros(0, dtype=np.int64)
    owner = np.searchsorted(hit_begin, np.arange(n), side="right") - 1 if n else np.zeros(0, dtype=np.int64)
    ordinal = before[:-1] - ops_begin[owner] if n else np.zeros(0, dtype=np.int64)
    if multi.any() and int(ordinal[multi].max()) > 255:
        raise ValueError("more than 256 multi-op hits in one read")
    heads["ops_index"] = np.where(multi, ordinal, 0)
    ext = np.zeros(int(multi.sum()), dtype=JOPS_DTYPE)
    ext["ops"][:, :9] = np.where(k[multi], full["ops"][multi], 0)
    return heads, ext, ops_begin.astype("<u4")


def assemble_join_batch(side: SideData, spliced) -> PackedJoinBatch:
    """Per-read bundles of long_spanning_reads (JoinSegmentsWorker rules, long_spanning_reads.cpp:2706-2785): per segment the contiguous
    hits followed by the junction-index hits on the genome; a read is kept when every segment has at least one hit.
    spliced: per segment (read_idx array, JHIT_FULL_DTYPE array) in stream order."""
    nseg = len(side.seg_hits)
    n, L = side.reads.shape
    counts = np.zeros((n, nseg), dtype=np.int64)
    for k in range(nseg):
        counts[:, k] = np.bincount(side.seg_hits[k]["read_idx"], minlength=n) + np.bincount(spliced[k][0], minlength=n)
    visit = (counts > 0).all(axis=1)
    sel = np.nonzero(visit)[0]
    nb = sel.shape[0]
    remap = np.full(n, -1, dtype=np.int64)
    remap[sel] = np.arange(nb)
    parts, keys = [], []
    for k in range(nseg):
        h = side.seg_hits[k]
        b = remap[h["read_idx"]]; m = b >= 0
        jh = np.zeros(int(m.sum()), dtype=JHIT_FULL_DTYPE)
        hm = h[m]
        jh["ref_id"] = hm["ref_id"]; jh["left"] = hm["left"]; jh["n_ops"] = 1; jh["flags"] = hm["flags"]; jh["mismatches"] = hm["edit_dist"]
        jh["ops"][:, 0] = (hm["read_len"].astype(np.uint32) << 4) | OP_MATCH
        parts.append(jh); keys.append(b[m] * (2 * nseg) + 2 * k)
        ridx, js = spliced[k]
        b = remap[ridx]; m = b >= 0
        parts.append(js[m]); keys.append(b[m] * (2 * nseg) + 2 * k + 1)
    allh = np.concatenate(parts); key = np.concatenate(keys)
    allh = allh[np.argsort(key, kind="stable")]
    bundles = np.zeros(nb, dtype=JBUNDLE_DTYPE)
    bundles["read_id"] = side.ids[sel]
    bundles["hit_begin"] = np.concatenate([[0], np.cumsum(counts[sel].sum(axis=1))])[:-1]
    bundles["read_len"] = L; bundles["n_segs"] = nseg
    heads, ext, ops_begin = pack_join_hits(allh, bundles["hit_begin"].astype(np.int64))
    bundles["ops_begin"] = ops_begin
    rw = (L + 63) // 64
    return PackedJoinBatch(nseg, rw, bundles, np.ascontiguousarray(counts[sel].astype("<u2")), pack_reads(side.reads[sel], rw), heads, ext)


def pack_join_side(wl: Workload, side: SideData, junctions: np.ndarray) -> PackedJoinBatch:
    """assemble_join_batch with the junction-index hits placed analytically (spliced_placements) and mapped back to the genome
    as SplicedBAMHitFactory does (bwt_map.cpp:1469-1770)."""
    nseg = len(side.seg_hits)
    spl = spliced_placements(wl, side, junctions)
    spliced = []
    for k in range(nseg):
        d = spl[k]
        js = np.zeros(int(d["read_idx"].shape[0]), dtype=JHIT_FULL_DTYPE)
        x = d["x"]; ln = d["ln"]
        js["ref_id"] = d["ref_id"]; js["left"] = d["jl"] - x + 1; js["n_ops"] = 3
        js["flags"] = np.where(d["anti"], HIT_ANTISENSE, 0) | (HIT_END if k == nseg - 1 else 0) | np.where(d["asplice"] > 0, JHIT_ANTISENSE_SPLICE, 0)
        js["mismatches"] = d["nm"]; js["splice_mms"] = d["smm"]
        js["ops"][:, 0] = (x.astype(np.uint32) << 4) | OP_MATCH
        js["ops"][:, 1] = ((d["jr"] - d["jl"] - 1).astype(np.uint32) << 4) | OP_REF_SKIP
        js["ops"][:, 2] = ((ln - x).astype(np.uint32) << 4) | OP_MATCH
        spliced.append((d["read_idx"], js))
    return assemble_join_batch(side, spliced)


def spliced_hits_for_sam(wl: Workload, side: SideData, junctions: np.ndarray, contigs):
    """Placements (vectorised) -> the per-segment record lists write_spliced_sam consumes; needs keep_truth."""
    key = {}
    for (name, ref, left_start, jl, jr, typ, seq) in contigs:
        if typ in ("ins", "del", "fus") or not isinstance(jr, int):
            continue
        key[(ref, jl, jr, 1 if name.endswith("|rev") else 0)] = (name, left_start, seq)
    names = wl.ref.names
    out = []
    for d in spliced_placements(wl, side, junctions, want_mm=True):
        rows = []
        ln = d["ln"]
        for i in range(d["read_idx"].shape[0]):
            ent = key.get((names[int(d["ref_id"][i]) - 1], int(d["jl"][i]), int(d["jr"][i]), int(d["asplice"][i])))
            if ent is None:
                continue
            name, left_start, cseq = ent
            x = int(d["x"][i]); pos0 = (int(d["jl"][i]) - left_start + 1) - x
            if pos0 < 0 or pos0 + ln > len(cseq):
                continue
            if d["nm"][i] == 0:
                md = str(ln)
            else:
                mm = d["mm"][i]; parts, run = [], 0
                for q in range(ln):
                    if mm[q]:
                        parts.append(str(run)); parts.append(cseq[pos0 + q] if cseq[pos0 + q] in "ACGT" else "N"); run = 0
                    else:
                        run += 1
                parts.append(str(run)); md = "".join(parts)
            rows.append(dict(read_idx=int(d["read_idx"][i]), contig=name, pos0=pos0, anti=bool(d["anti"][i]), fwd=d["fwd"][i], nm=int(d["nm"][i]), md=md))
        out.append(rows)
    return out
