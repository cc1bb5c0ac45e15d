#!/usr/bin/env python
"""bench.py -- measurement of the splice-junction hot path on B200 (contract: task statement + DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic 2x101 bp read pairs:
  stage 1 segment_juncs       thb_segjuncs_begin -> submit(left mates) -> submit(right mates) -> [N>1: NCCL all-gather
                              of the discovered sets] -> thb_segjuncs_finish
  stage 2 long_spanning_reads thb_join_begin(junction / indel sets) -> join(left mates) -> join(right mates)
Default workload at every N: BASELINE.json configs[2] -- the configuration the metric is quoted on (hg38-sized reference at 1/2/4/8
GPUs) -- sharded by read: 6.25 M pairs per GPU, so that N = 8 is the 50 M-pair configuration (weak scaling; every rank draws its own
read chunks over the same genome).  --workload chr20 (configs[1], 10 M pairs) and --workload indel (configs[3]) are one flag away;
their lines are kept under profiles/.

  value  : reads (mates) per second, whole job, inputs resident in HBM (CUDA events on the library's stream)
  e2e    : same metric through the C ABI with HOST (pinned) buffers: H2D copies + D2H results inside the timing
  roofline : dominant kernel, algorithmic bytes (SURVEY.md §8d formula on the actual task counts) / CUDA-event time; every kernel and
             the non-kernel parts of the step are listed beside it
  parity_checked : every rank, after the timed region: sharded + all-gathered sets == single context == CPU oracle on a 200 k-pair sample
  cpu_baseline / drop_in_cli (N = 1) : the reference's own binaries (oracle/_ref) and OUR executables on the same BAM / FASTA files

`--impl reference` times the reference CPU binary only (rank 0 alone under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

CHR20 = 64_444_167
# BASELINE configs[2]: 24 contigs of hg38's primary-assembly sizes (chr1..22, X, Y), 3.09 Gbp
HG38 = (248_956_422, 242_193_529, 198_295_559, 190_214_555, 181_538_259, 170_805_979, 159_345_973, 145_138_636, 138_394_717,
        133_797_422, 135_086_622, 133_275_309, 114_364_328, 107_043_718, 101_991_189, 90_338_345, 83_257_441, 80_373_285,
        58_617_616, 64_444_167, 46_709_983, 50_818_468, 156_040_895, 57_227_415)
WORKLOAD = "hg38"             # set from --workload: hg38 (configs[2], default) | chr20 (configs[1]) | indel (configs[3])
KERNELS = ("bundle", "hit", "rescue", "rescued_windows", "window_scan", "indel")
METRIC = "spliced reads aligned/s (segment_juncs + long_spanning_reads), 2x101bp"
UNIT = "reads/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------
# workload


def contig_lens(kind: str):
    return HG38 if kind == "hg38" else (CHR20,)


def default_pairs(kind: str) -> int:
    # hg38: BASELINE configs[2] (50 M pairs over 8 GPUs) sharded by read: 6.25 M pairs per GPU at every N (weak scaling; N = 8 is configs[2])
    return 6_250_000 if kind == "hg38" else 10_000_000


_REFDATA = {}


def get_refdata(kind: str, rank: int = 0, world: int = 1, barrier=None):
    """The synthetic genome of a workload kind.  With several ranks on one box rank 0 generates it once and the others map
    it from /dev/shm (an hg38-sized genome is 3.1 GB of codes + 1.2 GB of bit planes: not something to build 8 times)."""
    from tophat_b200 import synth
    if kind in _REFDATA:
        return _REFDATA[kind]
    cfg = synth.SynthConfig(contig_lens=contig_lens(kind), n_pairs=1, seed=20240611)
    t = time.time()
    if world == 1:
        rd = synth.make_reference(cfg)
    else:
        d = "/dev/shm/thb_bench_ref_%s_%s" % (kind, os.environ.get("MASTER_PORT", "0"))
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)
            rd0 = synth.make_reference(cfg)
            synth.save_refdata(rd0, d + ".tmp"); os.rename(d + ".tmp", d)
            del rd0
        barrier()
        rd = synth.load_refdata(d)
        barrier()
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)      # the mappings stay valid until the last rank exits
    log("[bench] rank %d: %s reference ready in %.1f s" % (rank, kind, time.time() - t))
    _REFDATA[kind] = rd
    return rd


def make_workload(pairs: int, rank: int, workers: int, keep_candidates: bool = False, keep_truth: bool = False, kind: str = None,
                  refdata=None, seed_base: int = None):
    from tophat_b200 import synth
    kind = kind or WORKLOAD
    chunk = 500_000 if pairs >= 500_000 else max(1000, pairs)
    nchunks = (pairs + chunk - 1) // chunk
    cfg = synth.SynthConfig(contig_lens=contig_lens(kind), n_pairs=pairs, seed=20240611, chunk=chunk,
                            indel_prob=0.5 if kind == "indel" else 0.0,
                            chunk_seed_base=rank * nchunks if seed_base is None else seed_base, keep_candidates=keep_candidates, keep_truth=keep_truth)
    t = time.time()
    wl = synth.generate(cfg, workers=workers, refdata=refdata if refdata is not None else get_refdata(kind))
    log("[bench] rank %d: generated %d pairs (%s) in %.1f s (%d workers)" % (rank, pairs, kind, time.time() - t, workers))
    return wl


def pack(wl):
    from tophat_b200 import synth
    t = time.time()
    bl = synth.pack_side(wl.left, wl.right, False)
    br = synth.pack_side(wl.right, wl.left, True, order_base=bl.n_bundles)
    log("[bench] packed %d + %d bundles in %.1f s" % (bl.n_bundles, br.n_bundles, time.time() - t))
    return [bl, br]


# ----------------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref binaries; the only place besides tests/ that executes oracle/)


class ReferenceArm:
    """Prepares BAM/FASTA inputs for a sample workload (generated with keep_truth) and times the reference's own
    segment_juncs followed by long_spanning_reads for both mates (juncs_db + the junction-index segment mapping that
    tophat.py runs in between are set-up, not timed, exactly as they are outside the GPU arm's timed region)."""

    def __init__(self, wl, sample_pairs: int, threads: int):
        from tophat_b200 import synth
        from oracle import pyoracle
        self.py = pyoracle
        if not pyoracle.have_reference():
            raise RuntimeError("oracle/_ref is not built (run __graft_entry__.build() where /root/reference exists)")
        self.sample_pairs = min(sample_pairs, wl.cfg.n_pairs)
        self.threads = threads
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        self.dir = tempfile.mkdtemp(prefix="thb_ref_", dir=base)
        t = time.time()
        sub = synth.subset(wl, self.sample_pairs) if self.sample_pairs < wl.cfg.n_pairs else wl
        self.nseg = len(sub.left.seg_hits)
        self.files = synth.write_pipeline_files(sub, self.dir)
        self.bams = pyoracle.make_bams(self.files, self.dir, self.nseg)
        self.opts = pyoracle.tophat_common_opts(50, 20)
        self.sj = os.path.join(pyoracle.REF_DIR, "segment_juncs"); self.lsr = os.path.join(pyoracle.REF_DIR, "long_spanning_reads")
        self.outs = pyoracle.run_segment_juncs(self.sj, self.files, self.bams, self.dir, self.nseg, opts=self.opts, threads=threads)
        self.jin = pyoracle.make_join_inputs(sub, self.files, self.outs, self.dir, self.nseg, fast=True)
        # a 1-pair input measures the fixed start-up (FASTA load of the same reference) of every binary run
        tiny = synth.subset(wl, 1)
        self.tdir = os.path.join(self.dir, "tiny"); os.makedirs(self.tdir)
        tf = synth.write_pipeline_files(tiny, self.tdir)
        tf["fasta"] = self.files["fasta"]; tf["header"] = self.files["header"]
        self.tfiles, self.tbams = tf, pyoracle.make_bams(tf, self.tdir, self.nseg)
        self.touts = pyoracle.run_segment_juncs(self.sj, self.tfiles, self.tbams, self.tdir, self.nseg, opts=self.opts)
        self.tjin = pyoracle.make_join_inputs(tiny, self.tfiles, self.touts, self.tdir, self.nseg, fast=True)
        log("[bench] reference arm inputs for %d pairs ready in %.1f s" % (self.sample_pairs, time.time() - t))

    def _run(self, files, bams, jin, outdir, threads, bins=None, tag=""):
        sj, lsr = bins if bins else (self.sj, self.lsr)
        t = time.perf_counter()
        outs = self.py.run_segment_juncs(sj, files, bams, outdir, self.nseg, opts=self.opts, threads=threads, tag=tag)
        for side in ("left", "right"):
            self.py.run_long_spanning_reads(lsr, files, bams, jin, outs, outdir, self.nseg, side=side, opts=self.opts, threads=threads, tag=tag)
        return time.perf_counter() - t

    def startup(self, bins=None) -> float:
        """Fixed cost of one step (three process starts: FASTA loads; ours also CUDA context + image upload), measured on a 1-pair
        input with the SAME -p and the same (warm) file cache as the step; best of two."""
        return min(self._run(self.tfiles, self.tbams, self.tjin, self.tdir, self.threads, bins, ".b200" if bins else "") for _ in range(2))

    def step(self, bins=None) -> float:
        return self._run(self.files, self.bams, self.jin, self.dir, self.threads, bins, ".b200" if bins else "")

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def usable_threads(sample_pairs: int) -> int:
    # -p N is honoured only when every .index side file has >= N entries (utils.cpp:75-79); an entry is written
    # every >= 1000 records, the sparsest stream (*.mapped.bam of one side) holds ~0.25 records per pair
    n = os.cpu_count() or 1
    return max(1, min(n, sample_pairs // 8000))


def reference_kind() -> str:
    # The reference binaries load the whole FASTA on every start (three starts per step): ~37 s each for an hg38-sized genome, which no
    # bounded sample can amortise.  Their per-read work does not depend on the genome size, so the CPU arm of the hg38 workload is timed
    # on the chr20-sized genome of the same generator (smaller working set: if anything this favours the reference).
    return "chr20" if WORKLOAD == "hg38" else WORKLOAD


def reference_note(st: float) -> str:
    n = "bounded sample from the GPU arm's generator; value excludes the %.2f s fixed start-up (three FASTA loads) measured on a 1-pair input at the same -p" % st
    if WORKLOAD == "hg38":
        n += "; timed on the chr20-sized genome: an hg38-sized FASTA costs the reference ~37 s per process start, per-read work is genome-size independent"
    return n


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    kind = reference_kind()
    wl = make_workload(args.ref_pairs, 0, max(1, (os.cpu_count() or 1)), keep_truth=True, kind=kind)
    threads = usable_threads(args.ref_pairs)
    arm = ReferenceArm(wl, args.ref_pairs, threads)
    try:
        st = arm.startup()
        for _ in range(args.warmup):
            arm.step()
        times = [arm.step() for _ in range(args.steps)]
    finally:
        arm.close()
    reads = 2 * arm.sample_pairs
    per = sum(times) / len(times)
    val = reads / (per - st) if per > st else reads / per
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.ref_pairs, 1, note=reference_note(st)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "host_cpu_count": os.cpu_count(), "kind": "reference",
                             "wall_s_per_step": per, "startup_s": st, "value_incl_startup": reads / per,
                             "sample": "%d pairs (%d reads), %s-sized genome, oracle/_ref segment_juncs + long_spanning_reads (left, right) -p%d of %d host "
                                       "CPUs, wall %.2f s/step incl. %.2f s start-up" % (arm.sample_pairs, reads, kind, threads, os.cpu_count() or 0, per, st)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


HOST_HANDOFF = False


def workload_config(pairs: int, world: int, note: str = "", join_inputs: str = ""):
    what = {"chr20": "BASELINE configs[1]: synthetic 2x101 bp pairs, chr20-sized (64,444,167 bp) reference, ",
            "hg38": "BASELINE configs[2] sharded by read (50 M pairs over 8 GPUs = 6.25 M per GPU): synthetic 2x101 bp pairs, hg38-sized reference "
                    "(24 contigs, 3.09 Gbp, 0.5% N; 1.2 GB image, not L2-resident), ",
            "indel": "BASELINE configs[3]: indel-heavy synthetic 2x101 bp pairs (1-3 bp indel in half of the mates), chr20-sized reference, "}[WORKLOAD]
    c = {"workload": what + "4 segments/mate (25/25/25/26), genome segment hits placed analytically (SURVEY.md 8d)",
         "join_inputs": join_inputs or "junction-index segment hits placed analytically",
         "pairs_per_gpu": pairs, "reads_per_gpu": 2 * pairs, "stages": "segment_juncs (junction / indel discovery) + long_spanning_reads (segment-chain join)",
         "l2": "inputs (>1 GB per step at the default size) exceed the 126 MB L2; no explicit flush",
         "parallelism": "read-shard x%d%s" % (world, " + NCCL all-gather of the junction/indel sets" if world > 1 else "")}
    if note:
        c["note"] = note
    return c


# ----------------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index; self.proc = None; self.path = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(prefix="thb_clk_", suffix=".csv"); os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                      "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            os.unlink(self.path)
        if not rows and shutil.which("nvidia-smi"):
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True).stdout
            rows = [l.strip().split(", ") for l in out.splitlines() if l.strip()]
        sm, mx, reasons = [], 0, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# parity inside the measured run


def _digest(res) -> str:
    import hashlib
    h = hashlib.sha256()
    for a in (res.junctions, res.deletions, res.insertions, res.fusions):
        h.update(np.ascontiguousarray(a).tobytes()); h.update(b"|")
    return h.hexdigest()


def parity_check(ctx, wl, P, rank: int, world: int, dist, total_pairs: int):
    """Run by EVERY rank after the timed region, on the configuration being measured: a sample of `total_pairs` pairs (every rank
    contributes the head of its own shard) goes through (1) the sharded path -- own batches + thb_segjuncs_allgather, (2) a
    single context over all ranks' batches in the reference's processing order (left mates, then right mates: the -p1 answer,
    segment_juncs.cpp:4752-4922) and (3) the CPU oracle (rank 0).  All three must agree record for record on every rank."""
    from tophat_b200 import synth
    from oracle import pyoracle
    t0 = time.time()
    sub = max(1000, min(wl.cfg.n_pairs, total_pairs // world))
    swl = synth.subset(wl, sub)
    mine = (synth.pack_side(swl.left, swl.right, False), synth.pack_side(swl.right, swl.left, True))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]
    order = [g[0] for g in gathered] + [g[1] for g in gathered]
    base = 0
    for b in order:
        b.order_base = base; base += b.n_bundles
    ctx.segjuncs_begin(P)
    ctx.segjuncs_submit(gathered[rank][0]); ctx.segjuncs_submit(gathered[rank][1])
    if world > 1:
        ctx.segjuncs_allgather()
    U = ctx.segjuncs_finish()
    ctx.segjuncs_begin(P)
    for b in order:
        ctx.segjuncs_submit(b)
    V = ctx.segjuncs_finish()
    du, dv = _digest(U), _digest(V)
    dorc = [None]
    if rank == 0:
        want, _ = pyoracle.segjuncs(P, wl.ref, order)
        dorc[0] = _digest(want)
    if world > 1:
        dist.broadcast_object_list(dorc, src=0)
    ok = {"sharded_union_equals_single_context": du == dv, "single_context_equals_cpu_oracle": dv == dorc[0]}
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        ok = {k: all(f[k] for f in flags) for k in ok}
    out = {"pairs": sub * world, "bundles": int(base), "junctions": int(len(V.junctions)), "deletions": int(len(V.deletions)),
           "insertions": int(len(V.insertions)), "ranks_checked": world, "seconds": round(time.time() - t0, 1)}
    out.update(ok)
    if not all(ok.values()):
        raise AssertionError("parity check failed on rank %d: %r" % (rank, out))
    return out


# ----------------------------------------------------------------------------------------------------
# GPU arm


def flank_join_inputs(ctx, wl, res0, P, peaks_for_flank=None, steps: int = 3):
    """Stage-2 inputs through the junction-flank matcher (thb_flank_*: the device replacement of juncs_db + bowtie-build + bowtie of the
    segments, tophat.py:2546-2600, 3686-3741) instead of placing the spliced segments analytically: index over the sets of the
    set-up pass, every segment of every unmapped read searched, placements mapped back to the genome on the device, then the
    per-read bundles of long_spanning_reads.  Returns the two join batches and the matcher's own measurements (CUDA events)."""
    import torch
    from tophat_b200 import capi, synth
    offs, lens = synth.segment_layout(wl.cfg.read_len, wl.cfg.segment_length)
    bounds = [int(o) for o in offs] + [int(offs[-1] + lens[-1])]
    nseg = len(lens); rw = (wl.cfg.read_len + 63) // 64
    FP = capi.FlankParams(int(P.segment_mismatches), int(P.max_seg_multihits), int(lens.min()), int(lens.max()), 3, 1)
    t0 = time.time()
    ctx.flank_begin(FP, res0.junctions, res0.deletions, res0.insertions, res0.fusions)
    begin_wall = time.time() - t0
    ft = ctx.flank_timing()
    info = {"replaces": "juncs_db <3> <max_seg_len> + bowtie-build + bowtie -v %d -k %d -m %d of every segment of the unmapped reads "
                        "(tophat.py:2546-2600, 3686-3741); outside the timed step, like the reference arm's own set-up" % (FP.max_mismatches, FP.max_multihits, FP.max_multihits),
            "index": {"contigs": int(ft.n_contigs), "entries": int(ft.n_index_entries), "device_ms": float(ft.index_ms), "wall_ms": begin_wall * 1e3},
            "sides": []}
    batches = []
    for side in (wl.left, wl.right):
        idx = np.nonzero(side.unmapped)[0]
        pin = torch.from_numpy(synth.pack_reads(side.reads[idx], rw)).pin_memory()
        dev = pin.cuda(); torch.cuda.synchronize()
        ms, vs = [], []
        for k in range(steps + 1):                                  # one warm-up
            ctx.flank_submit(len(idx), rw, bounds, device_ptr=dev.data_ptr(), copy=False)
            t = ctx.flank_timing()
            if k:
                ms.append((t.match_ms, t.post_ms)); vs.append(int(t.n_verified))
        w0 = time.time()
        hits = ctx.flank_submit(pin.numpy(), rw, bounds)            # the call a host makes: pinned host reads in, placements out
        host_ms = (time.time() - w0) * 1e3
        t = ctx.flank_timing()
        jh = ctx.flank_spliced_hits(int(P.min_anchor_len))
        keep = jh["n_ops"] > 0
        spliced = []
        for k in range(nseg):
            m = (hits["seg"] == k) & keep
            spliced.append((idx[hits["read"][m]], jh[m]))
        # the analytically placed spliced segments (the previous stage-2 input) must all be among the matcher's
        ana = synth.spliced_placements(wl, side, res0.junctions)
        def keyof(read, seg, x, left):
            return (read.astype(np.uint64) << np.uint64(40)) | (np.uint64(seg) << np.uint64(37)) | (x.astype(np.uint64) << np.uint64(32)) | (left.astype(np.int64).astype(np.uint64) & np.uint64(0xffffffff))
        got = np.concatenate([keyof(r, k, (j["ops"][:, 0] >> 4), j["left"]) for k, (r, j) in enumerate(spliced)]) if spliced else np.zeros(0, np.uint64)
        want = np.concatenate([keyof(d["read_idx"], k, d["x"], d["jl"] - d["x"] + 1) for k, d in enumerate(ana)])
        missing = int((~np.isin(want, got)).sum())
        subset = missing == 0
        match_ms = float(np.mean([a for a, _ in ms])); post_ms = float(np.mean([b for _, b in ms]))
        info["sides"].append({"reads": int(len(idx)), "segments": int(len(idx)) * nseg, "placements": int(len(hits)), "spliced_hits": int(keep.sum()),
                              "analytic_placements": int(want.shape[0]), "analytic_subset_of_matcher": subset, "analytic_missing": missing, "verified_candidates": int(np.mean(vs)),
                              "match_ms": match_ms, "post_ms": post_ms, "host_call_ms": host_ms, "h2d_bytes": int(pin.numel() * 8), "d2h_bytes": int(hits.nbytes),
                              "algorithmic_bytes": int(t.algorithmic_bytes)})
        if not subset:
            log("[bench] WARNING: %d analytically placed spliced segments are not among the matcher's placements" % missing)
        batches.append(synth.assemble_join_batch(side, spliced))
        del dev, pin
    tot_ms = sum(x["match_ms"] + x["post_ms"] for x in info["sides"])
    info["segments_per_s"] = sum(x["segments"] for x in info["sides"]) / (tot_ms * 1e-3) if tot_ms > 0 else 0.0
    info["reads_per_s"] = sum(x["reads"] for x in info["sides"]) / (tot_ms * 1e-3) if tot_ms > 0 else 0.0
    info["achieved_gbs"] = sum(x["algorithmic_bytes"] for x in info["sides"]) / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    return batches, info


def bind_to_gpu_numa(local_rank: int):
    """Pins this rank (and the threads / worker processes it starts) to the CPUs of the NUMA node its GPU hangs off, so the page-locked
    batch buffers it allocates are node-local to the GPU's PCIe root (first-touch policy).  Returns what it did for the JSON line."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return {"gpu": bdf, "numa_node": node, "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"gpu": bdf, "numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"gpu": bdf, "numa_node": node, "bound": True, "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError) as e:
        return {"bound": False, "why": repr(e)[:80]}


def run_gpu_arm(args, rank: int, local_rank: int, world: int):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from tophat_b200 import capi, synth
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local_rank) if not args.no_numa_bind else {"bound": False, "why": "--no-numa-bind"}
    barrier = None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=30))
        barrier = dist.barrier
    workers = max(1, min(len(os.sched_getaffinity(0)), (os.cpu_count() or 1) // world * 2))
    refdata = get_refdata(WORKLOAD, rank, world, barrier)
    wl = make_workload(args.pairs, rank, workers, keep_candidates=True, refdata=refdata)
    batches = pack(wl)
    n_reads = 2 * args.pairs
    P = capi.default_params(inner_dist_mean=50, inner_dist_std_dev=20)
    ctx = capi.Context(local_rank)
    ctx.ref_upload(wl.ref)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    def stage(batch_list, fields, to_struct):
        """device-resident copies (value) and pinned host copies (e2e) of packed batches + their C structs"""
        keep, dev_s, host_s = [], [], []
        for b in batch_list:
            t = {k: torch.from_numpy(np.ascontiguousarray(getattr(b, k)).view(np.uint8).reshape(-1)) for k in fields}
            d = {k: v.cuda() for k, v in t.items()}
            pt = {k: v.pin_memory() for k, v in t.items()}
            keep.append((d, pt))
            for src, dst in ((d, dev_s), (pt, host_s)):
                bc = to_struct(b)
                for k in fields:
                    setattr(bc, k, src[k].data_ptr())
                dst.append(bc)
        torch.cuda.synchronize()
        return keep, dev_s, host_s

    sj_fields = ("bundles", "seg_count", "reads", "hits", "partner_hits")
    sj_keep, sj_dev, sj_host = stage(batches, sj_fields, capi.batch_c)

    def segjuncs_pass(device_resident: bool, copy: bool = True):
        ctx.segjuncs_begin(P)
        for i in range(len(batches)):
            if device_resident:
                ctx.segjuncs_submit_device(sj_dev[i])
            else:
                ctx._check(ctx.lib.thb_segjuncs_submit(ctx.h, C.byref(sj_host[i])), "thb_segjuncs_submit")
        if world > 1:
            ctx.segjuncs_allgather()
        # copy=False (every step but the last): the C call alone -- the result sets land in the library's page-locked arrays either way,
        # building numpy views of them is the harness's business, not the path's
        if device_resident and not args.host_handoff and not copy:
            res = ctx.segjuncs_finish_resident()      # sets stay on the device for thb_join_begin_resident; their download runs behind stage 2
        else:
            res = ctx.segjuncs_finish(True) if copy else ctx.segjuncs_finish_raw()
        return res, ctx.timing()

    # stage 2 inputs: the junction-index segment hits depend on the junction set stage 1 finds (tophat.py:3686-3741 runs
    # juncs_db + bowtie in between); they are derived once from a set-up pass and reused by every timed step
    t0 = time.time()
    res0, _ = segjuncs_pass(True)
    jsets = capi.join_sets_from_results(res0)
    # page-locked copies of the junction / insertion sets (stage-2 inputs that thb_join_begin uploads every step)
    _jpin = [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).pin_memory() for a in jsets]
    jsets = tuple(t.numpy().view(a.dtype) for t, a in zip(_jpin, jsets))
    # The indel-heavy configuration keeps the analytically placed junction-index hits by default: its segment_juncs pass reports 2 M
    # insertion / deletion candidates per 10 M pairs, the matcher (like bowtie would) places 120 M segments per side on their contigs
    # and the join enumerates 1.3 G chains -- 2.2 s per step, measured once (profiles/r2y_bench_indel_matcher_inputs.json); it is a
    # different workload from round 1's, so it is opt-in there (--matcher-join-inputs).
    if args.analytic_join_inputs or (WORKLOAD == "indel" and not args.matcher_join_inputs):
        jbatches = [synth.pack_join_side(wl, wl.left, res0.junctions), synth.pack_join_side(wl, wl.right, res0.junctions)]
        flank = None
    else:
        jbatches, flank = flank_join_inputs(ctx, wl, res0, P, peaks_for_flank=None)
    log("[bench] rank %d: join batches %d + %d reads, %d segment hits (%.1f s)" % (
        rank, jbatches[0].n_bundles, jbatches[1].n_bundles, sum(int(b.hits.shape[0]) for b in jbatches), time.time() - t0))
    j_fields = ("bundles", "seg_count", "reads", "hits", "ops_ext")
    j_keep, j_dev, j_host = stage(jbatches, j_fields, capi.join_batch_c)
    h2d_bytes = sum(b.nbytes() for b in batches) + sum(b.nbytes() for b in jbatches) + int(jsets[0].nbytes + jsets[1].nbytes)

    wall = {}

    def clock(name, f, *a):
        """per-call host wall clock (every C-ABI call returns with its stream idle), reported with --breakdown"""
        t = time.perf_counter(); r = f(*a); wall[name] = wall.get(name, 0.0) + (time.perf_counter() - t); return r

    def segjuncs_pass_timed(device_resident: bool, copy: bool):
        clock("segjuncs_begin", ctx.segjuncs_begin, P)
        for i in range(len(batches)):
            if device_resident:
                clock("segjuncs_submit", ctx.segjuncs_submit_device, sj_dev[i])
            else:
                clock("segjuncs_submit", lambda: ctx._check(ctx.lib.thb_segjuncs_submit(ctx.h, C.byref(sj_host[i])), "thb_segjuncs_submit"))
        if world > 1:
            clock("allgather", ctx.segjuncs_allgather)
        if device_resident and not args.host_handoff and not copy:
            res = clock("segjuncs_finish", ctx.segjuncs_finish_resident)
        else:
            res = clock("segjuncs_finish", ctx.segjuncs_finish, True) if copy else clock("segjuncs_finish", ctx.segjuncs_finish_raw)
        return res, ctx.timing()

    def step(device_resident: bool, copy: bool = False):
        resident = device_resident and not args.host_handoff and not copy
        if args.breakdown:
            res, tm = segjuncs_pass_timed(device_resident, copy)
            if resident:
                clock("join_begin", ctx.join_begin_resident, P)
            else:
                clock("join_begin", ctx.join_begin, P, jsets[0], jsets[1])
        else:
            res, tm = segjuncs_pass(device_resident, copy)
            if resident:
                ctx.join_begin_resident(P)
            else:
                ctx.join_begin(P, jsets[0], jsets[1])
        n_joined, d2h = 0, 0
        for i in range(len(jbatches)):
            if device_resident:
                n_joined += clock("join_submit", ctx.join_submit_device, j_dev[i]) if args.breakdown else ctx.join_submit_device(j_dev[i])
            else:
                out = C.c_void_p(); n = C.c_uint64()
                ctx._check(ctx.lib.thb_join_submit(ctx.h, C.byref(j_host[i]), C.byref(out), C.byref(n)), "thb_join_submit")
                n_joined += int(n.value); d2h += int(n.value) * 128
        if resident:
            clock("segjuncs_fetch", ctx.segjuncs_fetch) if args.breakdown else ctx.segjuncs_fetch()      # the sets are on the host when the step ends
        return res, tm, ctx.join_timing(), n_joined, d2h

    def timed(device_resident: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(device_resident)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = dict(scan_ms=0.0, alg=0, launches=0, join_ms=0.0, join_alg=0, join_launches=0, enum_ms=0.0, merge_ms=0.0,
                   merge_simple_ms=0.0, merge_abutting_ms=0.0, merge_general_ms=0.0, finish_ms=0.0, begin_ms=0.0,
                   kms={k: 0.0 for k in KERNELS})
        e0.record(stream)
        for k_ in range(steps):
            res, tm, jt, n_joined, d2h = step(device_resident, copy=False)
            acc["scan_ms"] += tm.scan_kernel_ms; acc["alg"] += tm.algorithmic_bytes; acc["launches"] += tm.total_launches + jt.launches
            acc["join_ms"] += jt.kernel_ms; acc["join_alg"] += jt.algorithmic_bytes; acc["join_launches"] += jt.launches
            acc["enum_ms"] += jt.enum_ms; acc["merge_ms"] += jt.merge_ms
            acc["finish_ms"] += tm.finish_ms; acc["begin_ms"] += jt.begin_ms
            for k in ("merge_simple_ms", "merge_abutting_ms", "merge_general_ms"):
                acc[k] += getattr(jt, k)
            for k in KERNELS:
                acc["kms"][k] += getattr(tm, k + "_ms")
        e1.record(stream)
        torch.cuda.synchronize()
        # the last step's result sets (still in the library's page-locked arrays) as numpy copies for the checks below
        res = capi.SegJuncsResults.from_c(res, copy=True)
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        acc.update(ms=ms, res=res, tm=tm, jt=jt, n_joined=n_joined, d2h=d2h, scan_launches=tm.kernel_launches)
        return acc

    clk = ClockSampler(local_rank)
    if rank == 0:
        clk.start()
    A = timed(True, args.steps, args.warmup)
    wall_dev = {k: v / (args.steps + args.warmup) * 1e3 for k, v in wall.items()}; wall.clear()
    E = timed(False, args.steps, args.warmup)
    clocks = clk.stop() if rank == 0 else None
    res, res_h, tm = A["res"], E["res"], A["tm"]
    d2h_bytes = int(res_h.junctions.nbytes + res_h.deletions.nbytes + res_h.insertions.nbytes + res_h.fusions.nbytes) + E["d2h"]

    # the two arms must agree with each other (same sets / same number of alignments from device-resident and host submission)
    assert res.junctions.shape == res_h.junctions.shape and (res.junctions == res_h.junctions).all()
    assert A["n_joined"] == E["n_joined"]
    # size-independent properties of the full-size result (std::set semantics of the reference's containers): strictly increasing
    # in Junction order, spans inside [min_segment_intron - 16, max_segment_intron + segment_length + 16]
    for arr in (res.junctions, res.deletions):
        if len(arr) > 1:
            k1 = (arr["ref_id"].astype(np.uint64) << np.uint64(32)) | arr["left"].astype(np.uint64)
            k2 = (arr["right"].astype(np.uint64) << np.uint64(1)) | arr["antisense"].astype(np.uint64)
            assert ((k1[1:] > k1[:-1]) | ((k1[1:] == k1[:-1]) & (k2[1:] > k2[:-1]))).all(), "result set not strictly increasing"
    if len(res.junctions):
        span = res.junctions["right"].astype(np.int64) - res.junctions["left"].astype(np.int64)
        assert span.min() >= 50 - 16 and span.max() <= 500000 + 25 + 16
    # parity on this very configuration, every rank (N > 1: the all-gathered union against a single context and the CPU oracle)
    parity = parity_check(ctx, wl, P, rank, world, dist, args.parity_pairs) if args.parity_pairs > 0 else None

    total_reads = n_reads * world
    per_step = A["ms"] / args.steps
    value = total_reads / (per_step * 1e-3)
    e2e_val = total_reads / (E["ms"] / args.steps * 1e-3)

    line = None
    if rank == 0:
        peaks, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            mp_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peaks, peak_src = float(mp_["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        # Per-kernel split of SURVEY.md 8(d)'s B_segjuncs / B_join, on this run's actual task counts:
        #   scan_tile   : 16 (bundle header) + 40 (read) per bundle + 16 per segment hit + 16 per partner hit
        #   window_scan : 32 (descriptor) + 64 (two reference sectors) per window + 16 per emitted junction record
        #   rescue      : 32 + 128 per mate-anchor task;  indel: 32 + 64 per task + 16 per record
        #   join_tile   : 16 (header) per read + 16 per segment hit, + the share of [40 (read) per read + 32 per multi-op hit + 96 per merged
        #                 chain + 128 per output record] of the chains it merges itself (those without a closure)
        #   chain_merge : the same share for the chains with a closure + 128 per closure
        legacy_scan, legacy_join = "THB_SCAN_LEGACY" in os.environ, "THB_JOIN_TILE" not in os.environ
        n_b = sum(b.n_bundles for b in batches); n_h = sum(int(b.hits.shape[0]) + int(b.partner_hits.shape[0]) for b in batches)
        n_hh = sum(int(b.hits.shape[0]) for b in batches)
        steps = args.steps
        kbytes = {"window_scan": 96 * int(tm.n_windows) + 16 * int(tm.n_juncs_emitted), "rescue": 160 * int(tm.n_rescue_tasks),
                  "rescued_windows": 0, "indel": 96 * int(tm.n_indel_tasks) + 16 * (len(res.deletions) + len(res.insertions))}
        b_bundle, b_hit = 16 * n_b + 16 * (n_h - n_hh), 40 * n_b + 16 * n_hh
        b_enum = 16 * sum(b.n_bundles for b in jbatches) + 16 * sum(int(b.hits.shape[0]) for b in jbatches)
        jt_ = A["jt"]
        merge_total = A["join_alg"] / steps - b_enum
        n_ch = max(1, int(jt_.n_chains)); n_s, n_a = int(jt_.n_simple_chains), int(jt_.n_abutting_chains); n_g = max(0, n_ch - n_s - n_a)
        closure_bytes = 128.0 * int(jt_.n_closures)
        rest = max(0.0, merge_total - closure_bytes)
        kms = {k: A["kms"][k] / steps for k in KERNELS}
        if legacy_scan:
            kbytes["bundle"], kbytes["hit"] = b_bundle, b_hit
        else:
            kbytes["scan_tile"] = b_bundle + b_hit
            kms["scan_tile"] = kms.pop("bundle"); kms.pop("hit")
        if legacy_join:
            kbytes["chain_enum"] = b_enum
            kbytes["chain_merge_simple"] = rest * n_s / n_ch; kbytes["chain_merge_abut"] = rest * n_a / n_ch
            kms["chain_enum"] = A["enum_ms"] / steps
            kms["chain_merge_simple"] = A["merge_simple_ms"] / steps; kms["chain_merge_abut"] = A["merge_abutting_ms"] / steps
        else:
            kbytes["join_tile"] = b_enum + rest * (n_s + n_a) / n_ch
            kms["join_tile"] = A["enum_ms"] / steps
        kbytes["chain_merge"] = rest * n_g / n_ch + closure_bytes
        kms["chain_merge"] = A["merge_general_ms"] / steps
        allk = tuple(kms.keys())
        n_join_kernels = 4 if legacy_join else 2
        klaunch = {k: (A["scan_launches"] if k in KERNELS or k == "scan_tile" else max(1, A["join_launches"] // (n_join_kernels * steps))) for k in allk}
        dom = max(allk, key=lambda k: kms[k])
        dom_ms = kms[dom] / klaunch[dom]; dom_bytes = kbytes[dom] / klaunch[dom]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        tot_ms = sum(kms.values()); tot_bytes = A["alg"] / steps + A["join_alg"] / steps
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                ent = tj.get(WORKLOAD, tj)
                traffic = next((v for k, v in ent.items() if k.startswith(dom + "_kernel")), None)
                if traffic is not None and ent.get("pairs_per_launch"):
                    traffic = traffic * args.pairs / ent["pairs_per_launch"]       # captured at another launch size: linear in the reads
            except Exception:
                traffic = None
        if args.host_handoff:
            non_kernel = {"segjuncs_finish (set compaction, CUB sorts, decode, D2H of the sets)": A["finish_ms"] / steps,
                          "join_begin (set upload, validation, bucket index build)": A["begin_ms"] / steps}
        else:
            non_kernel = {"segjuncs_finish_resident (set compaction, CUB sorts, decode; the D2H of the sets runs behind stage 2)": A["finish_ms"] / steps,
                          "join_begin_resident (device-to-device hand-off of the sets, validation, bucket index build)": A["begin_ms"] / steps}
        non_kernel["other (queue counters read back, launch gaps, wait for the sets' download, Python between calls)"] = max(0.0, per_step - tot_ms - sum(non_kernel.values()))
        if flank is not None:
            # what one pass costs when the junction index is rebuilt and searched every step as well (sum of the measured parts; the
            # reference spends juncs_db + bowtie-build + 2 x nseg bowtie runs here, outside both arms' timed region)
            extra = flank["index"]["device_ms"] + sum(x["match_ms"] + x["post_ms"] for x in flank["sides"])
            flank["step_plus_junction_index"] = {"ms": per_step + extra, "reads_per_s": n_reads / ((per_step + extra) * 1e-3),
                                                 "note": "per GPU: ms_per_step + index build + search of both sides, device times"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": per_step, "higher_is_better": True, "scaling": "strong" if args.total_pairs > 0 else "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": dict(workload_config(args.pairs, world, join_inputs=(
                    "junction-index segment hits from the junction-flank matcher (thb_flank_*) over the sets of the set-up pass" if flank is not None
                    else "junction-index segment hits placed analytically")),
                    handoff="host" if HOST_HANDOFF else "device-resident sets (value arm); host arrays (e2e arm)"),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": E["ms"] / args.steps,
                        "api": "thb_segjuncs_begin/submit(host, pinned)/finish + thb_join_begin/submit(host, pinned); the packed batches are built "
                               "once outside the timed region (in the executables that is BAM decoding: see drop_in_cli)"},
                "gpu_launches": int(A["launches"]),
                "roofline": {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peaks, "unit": "GB/s",
                             "frac": achieved / peaks, "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture of this workload, DRAM read + write bytes per launch, scaled linearly to this launch size)" if traffic else None,
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms_per_launch": dom_ms,
                             "launches_per_step": klaunch[dom],
                             "per_kernel_ms_per_step": kms,
                             "per_kernel_gbs": {k: (kbytes[k] / (kms[k] * 1e-3) / 1e9 if kms[k] > 0 else 0.0) for k in allk},
                             "per_kernel_frac": {k: (kbytes[k] / (kms[k] * 1e-3) / 1e9 / peaks if kms[k] > 0 else 0.0) for k in allk},
                             "non_kernel_ms_per_step": non_kernel,
                             "all_kernels": {"ms_per_step": tot_ms, "algorithmic_bytes_per_step": tot_bytes,
                                             "achieved": tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0,
                                             "frac": (tot_bytes / (tot_ms * 1e-3) / 1e9 / peaks) if tot_ms > 0 else 0.0}},
                "clocks": clocks, "numa": numa, "flank_match": flank, "parity_checked": parity, "host_wall_ms_per_call_kind": wall_dev if args.breakdown else None,
                "host": {"cpu_count": os.cpu_count()},
                "results": {"junctions": int(len(res.junctions)), "deletions": int(len(res.deletions)),
                            "insertions": int(len(res.insertions)), "windows": int(tm.n_windows),
                            "indel_tasks": int(tm.n_indel_tasks), "rescue_tasks": int(tm.n_rescue_tasks),
                            "join_reads": int(sum(b.n_bundles for b in jbatches)), "joined_alignments": int(A["n_joined"]),
                            "join_chains": int(A["jt"].n_chains), "join_closures": int(A["jt"].n_closures),
                            "join_simple_chains": int(A["jt"].n_simple_chains), "join_abutting_chains": int(A["jt"].n_abutting_chains)}}
    ctx.close()
    del sj_keep, j_keep

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                os.sched_setaffinity(0, all_cpus)   # the CPU legs get every core of the box again
                cpu, cli = cpu_and_cli(args)
                line["cpu_baseline"] = cpu
                if cli is not None:
                    line["drop_in_cli"] = cli
            except Exception as e:  # the baseline leg must never hide the GPU result
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_and_cli(args):
    """cpu_baseline (the reference's own binaries) and drop_in_cli (OUR executables on the same BAM / FASTA files: BAM decode, bundle
    building, GPU work, text / BAM output -- what a tophat.py run sees per stage), both as plain wall clock of one step = segment_juncs +
    long_spanning_reads (left) + long_spanning_reads (right) at -p<threads>, file cache warm, nothing subtracted or clamped; the fixed
    start-up of a step (measured on a 1-pair input at the same -p) is reported beside it."""
    kind = reference_kind()
    threads = usable_threads(args.cli_pairs)
    swl = make_workload(args.cli_pairs, 0, os.cpu_count() or 1, keep_truth=True, kind=kind)
    arm = ReferenceArm(swl, args.cli_pairs, threads)
    cli = None
    try:
        reads = 2 * arm.sample_pairs
        st = arm.startup(); arm.step(); wall = arm.step()
        cpu = {"value": reads / (wall - st) if wall > st else reads / wall, "unit": UNIT, "cores": threads, "host_cpu_count": os.cpu_count(), "kind": "reference",
               "wall_s": wall, "startup_s": st, "value_incl_startup": reads / wall,
               "sample": "%d pairs (%d reads) from the same generator, %s-sized genome; oracle/_ref segment_juncs + long_spanning_reads (left, right) "
                         "-p%d of %d host CPUs; wall %.2f s of which %.2f s fixed start-up (three FASTA loads; `value` excludes it, "
                         "`value_incl_startup` does not)" % (arm.sample_pairs, reads, kind, threads, os.cpu_count() or 0, wall, st)}
        try:
            from tophat_b200 import build as _b
            ours = (os.path.join(_b.BIN_DIR, "segment_juncs"), os.path.join(_b.BIN_DIR, "long_spanning_reads"))
            if all(os.access(x, os.X_OK) for x in ours):
                st_o = arm.startup(ours); arm.step(ours); wall_o = arm.step(ours)
                # our output equals the reference's -p1 answer (tests/test_cli_*); the reference's own -pN output can lack junctions of
                # reads at its thread-partition boundaries, so here: every line the reference found must be in ours, and the count of
                # extra lines is reported (exact equality vs -p1 is checked when the sample is small enough for a -p1 run)
                cmp_ = {}
                for k in ("juncs", "insertions", "deletions"):
                    a = open(os.path.join(arm.dir, "segment.b200." + k)).read().splitlines(); b = open(arm.outs[k]).read().splitlines()
                    cmp_[k] = {"ours": len(a), "reference_pN": len(b), "reference_lines_missing_from_ours": len(set(b) - set(a))}
                same_p1 = None
                if arm.sample_pairs <= 150_000:
                    p1 = arm.py.run_segment_juncs(arm.sj, arm.files, arm.bams, arm.dir, arm.nseg, opts=arm.opts, threads=1, tag=".p1")
                    same_p1 = all(open(os.path.join(arm.dir, "segment.b200." + k)).read() == open(p1[k]).read() for k in ("juncs", "insertions", "deletions"))
                cli = {"value": reads / wall_o, "unit": UNIT, "wall_s": wall_o, "startup_s": st_o,
                       "value_excl_startup": reads / (wall_o - st_o) if wall_o > st_o else None,
                       "reference_wall_s": wall, "reference_startup_s": st, "reference_value": reads / wall,
                       "speedup_wall": wall / wall_o, "threads": threads, "pairs": arm.sample_pairs,
                       "segment_files_vs_reference_pN": cmp_, "segment_files_identical_to_reference_p1": same_p1,
                       "note": "our segment_juncs + long_spanning_reads executables on the reference arm's sample files, whole wall clock of a step "
                               "(process starts, CUDA context, FASTA / image load included)"}
        except Exception as e:
            cli = {"value": None, "note": "failed: %r" % (e,)}
    finally:
        arm.close()
    return cpu, cli


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("THB_BENCH_PAIRS", 0)), help="read pairs per GPU (default: by workload)")
    ap.add_argument("--ref-pairs", type=int, default=int(os.environ.get("THB_BENCH_REF_PAIRS", 250_000)),
                    help="pairs in the bounded sample the reference CPU binaries are timed on (--impl reference); >= 8000 x host CPUs lets -p use them all")
    ap.add_argument("--cli-pairs", type=int, default=int(os.environ.get("THB_BENCH_CLI_PAIRS", 500_000)),
                    help="pairs in the sample on which cpu_baseline and drop_in_cli are timed inside the GPU arm's run (N = 1)")
    ap.add_argument("--parity-pairs", type=int, default=200_000, help="pairs in the in-run parity check (0 = off)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("THB_BENCH_WORKLOAD", "hg38"), choices=["chr20", "hg38", "indel"],
                    help="hg38 = BASELINE configs[2] (the configuration the metric is quoted on: hg38-sized reference at 1/2/4/8 GPUs, sharded by read); "
                         "chr20 = configs[1]; indel = configs[3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--analytic-join-inputs", action="store_true",
                    help="place the spliced segments of stage 2 analytically (round 1) instead of through the junction-flank matcher")
    ap.add_argument("--matcher-join-inputs", action="store_true", help="--workload indel: stage-2 inputs through the junction-flank matcher as well")
    ap.add_argument("--total-pairs", type=int, default=0,
                    help="strong scaling: this many read pairs divided over the GPUs (default: --pairs per GPU, weak scaling)")
    ap.add_argument("--host-handoff", action="store_true",
                    help="device-resident arm: hand the sets from stage 1 to stage 2 through the host (thb_segjuncs_finish + thb_join_begin) "
                         "instead of thb_segjuncs_finish_resident + thb_join_begin_resident")
    ap.add_argument("--no-numa-bind", action="store_true", help="leave the rank on all CPUs instead of its GPU's NUMA node")
    ap.add_argument("--breakdown", action="store_true", help="also report the host wall clock of every C-ABI call kind (device-resident pass)")
    args = ap.parse_args()
    global WORKLOAD, HOST_HANDOFF
    WORKLOAD = args.workload
    HOST_HANDOFF = bool(args.host_handoff)
    if args.total_pairs > 0:
        # strong scaling: a fixed job divided over the GPUs (BASELINE configs[2] as a whole is --total-pairs 50000000)
        args.pairs = max(1, args.total_pairs // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
    if args.pairs <= 0:
        args.pairs = default_pairs(WORKLOAD)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
