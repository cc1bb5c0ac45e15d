#!/usr/bin/env python
"""bench.py -- measurement of the splice-junction hot path on B200 (contract: task statement + DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic 2x101 bp read pairs:
  stage 1 segment_juncs       thb_segjuncs_begin -> submit(left mates) -> submit(right mates) -> [N>1: NCCL all-gather
                              of the discovered sets] -> thb_segjuncs_finish
  stage 2 long_spanning_reads thb_join_begin(junction / indel sets) -> join(left mates) -> join(right mates)  Workload at N=1: BASELINE.json configs[1] (10 M pairs, chr20-sized reference).  For
N>1 every rank runs the same number of pairs (its own shard of reads, same reference): weak scaling.

  value  : reads (mates) per second, whole job, inputs resident in HBM (CUDA events on the library's stream)
  e2e    : same metric through the C ABI with HOST (pinned) buffers: H2D copies + D2H results inside the timing
  roofline : scan kernel, algorithmic bytes (SURVEY.md §8d formula on the actual task counts) / CUDA-event time
  cpu_baseline : the reference's own segment_juncs binary (oracle/_ref) on a bounded sample, same box

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
WORKLOAD = "chr20"            # set from --workload: chr20 (configs[1], default) | hg38 (configs[2]) | indel (configs[3])
KERNELS = ("bundle", "hit", "rescue", "rescued_windows", "window_scan", "indel")
METRIC = "spliced reads aligned/s (segment_juncs + long_spanning_reads), 2x101bp"
UNIT = "reads/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------
# workload


def make_workload(pairs: int, rank: int, workers: int, keep_candidates: bool = False, keep_truth: bool = False):
    from tophat_b200 import synth
    chunk = 500_000 if pairs >= 500_000 else max(1000, pairs)
    nchunks = (pairs + chunk - 1) // chunk
    cfg = synth.SynthConfig(contig_lens=HG38 if WORKLOAD == "hg38" else (CHR20,), n_pairs=pairs, seed=20240611, chunk=chunk,
                            indel_prob=0.5 if WORKLOAD == "indel" else 0.0,
                            chunk_seed_base=rank * nchunks, keep_candidates=keep_candidates, keep_truth=keep_truth)
    t = time.time()
    wl = synth.generate(cfg, workers=workers)
    log("[bench] rank %d: generated %d pairs in %.1f s (%d workers)" % (rank, pairs, time.time() - t, workers))
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
        return min(self._run(self.tfiles, self.tbams, self.tjin, self.tdir, 1, bins, ".b200" if bins else "") for _ in range(2))

    def step(self, bins=None) -> float:
        return self._run(self.files, self.bams, self.jin, self.dir, self.threads, bins, ".b200" if bins else "")

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def usable_threads(sample_pairs: int) -> int:
    # -p N is honoured only when every .index side file has >= N entries (utils.cpp:75-79); an entry is written
    # every >= 1000 records, the sparsest stream (*.mapped.bam of one side) holds ~0.25 records per pair
    n = os.cpu_count() or 1
    return max(1, min(n, sample_pairs // 8000))


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    wl = make_workload(args.ref_pairs, 0, max(1, (os.cpu_count() or 1)), keep_truth=True)
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
    work = max(per - st, 1e-6)
    val = reads / work
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.ref_pairs, 1, note="bounded sample of the GPU arm's workload; value excludes the "
                                      "%.2f s fixed start-up (FASTA load) measured on a 1-pair input" % st),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": "%d pairs (%d reads), oracle/_ref segment_juncs + long_spanning_reads (left, right) -p%d, wall %.2f s/step incl. %.2f s start-up"
                                       % (arm.sample_pairs, reads, threads, per, st)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(pairs: int, world: int, note: str = ""):
    what = {"chr20": "BASELINE configs[1]: synthetic 2x101 bp pairs, chr20-sized (64,444,167 bp) reference, ",
            "hg38": "BASELINE configs[2]: synthetic 2x101 bp pairs, hg38-sized reference (24 contigs, 3.09 Gbp, 0.5% N), ",
            "indel": "BASELINE configs[3]: indel-heavy synthetic 2x101 bp pairs (1-3 bp indel in half of the mates), chr20-sized reference, "}[WORKLOAD]
    c = {"workload": what + "4 segments/mate (25/25/25/26), segment hits placed analytically (SURVEY.md 8d)",
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
# GPU arm


def run_gpu_arm(args, rank: int, local_rank: int, world: int):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from tophat_b200 import capi, synth
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    workers = max(1, (os.cpu_count() or 1) // world)
    wl = make_workload(args.pairs, rank, workers, keep_candidates=True)
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
        res = ctx.segjuncs_finish(copy)       # copy=False: views of the library's page-locked result arrays (no Python-side memcpy)
        return res, ctx.timing()

    # stage 2 inputs: the junction-index segment hits depend on the junction set stage 1 finds (tophat.py:3686-3741 runs
    # juncs_db + bowtie in between); they are derived once from a set-up pass and reused by every timed step
    t0 = time.time()
    res0, _ = segjuncs_pass(True)
    jsets = capi.join_sets_from_results(res0)
    # page-locked copies of the junction / insertion sets (stage-2 inputs that thb_join_begin uploads every step)
    _jpin = [torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).pin_memory() for a in jsets]
    jsets = tuple(t.numpy().view(a.dtype) for t, a in zip(_jpin, jsets))
    jbatches = [synth.pack_join_side(wl, wl.left, res0.junctions), synth.pack_join_side(wl, wl.right, res0.junctions)]
    log("[bench] rank %d: join batches %d + %d reads, %d segment hits (%.1f s)" % (
        rank, jbatches[0].n_bundles, jbatches[1].n_bundles, sum(int(b.hits.shape[0]) for b in jbatches), time.time() - t0))
    j_fields = ("bundles", "seg_count", "reads", "hits", "ops_ext")
    j_keep, j_dev, j_host = stage(jbatches, j_fields, capi.join_batch_c)
    h2d_bytes = sum(b.nbytes() for b in batches) + sum(b.nbytes() for b in jbatches) + int(jsets[0].nbytes + jsets[1].nbytes)

    def step(device_resident: bool, copy: bool = False):
        res, tm = segjuncs_pass(device_resident, copy)
        ctx.join_begin(P, jsets[0], jsets[1])
        n_joined, d2h = 0, 0
        for i in range(len(jbatches)):
            if device_resident:
                n_joined += ctx.join_submit_device(j_dev[i])
            else:
                out = C.c_void_p(); n = C.c_uint64()
                ctx._check(ctx.lib.thb_join_submit(ctx.h, C.byref(j_host[i]), C.byref(out), C.byref(n)), "thb_join_submit")
                n_joined += int(n.value); d2h += int(n.value) * 128
        return res, tm, ctx.join_timing(), n_joined, d2h

    def timed(device_resident: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(device_resident)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = dict(scan_ms=0.0, alg=0, launches=0, join_ms=0.0, join_alg=0, join_launches=0, enum_ms=0.0, merge_ms=0.0,
                   merge_simple_ms=0.0, merge_abutting_ms=0.0, merge_general_ms=0.0,
                   kms={k: 0.0 for k in KERNELS})
        e0.record(stream)
        for k_ in range(steps):
            res, tm, jt, n_joined, d2h = step(device_resident, copy=(k_ == steps - 1))     # the last step's sets are kept for the checks below
            acc["scan_ms"] += tm.scan_kernel_ms; acc["alg"] += tm.algorithmic_bytes; acc["launches"] += tm.total_launches + jt.launches
            acc["join_ms"] += jt.kernel_ms; acc["join_alg"] += jt.algorithmic_bytes; acc["join_launches"] += jt.launches
            acc["enum_ms"] += jt.enum_ms; acc["merge_ms"] += jt.merge_ms
            for k in ("merge_simple_ms", "merge_abutting_ms", "merge_general_ms"):
                acc[k] += getattr(jt, k)
            for k in KERNELS:
                acc["kms"][k] += getattr(tm, k + "_ms")
        e1.record(stream)
        torch.cuda.synchronize()
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
        #   bundle      : 16 (bundle header) per bundle + 16 per partner hit
        #   hit         : 40 (read) per bundle + 16 per segment hit
        #   window_scan : 32 (descriptor) + 64 (two reference sectors) per window + 16 per emitted junction record
        #   rescue      : 32 + 128 per mate-anchor task;  indel: 32 + 64 per task + 16 per record
        #   chain_enum  : 16 (header) per read + 16 per segment hit (the 16-byte wire record)
        #   chain_merge : 40 (read) per read + 32 per multi-op hit + 128 per closure + 96 per merged chain + 128 per output record
        n_b = sum(b.n_bundles for b in batches); n_h = sum(int(b.hits.shape[0]) + int(b.partner_hits.shape[0]) for b in batches)
        n_hh = sum(int(b.hits.shape[0]) for b in batches)
        steps = args.steps
        kbytes = {"bundle": 16 * n_b + 16 * (n_h - n_hh), "hit": 40 * n_b + 16 * n_hh,
                  "window_scan": 96 * int(tm.n_windows) + 16 * int(tm.n_juncs_emitted), "rescue": 160 * int(tm.n_rescue_tasks),
                  "rescued_windows": 0, "indel": 96 * int(tm.n_indel_tasks) + 16 * (len(res.deletions) + len(res.insertions)),
                  "chain_enum": 16 * sum(b.n_bundles for b in jbatches) + 16 * sum(int(b.hits.shape[0]) for b in jbatches)}
        # the three merge kernels share the rest of B_join by their number of chains; the closure terms (128 per closure) belong
        # to the general kernel alone
        jt_ = A["jt"]
        merge_total = A["join_alg"] / steps - kbytes["chain_enum"]
        n_ch = max(1, int(jt_.n_chains)); n_s, n_a = int(jt_.n_simple_chains), int(jt_.n_abutting_chains); n_g = max(0, n_ch - n_s - n_a)
        closure_bytes = 128.0 * int(jt_.n_closures)
        rest = max(0.0, merge_total - closure_bytes)
        kbytes["chain_merge_simple"] = rest * n_s / n_ch
        kbytes["chain_merge_abut"] = rest * n_a / n_ch
        kbytes["chain_merge"] = rest * n_g / n_ch + closure_bytes
        kms = {k: A["kms"][k] / steps for k in KERNELS}; kms["chain_enum"] = A["enum_ms"] / steps
        kms["chain_merge_simple"] = A["merge_simple_ms"] / steps; kms["chain_merge_abut"] = A["merge_abutting_ms"] / steps
        kms["chain_merge"] = A["merge_general_ms"] / steps
        klaunch = {k: A["scan_launches"] for k in KERNELS}
        allk = KERNELS + ("chain_enum", "chain_merge_simple", "chain_merge_abut", "chain_merge")
        for k in allk[len(KERNELS):]:
            klaunch[k] = max(1, A["join_launches"] // (4 * steps))
        dom = max(allk, key=lambda k: kms[k])
        dom_ms = kms[dom] / klaunch[dom]; dom_bytes = kbytes[dom] / klaunch[dom]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        tot_ms = sum(kms.values()); tot_bytes = A["alg"] / steps + A["join_alg"] / steps
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = next((v for k, v in tj.items() if k.startswith(dom + "_kernel")), None)
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": workload_config(args.pairs, world),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": E["ms"] / args.steps,
                        "api": "thb_segjuncs_begin/submit(host, pinned)/finish + thb_join_begin/submit(host, pinned)"},
                "gpu_launches": int(A["launches"]),
                "roofline": {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peaks, "unit": "GB/s",
                             "frac": achieved / peaks, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms_per_launch": dom_ms,
                             "launches_per_step": klaunch[dom],
                             "per_kernel_ms_per_step": kms,
                             "per_kernel_gbs": {k: (kbytes[k] / (kms[k] * 1e-3) / 1e9 if kms[k] > 0 else 0.0) for k in allk},
                             "all_kernels": {"ms_per_step": tot_ms, "algorithmic_bytes_per_step": tot_bytes,
                                             "achieved": tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0,
                                             "frac": (tot_bytes / (tot_ms * 1e-3) / 1e9 / peaks) if tot_ms > 0 else 0.0}},
                "clocks": clocks,
                "results": {"junctions": int(len(res.junctions)), "deletions": int(len(res.deletions)),
                            "insertions": int(len(res.insertions)), "windows": int(tm.n_windows),
                            "indel_tasks": int(tm.n_indel_tasks), "rescue_tasks": int(tm.n_rescue_tasks),
                            "join_reads": int(sum(b.n_bundles for b in jbatches)), "joined_alignments": int(A["n_joined"]),
                            "join_chains": int(A["jt"].n_chains), "join_closures": int(A["jt"].n_closures),
                            "join_simple_chains": int(A["jt"].n_simple_chains), "join_abutting_chains": int(A["jt"].n_abutting_chains)}}
    ctx.close()

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                threads = usable_threads(args.ref_pairs)
                swl = make_workload(args.ref_pairs, 0, os.cpu_count() or 1, keep_truth=True)
                arm = ReferenceArm(swl, args.ref_pairs, threads)
                cli = None
                try:
                    st = arm.startup(); arm.step(); wall = arm.step()
                    # the drop-in executables themselves (tophat_b200/bin, C++ hosts over the C ABI) on the very same files: BAM
                    # decode, bundle building, GPU work, text / BAM output -- what a tophat.py run would see per stage
                    try:
                        from tophat_b200 import build as _b
                        ours = (os.path.join(_b.BIN_DIR, "segment_juncs"), os.path.join(_b.BIN_DIR, "long_spanning_reads"))
                        if all(os.access(x, os.X_OK) for x in ours):
                            st_o = arm.startup(ours); arm.step(ours); wall_o = arm.step(ours)
                            # compared with the reference at -p1: its own -pN output can lack junctions of reads at the thread
                            # partition boundaries (observed: 1 of 28,969 at -p12 on this sample), ours equals the -p1 answer
                            p1 = arm.py.run_segment_juncs(arm.sj, arm.files, arm.bams, arm.dir, arm.nseg, opts=arm.opts, threads=1, tag=".p1")
                            same = all(open(os.path.join(arm.dir, "segment.b200." + k)).read() == open(p1[k]).read()
                                       for k in ("juncs", "insertions", "deletions"))
                            cli = {"value": 2 * arm.sample_pairs / max(wall_o - st_o, 1e-6), "unit": UNIT, "wall_s": wall_o, "startup_s": st_o,
                                   "reference_wall_s": wall, "reference_startup_s": st, "segment_files_identical_to_reference_p1": same,
                                   "note": "our segment_juncs + long_spanning_reads executables on the reference arm's sample files; "
                                           "start-up (CUDA context, FASTA load, image upload) measured on a 1-pair input and excluded like the reference's"}
                    except Exception as e:
                        cli = {"value": None, "note": "failed: %r" % (e,)}
                finally:
                    arm.close()
                if cli is not None:
                    line["drop_in_cli"] = cli
                reads = 2 * arm.sample_pairs
                line["cpu_baseline"] = {"value": reads / max(wall - st, 1e-6), "unit": UNIT, "cores": threads, "kind": "reference",
                                        "sample": "%d pairs (%d reads) from the same generator and reference; oracle/_ref segment_juncs + "
                                                  "long_spanning_reads (left, right) -p%d; wall %.2f s of which %.2f s fixed start-up "
                                                  "(FASTA loads, excluded)" % (arm.sample_pairs, reads, threads, wall, st)}
            except Exception as e:  # the baseline leg must never hide the GPU result
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("THB_BENCH_PAIRS", 10_000_000)), help="read pairs per GPU")
    ap.add_argument("--ref-pairs", type=int, default=int(os.environ.get("THB_BENCH_REF_PAIRS", 100_000)),
                    help="pairs in the bounded sample the reference CPU binary is timed on")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chr20", choices=["chr20", "hg38", "indel"],
                    help="chr20 = BASELINE configs[1] (the default, the configuration the metric is quoted on at one GPU); hg38 = configs[2]; indel = configs[3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
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
