"""Once-off check at scale: our segment_juncs executable (-p <all cores>, GPU) against the reference's own
segment_juncs -p1 (oracle/_ref) on the same BAM / FASTA files; the three text outputs must be byte-identical.
usage: python scripts/p1_compare.py --pairs 10000000 [--workload chr20] --out profiles/r2_p1_compare.json"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=10_000_000)
    ap.add_argument("--workload", default="chr20")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from tophat_b200 import synth, build
    from oracle import pyoracle
    log = {"pairs": args.pairs, "workload": args.workload, "host_cpus": os.cpu_count()}
    t = time.time()
    wl = bench.make_workload(args.pairs, 0, os.cpu_count() or 1, keep_truth=True, kind=args.workload)
    log["generate_s"] = time.time() - t
    d = tempfile.mkdtemp(prefix="thb_p1_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        t = time.time()
        files = synth.write_pipeline_files(wl, d)
        nseg = len(wl.left.seg_hits)
        bams = pyoracle.make_bams(files, d, nseg)
        log["write_inputs_s"] = time.time() - t
        log["input_bytes"] = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".bam"))
        opts = pyoracle.tophat_common_opts(50, 20)
        ours = os.path.join(build.BIN_DIR, "segment_juncs")
        threads = os.cpu_count() or 1
        for k in range(2):
            t = time.perf_counter()
            o = pyoracle.run_segment_juncs(ours, files, bams, d, nseg, opts=opts, threads=threads, tag=".b200")
            log["ours_p%d_wall_s_run%d" % (threads, k)] = time.perf_counter() - t
        t = time.perf_counter()
        r = pyoracle.run_segment_juncs(os.path.join(pyoracle.REF_DIR, "segment_juncs"), files, bams, d, nseg, opts=opts, threads=1, tag=".p1")
        log["reference_p1_wall_s"] = time.perf_counter() - t
        same = {}
        for k in ("juncs", "insertions", "deletions"):
            a = open(o[k], "rb").read(); b = open(r[k], "rb").read()
            same[k] = {"identical": a == b, "bytes": len(b), "lines": b.count(b"\n")}
        log["outputs"] = same
        log["all_identical"] = all(v["identical"] for v in same.values())
        log["speedup_vs_p1"] = log["reference_p1_wall_s"] / log["ours_p%d_wall_s_run1" % threads]
    finally:
        shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(log, indent=1))
    if args.out:
        json.dump(log, open(args.out, "w"), indent=1)
    if not log.get("all_identical"):
        sys.exit(1)


if __name__ == "__main__":
    main()
