#!/bin/bash
# where the start-up of the executables goes: a tiny input, TOPHAT_GPU_STATS=1, a few runs each
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/startup_probe.log 2>&1
import os, sys, time, subprocess, tempfile, json
sys.path.insert(0, os.getcwd())
from tophat_b200 import synth, build
from oracle import pyoracle
import bench
wl = bench.make_workload(2000, 0, 8, keep_truth=True, kind="chr20")
d = tempfile.mkdtemp(dir="/dev/shm")
files = synth.write_pipeline_files(wl, d); bams = pyoracle.make_bams(files, d, 4)
opts = pyoracle.tophat_common_opts(50, 20)
env = dict(os.environ, TOPHAT_GPU_STATS="1", THB_TRACE="1")
for exe in (os.path.join(build.BIN_DIR, "segment_juncs"), os.path.join(pyoracle.REF_DIR, "segment_juncs")):
    for k in range(3):
        t = time.perf_counter()
        outs = pyoracle.run_segment_juncs(exe, files, bams, d, 4, opts=opts, threads=16, tag=".x", env=env) if "b200" in exe or True else None
        print(os.path.basename(os.path.dirname(exe)), "run", k, "wall %.3f s" % (time.perf_counter() - t), flush=True)
    log = [f for f in os.listdir(d) if f.startswith("segment_juncs") and f.endswith(".log")]
    for f in log:
        print(f, open(os.path.join(d, f)).read()[-1500:])
t = time.perf_counter(); subprocess.run([sys.executable, "-c", "import ctypes; l = ctypes.CDLL('%s'); p = ctypes.c_void_p(); import time; t=time.perf_counter(); print('thb_create rc', l.thb_create(0, ctypes.byref(p)), 'in %.3f s' % (time.perf_counter()-t))" % os.path.join(os.getcwd(), "tophat_b200", "libtophat_b200.so")]); print("python ctypes process %.3f s" % (time.perf_counter() - t))
PY
tail -40 gpurun_out/startup_probe.log | cut -c1-600
