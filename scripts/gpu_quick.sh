#!/bin/bash
# quick GPU check: join parity tests + a 4M-pair bench line (per-kernel times)
# usage: scripts/gpu_quick.sh <tag> [pairs] [pytest -k expression]
TAG=${1:-q}; PAIRS=${2:-4000000}; KEXPR=${3:-"join or long_spanning"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --pairs $PAIRS --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)  frac(all) %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["all_kernels"]["frac"]))
print({k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms_per_step"].items()})
print(d["results"])
PY
