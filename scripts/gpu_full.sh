#!/bin/bash
# full check: GPU test suite + the default bench line (hg38 workload, cpu_baseline + drop_in_cli) [+ reference arm]
TAG=${1:-full}; EXTRA=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi_$TAG.txt; free -g >> gpurun_out/smi_$TAG.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
THB_TRACE=1 timeout 1500 python bench.py --breakdown $EXTRA > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)  frac(all) %.4f  dom %s %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["all_kernels"]["frac"], d["roofline"]["kernel"], d["roofline"]["frac"]))
    print({k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms_per_step"].items()})
    print("host wall", {k: round(v, 3) for k, v in (d.get("host_wall_ms_per_call_kind") or {}).items()})
    print("cpu_baseline", {k: v for k, v in d.get("cpu_baseline", {}).items() if k != "sample"})
    print("drop_in_cli", {k: v for k, v in d.get("drop_in_cli", {}).items() if k not in ("note",)})
    print("parity", d.get("parity_checked"))
except Exception as e:
    print("no line:", e)
PY
grep "thb trace" gpurun_out/bench_$TAG.err | tail -3
tail -5 gpurun_out/bench_$TAG.err | cut -c1-300
