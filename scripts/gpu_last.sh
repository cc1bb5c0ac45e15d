#!/bin/bash
# last check of the round on one GPU: whole GPU suite, smoke(), the default bench line, the reference arm line
TAG=${1:-last}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log | cut -c1-200
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_$TAG.json") if l.startswith("{")][-1])
print("value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)  frac(all) %.4f  dom %s %.3f traffic %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["all_kernels"]["frac"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"]))
print("cpu_baseline", {k: v for k, v in d.get("cpu_baseline", {}).items() if k != "sample"})
print("drop_in_cli", {k: v for k, v in d.get("drop_in_cli", {}).items() if k in ("value", "wall_s", "startup_s", "reference_wall_s", "speedup_wall")})
print("flank", d["flank_match"]["index"], d["flank_match"]["reads_per_s"], d["flank_match"].get("step_plus_junction_index"))
print("parity", d.get("parity_checked"))
PY
