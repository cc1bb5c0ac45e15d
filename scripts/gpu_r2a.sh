#!/bin/bash
# round-2 first GPU session: parity suite, A/B of the tile kernels against round 1's queue kernels, hg38 default line
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi_$TAG.txt; free -g >> gpurun_out/smi_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)  frac(all) %.4f  dom %s %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["all_kernels"]["frac"], d["roofline"]["kernel"], d["roofline"]["frac"]))
    print({k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms_per_step"].items()})
    print({k: round(v, 3) for k, v in d["roofline"]["non_kernel_ms_per_step"].items()})
    print(d.get("parity_checked"))
except Exception as e:
    print("no line:", e)
PY
}
for mode in tile legacy; do
  if [ $mode = legacy ]; then export THB_JOIN_LEGACY=1 THB_SCAN_LEGACY=1; else unset THB_JOIN_LEGACY THB_SCAN_LEGACY; fi
  timeout 600 python bench.py --workload chr20 --pairs 4000000 --steps 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_chr20_$mode.json 2> gpurun_out/bench_${TAG}_chr20_$mode.err; echo "bench chr20 $mode exit $?"
  show gpurun_out/bench_${TAG}_chr20_$mode.json
done
unset THB_JOIN_LEGACY THB_SCAN_LEGACY
timeout 900 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_hg38.json 2> gpurun_out/bench_${TAG}_hg38.err; echo "bench hg38 exit $?"
show gpurun_out/bench_${TAG}_hg38.json
tail -3 gpurun_out/bench_${TAG}_hg38.err
