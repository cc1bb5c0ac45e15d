#!/bin/bash
# round-end evidence on one GPU: GPU suite, default bench line (+ cpu_baseline, drop_in_cli), chr20 / indel lines, ncu launch list of the default command
TAG=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 1200 python bench.py --breakdown > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 600 python bench.py --workload chr20 --steps 5 --no-cpu-baseline --breakdown > gpurun_out/bench_${TAG}_chr20.json 2> gpurun_out/bench_${TAG}_chr20.err; echo "chr20 exit $?"
timeout 600 python bench.py --workload indel --steps 5 --no-cpu-baseline --breakdown > gpurun_out/bench_${TAG}_indel.json 2> gpurun_out/bench_${TAG}_indel.err; echo "indel exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-pairs 0 > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu exit $?"
python - <<PY
import json
for t in ("", "_chr20", "_indel"):
    try:
        d = json.load(open("gpurun_out/bench_$TAG%s.json" % t))
        print(t or "hg38", "value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)  frac(all) %.4f  dom %s %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["all_kernels"]["frac"], d["roofline"]["kernel"], d["roofline"]["frac"]))
        print("   ", {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms_per_step"].items()})
        if "cpu_baseline" in d: print("    cpu_baseline", {k: v for k, v in d["cpu_baseline"].items() if k != "sample"})
        if "drop_in_cli" in d: print("    drop_in_cli", {k: v for k, v in d["drop_in_cli"].items() if k in ("value", "wall_s", "startup_s", "reference_wall_s", "speedup_wall")})
        print("    parity", d.get("parity_checked"))
    except Exception as e:
        print(t, "no line:", e)
PY
