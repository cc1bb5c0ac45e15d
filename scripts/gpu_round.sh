#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu full capture of the scan kernel.
# usage: scripts/gpu_round.sh <tag> [pairs]
TAG=${1:-r1}; PAIRS=${2:-10000000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --pairs $PAIRS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cat gpurun_out/bench_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --pairs $PAIRS --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bundle_kernel|hit_kernel|rescue_kernel|rescued_windows_kernel|window_scan_kernel|indel_kernel|chain_enum_kernel|chain_merge" -s 22 -c 22 -f -o gpurun_out/prof_$TAG \
  python bench.py --pairs $PAIRS --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm exit $?"; cut -c1-600 gpurun_out/bench_ref_$TAG.json
