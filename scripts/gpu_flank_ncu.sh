#!/bin/bash
# flank matcher: probe with the sample parity check, then an ncu --set full capture of its kernels.  usage: scripts/gpu_flank_ncu.sh <tag> [pairs]
TAG=$1; PAIRS=${2:-1000000}
mkdir -p gpurun_out
timeout 1500 python scripts/flank_probe.py --workload hg38 --pairs $PAIRS --check-reads 200 --out gpurun_out/flank_$TAG.json > gpurun_out/flank_$TAG.log 2>&1; echo "probe exit $?"
grep "probe\]" gpurun_out/flank_$TAG.log | grep -v "step" | cut -c1-400; tail -3 gpurun_out/flank_$TAG.log | cut -c1-300
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"flank_" -c 9 -f -o gpurun_out/prof_$TAG \
  python scripts/flank_probe.py --workload hg38 --pairs $PAIRS --steps 1 --check-reads 0 > gpurun_out/ncu_flank_$TAG.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/prof_$TAG.ncu-rep
