#!/bin/bash
# flank matcher: GPU parity tests + probe at scale.  usage: scripts/gpu_flank.sh <tag> [pairs] [workload]
TAG=$1; PAIRS=${2:-1000000}; WL=${3:-hg38}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flank.py -m gpu -x -q > gpurun_out/pytest_flank_$TAG.log 2>&1; tail -3 gpurun_out/pytest_flank_$TAG.log
timeout 1500 python scripts/flank_probe.py --workload $WL --pairs $PAIRS --out gpurun_out/flank_$TAG.json > gpurun_out/flank_$TAG.log 2>&1; echo "probe exit $?"
grep "probe\]" gpurun_out/flank_$TAG.log | cut -c1-700; tail -4 gpurun_out/flank_$TAG.log | cut -c1-400
