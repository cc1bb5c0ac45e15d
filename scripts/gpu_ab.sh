#!/bin/bash
# A/B of kernel variants + ncu capture.  usage: scripts/gpu_ab.sh <tag> <pairs> <workload> "<variant env lists separated by ;>" [ncu kernel regex]
TAG=$1; PAIRS=${2:-4000000}; WL=${3:-chr20}; VARIANTS=${4:-"none"}; KRE=$5
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)  frac(all) %.4f  dom %s %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["all_kernels"]["frac"], d["roofline"]["kernel"], d["roofline"]["frac"]))
    print({k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms_per_step"].items()})
    print({k: round(v, 3) for k, v in d["roofline"]["per_kernel_frac"].items()})
    print({k[:14]: round(v, 3) for k, v in d["roofline"]["non_kernel_ms_per_step"].items()}); print("host wall per call kind", {k: round(v, 3) for k, v in (d.get("host_wall_ms_per_call_kind") or {}).items()})
except Exception as e:
    print("no line:", e)
PY
}
IFS=';' read -ra VS <<< "$VARIANTS"
i=0
for v in "${VS[@]}"; do
  i=$((i+1))
  ( if [ "$v" != "none" ]; then for kv in $v; do export "$kv"; done; fi
    timeout 900 python bench.py --workload $WL --pairs $PAIRS --steps 5 --no-cpu-baseline --parity-pairs 50000 --breakdown $THB_BENCH_EXTRA > gpurun_out/bench_${TAG}_v$i.json 2> gpurun_out/bench_${TAG}_v$i.err; echo "variant $i [$v] exit $?" )
  show gpurun_out/bench_${TAG}_v$i.json
done
if [ -n "$KRE" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 8 -c 8 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload $WL --pairs $PAIRS --steps 1 --warmup 1 --no-cpu-baseline --parity-pairs 0 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
  ls -la gpurun_out/prof_$TAG.ncu-rep
fi
