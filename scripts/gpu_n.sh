#!/bin/bash
# bench at N ranks with the per-call host breakdown.  usage: scripts/gpu_n.sh <tag> <N> [extra bench args]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 --breakdown "$@" > gpurun_out/bench_${TAG}_n$N.out 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench N=$N exit $?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${TAG}_n$N.out") if l.startswith("{")][-1])
    print("N=%d value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print({k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms_per_step"].items()})
    print({k[:14]: round(v, 3) for k, v in d["roofline"]["non_kernel_ms_per_step"].items()})
    print("host wall", {k: round(v, 3) for k, v in (d.get("host_wall_ms_per_call_kind") or {}).items()})
    print("parity", d.get("parity_checked"))
except Exception as e:
    print("no line:", e)
PY
