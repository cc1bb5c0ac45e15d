#!/bin/bash
# multi-GPU session (gpurun --gpus N): parity of the sharded path on both stages (incl. --fusion-search), then the default bench at N ranks
TAG=${1:-multi}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nvidia-smi topo -m >> gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_$TAG.log
tail -3 gpurun_out/pytest_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_rank_check.py > gpurun_out/multi_rank_check_$TAG.log 2>&1; echo "multi_rank_check exit $?"
grep -E "rank 0\]|MULTI|MISMATCH" gpurun_out/multi_rank_check_$TAG.log | tail -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench N=$N exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_n$N.json"))
    print("N=%d value %.4g reads/s  step %.3f ms  e2e %.4g reads/s (%.1f ms)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("parity", d.get("parity_checked"))
except Exception as e:
    print("no line:", e)
PY
tail -4 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
