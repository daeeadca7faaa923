#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-s}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -p no:cacheprovider -k "full_spectrum or two_engines" 2>&1 | tail -n 4
timeout 300 $TR --master-port 29511 scripts/run_sharded_nccl.py small 24 2>&1 | grep -vE "^W|^\*|OMP_NUM" | tail -n 6
timeout 300 $TR --master-port 29512 scripts/run_sharded_nccl.py cfg2 16 2>&1 | grep -vE "^W|^\*|OMP_NUM" | tail -n 6
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu_$tag.json 2> gpurun_out/bench_2gpu_$tag.err
echo "bench 2gpu rc=$?"; grep -vE "^W|^\*|OMP_NUM" gpurun_out/bench_2gpu_$tag.err | tail -n 12
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_2gpu_$tag.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","boundary_check","collective","timed_region"):
        print(k, d.get(k))
except Exception as e:
    print("no bench json", e)
PY
