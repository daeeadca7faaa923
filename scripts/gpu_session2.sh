#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-s}
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -vE "^\s*$" > gpurun_out/pytest_$tag.log
echo "pytest done"; grep -E "passed|failed|FAILED|worst|Error" gpurun_out/pytest_$tag.log | tail -n 25
timeout 600 python bench.py --steps 8 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","e2e_per_call","e2e_batched","stage_ms_sequential_block"):
        print(k, d.get(k))
    for k in d:
        if k.startswith("roofline"):
            r=d[k]; print(k, r.get("achieved"), r.get("unit"), r.get("frac"), r.get("ms_per_block", r.get("kernel_ms_per_block")))
except Exception as e:
    print("no bench json", e)
PY
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/bench_cfg4_$tag.json 2> gpurun_out/bench_cfg4_$tag.err
echo "cfg4 rc=$?"; tail -c 1500 gpurun_out/bench_cfg4_$tag.err; head -c 2500 gpurun_out/bench_cfg4_$tag.json
APV_CFG5_CLIPS=2 timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 > gpurun_out/bench_cfg5_$tag.json 2> gpurun_out/bench_cfg5_$tag.err
echo "cfg5 rc=$?"; tail -c 1500 gpurun_out/bench_cfg5_$tag.err; head -c 3000 gpurun_out/bench_cfg5_$tag.json
