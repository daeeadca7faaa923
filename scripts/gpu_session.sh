#!/bin/bash
# One GPU session: the -m gpu suite, then a short bench run; logs under gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
tag=${1:-s}
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -n 30 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","e2e_per_call","e2e_batched","boundary_check","stage_ms_sequential_block","stage_ms_pipelined_last_block","alt_structured_stats","clocks"):
        print(k, d.get(k))
    for k in d:
        if k.startswith("roofline"):
            r=d[k]; print(k, r.get("achieved"), r.get("unit"), r.get("frac"), r.get("ms_per_block", r.get("kernel_ms_per_block")))
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("no bench json", e)
PY
