#!/bin/bash
# round 2, second session: the driver's sequence on one GPU (tests, smoke, bench as the driver runs it, cfg-2 / cfg-4 lines)
set -u
mkdir -p gpurun_out
tag=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/pytest_$tag.log; cat gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg3_$tag.json 2> gpurun_out/bench_cfg3_$tag.err; echo "bench rc=$?"
timeout 300 python bench.py --workload cfg2 --steps 40 --warmup 5 > gpurun_out/bench_cfg2_$tag.json 2> gpurun_out/bench_cfg2_$tag.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/bench_cfg4_$tag.json 2> gpurun_out/bench_cfg4_$tag.err; echo "cfg4 rc=$?"
python - <<PY
import json
for w in ("cfg3", "cfg2", "cfg4"):
    try:
        d = json.loads(open("gpurun_out/bench_%s_$tag.json" % w).read().strip().splitlines()[-1])
        print(w, d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
        if w == "cfg3":
            for k in ("e2e_per_call", "e2e_batched", "alt_structured_stats", "cpu_baseline", "stage_ms_sequential_block", "roofline_tridiag"):
                print("  ", k, d.get(k))
    except Exception as e:
        print(w, "no line", e)
PY
