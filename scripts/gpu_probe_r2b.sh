#!/bin/bash
# round-2 probe: fused SYRK reduction (parity + timing), GEMM shapes, chase grid size
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python scripts/one_block.py cfg3 3 2>&1 | tail -1 | cut -c1-260
APV_SYRK_GROUP=4 timeout 200 python scripts/one_block.py cfg3 3 2>&1 | tail -1 | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k "cfg3 or cfg2" 2>&1 | grep -E "worst|passed|failed" | cut -c1-300
timeout 100 python scripts/bench_gemm_shapes.py
for g in 24 74; do APV_CHASE_G=$g APV_OB_STATS=2 APV_TS_DEBUG=1 timeout 200 python scripts/one_block.py cfg3 2 2>&1 | grep -E "two-stage ms: panel" | tail -1; done
