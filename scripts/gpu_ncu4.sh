#!/bin/bash
# final captures of the second session of round 2: launch list of three cfg-3 blocks + ncu --set full of the SYRK (ring of
# partial tiles, fused sum) and of the rank-64 band update GEMM (L2 prefetch of the C tile)
set -u
mkdir -p gpurun_out
OB="python scripts/one_block.py cfg3 3"
$OB > gpurun_out/plain_ob4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4000 --csv --log-file gpurun_out/launches_r02c_cfg3.csv $OB > gpurun_out/ncu_l4.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:syrk_toeplitz -s 2 -c 1 -f -o gpurun_out/prof_r02c_syrk $OB > gpurun_out/ncu_full_syrk4.log 2>&1
echo "full syrk rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:gemm_async_kernel<1, 64>" -s 1000 -c 1 -f -o gpurun_out/prof_r02c_gemm_band $OB > gpurun_out/ncu_full_gemm4.log 2>&1
echo "full gemm rc=$?"
