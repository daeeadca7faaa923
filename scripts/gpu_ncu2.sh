#!/bin/bash
set -u
mkdir -p gpurun_out
OB="python scripts/one_block.py cfg3 3"
$OB > gpurun_out/plain_ob2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4000 --csv --log-file gpurun_out/launches_r02_cfg3_final.csv $OB > gpurun_out/ncu_l1f.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:syrk_toeplitz -s 2 -c 1 -f -o gpurun_out/prof_r02_syrk_final $OB > gpurun_out/ncu_full_syrk_final.log 2>&1
echo "full syrk rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dc_secular -s 6 -c 1 -f -o gpurun_out/prof_r02_dc_secular env APV_OB_V=0 python scripts/one_block.py cfg3 1 > gpurun_out/ncu_full_dc.log 2>&1
echo "full dc rc=$?"
