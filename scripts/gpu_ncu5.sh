#!/bin/bash
# final launch list of round 2 (three cfg-3 blocks) + ncu --set full of the register-slab panel QR (first, largest panel)
set -u
mkdir -p gpurun_out
OB="python scripts/one_block.py cfg3 3"
$OB > gpurun_out/plain_ob5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4000 --csv --log-file gpurun_out/launches_r02d_cfg3.csv $OB > gpurun_out/ncu_l5.log 2>&1
echo "launch list rc=$?"
APV_OB_STATS=2 ncu --set full --clock-control none --import-source on -k regex:sb_panel_qr -s 5 -c 1 -f -o gpurun_out/prof_r02d_qr python scripts/one_block.py cfg3 1 > gpurun_out/ncu_qr5.log 2>&1
echo "full qr rc=$?"
