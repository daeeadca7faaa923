#!/bin/bash
# round 2, second session: launch list of one cfg-3 block + ncu --set full of the SYRK (one launch over 16 microphones with the
# fused tree sum) and of the rolled Cholesky diagonal-block kernel
set -u
mkdir -p gpurun_out
OB="python scripts/one_block.py cfg3 3"
$OB > gpurun_out/plain_ob3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4000 --csv --log-file gpurun_out/launches_r02b_cfg3.csv $OB > gpurun_out/ncu_l3.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:syrk_toeplitz -s 2 -c 1 -f -o gpurun_out/prof_r02b_syrk $OB > gpurun_out/ncu_full_syrk3.log 2>&1
echo "full syrk rc=$?"
ncu --set full --clock-control none --import-source on -k regex:chol_diag -s 140 -c 1 -f -o gpurun_out/prof_r02b_chol_diag $OB > gpurun_out/ncu_full_chol3.log 2>&1
echo "full chol_diag rc=$?"
