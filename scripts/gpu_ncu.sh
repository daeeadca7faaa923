#!/bin/bash
# ncu captures of round 2 (one GPU).  Every profiled command line is first run plain.
set -u
mkdir -p gpurun_out
OB="python scripts/one_block.py cfg3 3"
$OB > gpurun_out/plain_ob.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 930 -c 1850 --csv --log-file gpurun_out/launches_r02_cfg3.csv $OB > gpurun_out/ncu_l1.log 2>&1
echo "launch list rc=$?"
for k in syrk_toeplitz sb2st_chase render_kernel sb_panel_qr; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r02_$k $OB > gpurun_out/ncu_full_$k.log 2>&1
  echo "full $k rc=$?"
done
# solver choice at the crossover size (cfg-2, n = 1024): one-stage (eig_mode 1) against two-stage (eig_mode 3)
for m in 1 3; do
  APV_OB_EIG_MODE=$m python scripts/one_block.py cfg2 3 > gpurun_out/plain_ob_cfg2_$m.log 2>&1 && \
  APV_OB_EIG_MODE=$m ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 3000 --csv --log-file gpurun_out/launches_r02_cfg2_eig$m.csv python scripts/one_block.py cfg2 3 > gpurun_out/ncu_l2_$m.log 2>&1
  echo "cfg2 eig_mode $m rc=$?"
done
# full-spectrum route (cfg-4: V = n = 4096): one block
APV_OB_V=0 python scripts/one_block.py cfg3 2 > gpurun_out/plain_ob_full.log 2>&1 && \
APV_OB_V=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4000 --csv --log-file gpurun_out/launches_r02_cfg4.csv python scripts/one_block.py cfg3 2 > gpurun_out/ncu_l3.log 2>&1
echo "cfg4 launch list rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02*
