#!/bin/bash
# ncu evidence for the three hot kernels (run only after the plain command exits 0).
mkdir -p gpurun_out
python scripts/profile_step.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1
python scripts/profile_step.py 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gcn_kernel|inproj_kernel|gru_recur_kernel' -s 3 -c 3 -f -o gpurun_out/prof python scripts/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
python scripts/profile_step.py 2 4096 tf32x3 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'inproj_tc_kernel' -s 1 -c 1 -f -o gpurun_out/prof_tc python scripts/profile_step.py 2 4096 tf32x3 > gpurun_out/ncu_tc.log 2>&1
echo "rc_tc=$?"; ls -la gpurun_out | grep ncu-rep
