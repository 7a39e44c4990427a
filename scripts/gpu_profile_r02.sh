#!/bin/bash
# Round-2 ncu evidence (each ncu run only after the same command exited 0 without ncu).
mkdir -p gpurun_out
# 1. launch list of the forward step, FP32 path and tensor path (3 forwards each)
python scripts/profile_step.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1
echo "launch list fp32 rc=$?"
python scripts/profile_step.py 3 4096 tensor > gpurun_out/plain_t.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tensor.csv python scripts/profile_step.py 3 4096 tensor > gpurun_out/ncu_launch_t.log 2>&1
echo "launch list tensor rc=$?"
# 2. full set: the FP32 path's three kernels, then the tensor path's two
python scripts/profile_step.py 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gcn_rows_kernel|inproj_kernel|gru_recur_unit_kernel' -s 3 -c 3 -f -o gpurun_out/prof python scripts/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1
echo "full fp32 rc=$?"
python scripts/profile_step.py 2 4096 tensor > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'inproj_tc2_kernel|gru_recur_tc_kernel' -s 2 -c 2 -f -o gpurun_out/prof_tensor python scripts/profile_step.py 2 4096 tensor > gpurun_out/ncu_full_t.log 2>&1
echo "full tensor rc=$?"; ls -la gpurun_out/*.ncu-rep
