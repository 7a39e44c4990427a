#!/bin/bash
# A/B (bit-identity + stage times) of the FP32 kernel generations, then ncu --set full of the three hot kernels.
mkdir -p gpurun_out
timeout 300 python scripts/kernel_ab.py --quick > gpurun_out/ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/ab.log
python scripts/profile_step.py 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gcn_rows_kernel|inproj_kernel|gru_recur_unit_kernel' -s 3 -c 3 -f -o gpurun_out/prof_r02 python scripts/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
