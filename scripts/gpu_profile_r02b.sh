#!/bin/bash
# Round-2 evidence beyond the forward kernels (each ncu run only after the same command exited 0 without ncu):
# launch lists and full captures of the training step (configs[4] shard of 512 windows, and 4096 windows per GPU)
# and of the CSR path (configs[3]).
mkdir -p gpurun_out
for B in 512 4096; do
  python scripts/train_step_once.py 2 $B > gpurun_out/train_plain_$B.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_$B.csv \
      python scripts/train_step_once.py 2 $B > gpurun_out/train_ncu_$B.log 2>&1
  echo "train launch list B=$B rc=$?"
done
# 13 of a step's launches match the filter; capture the second step
python scripts/train_step_once.py 2 512 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gcn_rows_kernel|inproj_kernel|gru_recur|mse_grad_kernel|gru_bwd|sgemm_kernel|gemm_kb_kernel|rows_to_tiles_kernel|gcn_bwd_rows_kernel|gcn_bwd_finish_kernel|adam_kernel' -s 13 -c 13 -f -o gpurun_out/prof_train_512 \
    python scripts/train_step_once.py 2 512 > gpurun_out/ncu_train_512.log 2>&1
echo "train capture 512 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'gcn_bwd_rows_kernel|gemm_kb_kernel|gru_bwd|sgemm_kernel' -s 4 -c 4 -f \
    -o gpurun_out/prof_train_4096 python scripts/train_step_once.py 2 4096 > gpurun_out/ncu_train_4096.log 2>&1
echo "train capture 4096 rc=$?"
python bench.py --workload fwd4096 --steps 1 --warmup 3 > gpurun_out/sparse_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/sparse_launches.csv \
    python bench.py --workload fwd4096 --steps 1 --warmup 3 > gpurun_out/ncu_sparse_l.log 2>&1
echo "sparse launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'gcn_sparse_plan_kernel|fmajor_to_tiles_kernel|gcn_sparse_row_kernel' -s 2 -c 2 -f \
    -o gpurun_out/prof_sparse python bench.py --workload fwd4096 --steps 1 --warmup 3 > gpurun_out/ncu_sparse.log 2>&1
echo "sparse capture rc=$?"
ls -la gpurun_out/*.ncu-rep
