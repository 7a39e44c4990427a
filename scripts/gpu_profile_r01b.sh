#!/bin/bash
# Round-1 evidence beyond the forward kernels: launch list of the bench command itself, full captures of the
# training-step kernels (configs[4] shard: 512 windows) and of the scaled-shape GCN kernel (configs[3]).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "bench launch list rc=$?"
python scripts/train_step_once.py 2 512 > gpurun_out/train_plain_512.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sgemm_kernel|gcn_bwd_kernel|gru_bwd_kernel|gru_recur|adam_kernel|mse_grad|inproj_kernel|rows_to_tiles|gcn_kernel' \
    -s 12 -c 12 -f -o gpurun_out/prof_train python scripts/train_step_once.py 2 512 > gpurun_out/ncu_train.log 2>&1
echo "train capture rc=$?"
python bench.py --workload fwd4096 --steps 1 --warmup 3 > gpurun_out/sparse_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gcn_sparse_row_kernel|rows_to_tiles' -s 2 -c 2 -f \
    -o gpurun_out/prof_sparse python bench.py --workload fwd4096 --steps 1 --warmup 3 > gpurun_out/ncu_sparse.log 2>&1
echo "sparse capture rc=$?"
ls -la gpurun_out/*.ncu-rep
