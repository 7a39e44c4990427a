#!/bin/bash
# Runs ON the GPU box after gpu_profile_r02.sh / gpu_profile_r02b.sh: turns the .ncu-rep captures into the small
# tracked summaries (gpurun merges at most 64 MiB back), then drops the captures.
set -x
mkdir -p gpurun_out/profiles_out
python scripts/make_profile_summary.py r02 prof.ncu-rep launches.csv > /dev/null
python scripts/make_profile_summary.py r02_tensor prof_tensor.ncu-rep launches_tensor.csv > /dev/null
python scripts/ncu_table.py gpurun_out/prof_train_512.ncu-rep profiles/r02_train_ncu_summary \
  "ncu summary r02 — training step, 512 windows of the 34-station model (BASELINE configs[4] per-GPU shard); command: python scripts/train_step_once.py 2 512" > /dev/null
python scripts/ncu_table.py gpurun_out/prof_train_4096.ncu-rep profiles/r02_train4096_ncu_summary \
  "ncu summary r02 — backward kernels of the training step at 4096 windows per GPU; command: python scripts/train_step_once.py 2 4096" > /dev/null
python scripts/ncu_table.py gpurun_out/prof_sparse.ncu-rep profiles/r02_sparse_ncu_summary \
  "ncu summary r02 — CSR path, 4096-station kNN graph, hidden 128, T=24, 256 windows (BASELINE configs[3]); command: python bench.py --workload fwd4096 --steps 1 --warmup 3" > /dev/null
cp gpurun_out/launches.csv profiles/r02_launches_raw.csv
cp gpurun_out/launches_tensor.csv profiles/r02_launches_tensor_raw.csv
for B in 512 4096; do cp gpurun_out/train_launches_$B.csv profiles/r02_train_launches_${B}_raw.csv; done
cp gpurun_out/sparse_launches.csv profiles/r02_sparse_launches_raw.csv
cp profiles/r02_* gpurun_out/profiles_out/
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
