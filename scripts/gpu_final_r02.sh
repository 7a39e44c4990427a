#!/bin/bash
# Final evidence of a round: GPU tests, smoke, the N=1 bench line and reference arm, then the FP32 forward's ncu
# launch list and full capture (each ncu run only after the same command exited 0 without ncu) and the summaries.
mkdir -p gpurun_out/profiles_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/final_bench_ref_n1.json 2>/dev/null; echo "ref rc=$?"
python scripts/profile_step.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1
echo "launch list fp32 rc=$?"
python scripts/profile_step.py 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gcn_rows_kernel|inproj_kernel|gru_recur_unit_kernel' -s 3 -c 3 -f -o gpurun_out/prof python scripts/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1
echo "full fp32 rc=$?"
python scripts/make_profile_summary.py r02 prof.ncu-rep launches.csv > /dev/null
cp gpurun_out/launches.csv profiles/r02_launches_raw.csv
cp profiles/r02_ncu_summary.* profiles/r02_launches.csv profiles/r02_launches_raw.csv gpurun_out/profiles_out/
rm -f gpurun_out/*.ncu-rep
