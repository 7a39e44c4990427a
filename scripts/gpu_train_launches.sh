mkdir -p gpurun_out
for B in 512 4096; do
  python scripts/train_step_once.py 2 $B > gpurun_out/train_plain_$B.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_$B.csv \
      python scripts/train_step_once.py 2 $B > gpurun_out/train_ncu_$B.log 2>&1
  echo "train launch list B=$B rc=$?"
done
