#!/bin/bash
# run S (single GPU, ~1.5 min): ncu of the augmentation kernel (launch list + --set full raw page of the 4 passes)
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:augment_pass -c 12 --csv \
  --log-file gpurun_out/r02_launches_augment_s.csv python tools/bench_augment.py --reps 1 --cpu-samples 0 > gpurun_out/r02_ncu_s.log 2>&1
echo "launch list rc=$?"
timeout 160 ncu --set full --clock-control none --import-source on -k regex:augment_pass -c 4 -o /tmp/aug_s \
  python tools/bench_augment.py --reps 1 --cpu-samples 0 >> gpurun_out/r02_ncu_s.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/aug_s.ncu-rep --page raw --csv > gpurun_out/r02_ncu_augment_raw.csv 2>> gpurun_out/r02_ncu_s.log
ls -la gpurun_out | tail -5
