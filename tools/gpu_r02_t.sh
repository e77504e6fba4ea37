#!/bin/bash
# run T (single GPU, ~1 min): the restructured augmentation kernel (row loops, hoisted op constants) — parity tests,
# timing, launch list
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_augment_gpu.py -x -q > gpurun_out/r02_gpu_tests_augment_t.log 2>&1; echo "augment tests rc=$?"
tail -2 gpurun_out/r02_gpu_tests_augment_t.log
timeout 50 python tools/bench_augment.py > gpurun_out/r02_augment_timing_t.json 2> gpurun_out/r02_augment_timing_t.err; echo "timing rc=$?"; cat gpurun_out/r02_augment_timing_t.json
timeout 60 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:augment_pass -c 8 --csv \
  --log-file gpurun_out/r02_launches_augment_t.csv python tools/bench_augment.py --reps 1 --cpu-samples 0 > gpurun_out/r02_ncu_t.log 2>&1
echo "launch list rc=$?"; grep -c augment_pass gpurun_out/r02_launches_augment_t.csv
