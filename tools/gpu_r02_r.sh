#!/bin/bash
# run R (single GPU, ~4 min of commands): the device-side augmentation (f4) — parity tests, smoke of the rebuilt
# library, the bench line with the augmentation inside the e2e region, kernel timing, memcheck of the new kernel
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_augment_gpu.py -x -q > gpurun_out/r02_gpu_tests_augment_r.log 2>&1; echo "augment tests rc=$?"
tail -3 gpurun_out/r02_gpu_tests_augment_r.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_r.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_r.log
timeout 60 python tools/bench_augment.py > gpurun_out/r02_augment_timing_r.json 2> gpurun_out/r02_augment_timing_r.err; echo "timing rc=$?"; cat gpurun_out/r02_augment_timing_r.json
timeout 200 python bench.py --device-augment --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_r_device_augment.json 2> gpurun_out/r02_bench_r.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_r_device_augment.json"))
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_augment_gpu.py -q -k "dense_source or golden" > gpurun_out/r02_sanitizer_memcheck_augment.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck_augment.log
