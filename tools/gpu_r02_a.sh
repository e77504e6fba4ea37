#!/bin/bash
# round 2, GPU call A: full GPU test suite (new 512x512 parity tests), bench lines, ncu of the HBM-bound kernels,
# compute-sanitizer on the kernel tests.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02a_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=20 ) > $O/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02a_pytest.log
cp $O/parity.log $O/r02a_parity.log 2>/dev/null
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r02a_bench.json 2> $O/r02a_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --fp32-inputs --no-cpu-baseline > $O/r02a_bench_fp32in.json 2>> $O/r02a_bench.err
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 > $O/r02a_bench_ref.json 2>> $O/r02a_bench.err
# launch list of one step + per-plan trace
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02a_launches_warp_step.csv python tools/profile_step.py > $O/r02a_profile_step.log 2>&1
SN_TRACE=1 timeout 300 python tools/profile_step.py > $O/r02a_plan_trace.txt 2>&1
# ncu --set full of the HBM-bound kernels (a few launches of each, inside one training step)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'norm_act_fwd_v4|norm_act_bwd_apply_v4|norm_act_bwd_reduce_v4|pack_concat|plane_stats_kernel|pack_weights_kernel|adamw' \
  -c 40 -o $O/r02a_elementwise python tools/profile_step.py > $O/r02a_ncu_elementwise.log 2>&1
# sanitizers on the kernel tests (small shapes)
K='test_conv_forward and nsplit-3 or test_conv_backward or test_instance_norm or test_residual or test_tanh or test_losses or test_roi'
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "($K) and not baseline" ) > $O/r02a_sanitizer_memcheck.log 2>&1
( time timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv3r-2-128-128 or convT4s2-2-128-64 or conv4s2-2-64-128 or instance_norm" ) > $O/r02a_sanitizer_racecheck.log 2>&1
( time timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv3r-2-128-128 or convT4s2-2-128-64 or conv4s2-2-64-128 or instance_norm" ) > $O/r02a_sanitizer_synccheck.log 2>&1
tail -5 $O/r02a_pytest.log; cat $O/r02a_bench.json | head -c 1500; echo; tail -3 $O/r02a_sanitizer_memcheck.log
