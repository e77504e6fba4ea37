#!/bin/bash
# round 2, GPU call I: final tree — full GPU suite, bench lines, launch list, plan trace, ncu --set full of one resblock
# dgrad launch (tap GEMM) and one resblock wgrad launch, ncu of the HBM-bound kernels (CSV)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > $O/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02i_pytest.log
cp $O/parity.log $O/r02i_parity.log 2>/dev/null
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02i_bench.json 2> $O/r02i_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --model texture --perceptual --no-cpu-baseline > $O/r02i_bench_texture_perceptual.json 2>> $O/r02i_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --model joint --perceptual --no-cpu-baseline > $O/r02i_bench_joint.json 2>> $O/r02i_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --fp32-inputs --no-cpu-baseline > $O/r02i_bench_fp32_inputs.json 2>> $O/r02i_bench.err
SN_NO_WGRAD_OVERLAP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02i_bench_no_wgrad_overlap.json 2>> $O/r02i_bench.err
SN_NO_GRAPH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02i_bench_nograph.json 2>> $O/r02i_bench.err
SN_NO_GRAPH=1 SN_TRACE=1 timeout 300 python tools/profile_step.py > $O/r02i_plan_trace.txt 2>&1
SN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02i_launches_warp_step.csv python tools/profile_step.py > $O/r02i_profile_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tap_gemm_kernel --launch-skip 8 --launch-count 1 \
  -o /tmp/r02i_tapgemm python tools/bench_layer.py conv3r 16 1024 1024 32 32 3 > $O/r02i_ncu_tapgemm.log 2>&1
ncu -i /tmp/r02i_tapgemm.ncu-rep --page raw --csv > $O/r02i_ncu_tapgemm_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm_kernel --launch-skip 2 --launch-count 1 \
  -o /tmp/r02i_wgrad python tools/bench_layer.py conv3r 16 1024 1024 32 32 3 > $O/r02i_ncu_wgrad.log 2>&1
ncu -i /tmp/r02i_wgrad.ncu-rep --page raw --csv > $O/r02i_ncu_wgrad_raw.csv 2>/dev/null
SN_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'norm_act_fwd_v4|norm_act_bwd_apply_v4|norm_act_bwd_reduce_v4|pack_concat|plane_stats_kernel|pack_weights_multi|weight_scale_multi|adamw|ce_tanh|to_one|sum_grads|bias_grad' \
  -c 70 -o /tmp/r02i_elementwise python tools/profile_step.py > $O/r02i_ncu_elementwise.log 2>&1
ncu -i /tmp/r02i_elementwise.ncu-rep --page raw --csv > $O/r02i_ncu_elementwise_raw.csv 2>/dev/null
tail -5 $O/r02i_pytest.log; head -c 500 $O/r02i_bench.json
