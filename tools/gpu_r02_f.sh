#!/bin/bash
# round 2, GPU call F: full GPU suite (oracle gate path fixed), bench lines (warp / texture / joint / reference arm),
# launch list, plan trace, ncu of the HBM-bound kernels (CSV only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > $O/r02f_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02f_pytest.log
cp $O/parity.log $O/r02f_parity.log 2>/dev/null
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r02f_bench.json 2> $O/r02f_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --model texture --perceptual --no-cpu-baseline > $O/r02f_bench_texture_perceptual.json 2>> $O/r02f_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --model joint --perceptual --no-cpu-baseline > $O/r02f_bench_joint.json 2>> $O/r02f_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02f_bench_reference.json 2>> $O/r02f_bench.err
SN_NO_GRAPH=1 SN_TRACE=1 timeout 300 python tools/profile_step.py > $O/r02f_plan_trace.txt 2>&1
SN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02f_launches_warp_step.csv python tools/profile_step.py > $O/r02f_profile_step.log 2>&1
SN_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'norm_act_fwd_v4|norm_act_bwd_apply_v4|norm_act_bwd_reduce_v4|pack_concat|plane_stats_kernel|pack_weights_multi|weight_scale_multi|adamw|ce_tanh|to_one|sum_grads|bias_grad_v8' \
  -c 70 -o /tmp/r02f_elementwise python tools/profile_step.py > $O/r02f_ncu_elementwise.log 2>&1
ncu -i /tmp/r02f_elementwise.ncu-rep --page raw --csv > $O/r02f_ncu_elementwise_raw.csv 2>/dev/null
tail -12 $O/r02f_pytest.log; head -c 500 $O/r02f_bench.json; echo; head -c 300 $O/r02f_bench_texture_perceptual.json; echo; head -c 300 $O/r02f_bench_joint.json; echo; head -c 300 $O/r02f_bench_reference.json
