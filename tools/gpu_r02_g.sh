#!/bin/bash
# round 2, GPU call G: unrolled IN/act kernels, direct pack_concat, 16-byte pack stores, to-one tuning: tests + A/B benches + ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/r02g_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02g_pytest.log
cp $O/parity.log $O/r02g_parity.log 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02g_bench.json 2> $O/r02g_bench.err
SN_EW_V4=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02g_bench_ew_v4.json 2>> $O/r02g_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --model texture --perceptual --no-cpu-baseline > $O/r02g_bench_texture_perceptual.json 2>> $O/r02g_bench.err
SN_NO_GRAPH=1 SN_TRACE=1 timeout 300 python tools/profile_step.py > $O/r02g_plan_trace.txt 2>&1
SN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02g_launches_warp_step.csv python tools/profile_step.py > $O/r02g_profile_step.log 2>&1
SN_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'norm_act_fwd_v4u|norm_act_bwd_apply_v4u|norm_act_bwd_reduce_v4u|pack_concat|plane_stats_kernel|pack_weights_multi|weight_scale_multi|adamw|ce_tanh|to_one|sum_grads|bias_grad_v8' \
  -c 70 -o /tmp/r02g_elementwise python tools/profile_step.py > $O/r02g_ncu_elementwise.log 2>&1
ncu -i /tmp/r02g_elementwise.ncu-rep --page raw --csv > $O/r02g_ncu_elementwise_raw.csv 2>/dev/null
tail -6 $O/r02g_pytest.log; head -c 400 $O/r02g_bench.json; echo; head -c 400 $O/r02g_bench_ew_v4.json
