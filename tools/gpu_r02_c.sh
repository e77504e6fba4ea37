#!/bin/bash
# round 2, GPU call C: per-stage backward errors at 512x512 with the device's gates imposed on the oracle; the tests that
# failed in call B with their assertion output; ncu of the HBM-bound kernels (CSV only: the .ncu-rep stays on the box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( timeout 900 python tests/tools/diag_engine.py 512 1 ) > $O/r02c_diag_512_b1.log 2>&1
( timeout 300 python tests/tools/diag_engine.py 128 2 ) > $O/r02c_diag_128_b2.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -k "graph or to_one or pack_table or compact or baseline or stacked or head" ) > $O/r02c_pytest_new.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -x -k "step_matches_oracle_512 and eval" ) > $O/r02c_pytest_512.log 2>&1
cp $O/parity.log $O/r02c_parity.log 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02c_bench.json 2> $O/r02c_bench.err
SN_HEAD_STACKED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02c_bench_head_unstacked.json 2>> $O/r02c_bench.err
SN_NO_GRAPH=1 SN_TRACE=1 timeout 300 python tools/profile_step.py > $O/r02c_plan_trace.txt 2>&1
SN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02c_launches_warp_step.csv python tools/profile_step.py > $O/r02c_profile_step.log 2>&1
SN_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'norm_act_fwd_v4|norm_act_bwd_apply_v4|norm_act_bwd_reduce_v4|pack_concat|plane_stats_kernel|pack_weights_multi|weight_scale_multi|adamw|ce_tanh|to_one|sum_grads|bias_grad_v8' \
  -c 70 -o /tmp/r02c_elementwise python tools/profile_step.py > $O/r02c_ncu_elementwise.log 2>&1
ncu -i /tmp/r02c_elementwise.ncu-rep --page raw --csv > $O/r02c_ncu_elementwise_raw.csv 2>/dev/null
ls -la /tmp/r02c_elementwise.ncu-rep >> $O/r02c_ncu_elementwise.log
tail -45 $O/r02c_diag_512_b1.log; tail -30 $O/r02c_pytest_new.log; head -c 400 $O/r02c_bench.json; echo; head -c 400 $O/r02c_bench_head_unstacked.json
