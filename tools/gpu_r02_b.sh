#!/bin/bash
# round 2, GPU call B: where do the 512x512 step gradients diverge from the fp64 oracle? + full GPU suite + bench A/B +
# ncu of the HBM-bound kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( timeout 900 python tests/tools/diag_engine.py 512 1 ) > $O/r02b_diag_512_b1.log 2>&1
( timeout 300 python tests/tools/diag_engine.py 128 1 ) > $O/r02b_diag_128_b1.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > $O/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02b_pytest.log
cp $O/parity.log $O/r02b_parity.log 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench.json 2> $O/r02b_bench.err
SN_NO_GRAPH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_nograph.json 2>> $O/r02b_bench.err
SN_NO_GRAPH=1 SN_PACK_PER_LAYER=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_nograph_perlayerpack.json 2>> $O/r02b_bench.err
SN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02b_launches_warp_step.csv python tools/profile_step.py > $O/r02b_profile_step.log 2>&1
SN_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'norm_act_fwd_v4|norm_act_bwd_apply_v4|norm_act_bwd_reduce_v4|pack_concat|plane_stats_kernel|pack_weights_multi|weight_scale_multi|adamw|ce_tanh|to_one|sum_grads|bias_grad_v8' \
  -c 70 -o $O/r02b_elementwise python tools/profile_step.py > $O/r02b_ncu_elementwise.log 2>&1
tail -40 $O/r02b_diag_512_b1.log; tail -15 $O/r02b_pytest.log; head -c 600 $O/r02b_bench.json; echo; head -c 300 $O/r02b_bench_nograph.json
