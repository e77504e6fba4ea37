#!/bin/bash
# round 2, GPU call L: final tree — full GPU suite, compute-sanitizer over the kernel tests (incl. the round-2 kernels),
# the driver's bench command and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > $O/r02l_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02l_pytest.log
cp $O/parity.log $O/r02l_parity.log 2>/dev/null
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02l_bench.json 2> $O/r02l_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02l_bench_reference.json 2>> $O/r02l_bench.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02l_smoke.log 2>&1
K='(test_conv_forward and nsplit-3) or test_conv_backward or test_instance_norm_block_fwd_bwd or test_residual or test_tanh or test_losses or test_roi or test_to_one or test_fused_bias or test_pack_planes'
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "($K) and not baseline" ) > $O/r02l_sanitizer_memcheck.log 2>&1
( time timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_engine_gpu.py -q -x -k "pack_table or compact or test_warp_engine_forward or test_graph" ) > $O/r02l_sanitizer_memcheck_engine.log 2>&1
( time timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv3r-2-128-128 or convT4s2-2-128-64 or conv4s2-2-64-128 or head-2-192 or test_to_one or test_fused_bias or (instance_norm and not baseline)" ) > $O/r02l_sanitizer_racecheck.log 2>&1
( time timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv3r-2-128-128 or convT4s2-2-128-64 or conv4s2-2-64-128 or head-2-192 or test_to_one or test_fused_bias or (instance_norm and not baseline)" ) > $O/r02l_sanitizer_synccheck.log 2>&1
tail -5 $O/r02l_pytest.log; head -c 400 $O/r02l_bench.json; echo; tail -3 $O/r02l_smoke.log; for t in memcheck memcheck_engine racecheck synccheck; do tail -6 $O/r02l_sanitizer_$t.log | head -3; done
