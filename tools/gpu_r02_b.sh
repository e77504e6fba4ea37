#!/bin/bash
# round 2, GPU call B: where do the 512x512 step gradients diverge from the fp64 oracle? + the tests added since call A
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( timeout 900 python tests/tools/diag_engine.py 512 1 ) > $O/r02b_diag_512_b1.log 2>&1
( timeout 300 python tests/tools/diag_engine.py 128 1 ) > $O/r02b_diag_128_b1.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_engine_gpu.py::test_warp_model_step_matches_oracle_512 --deselect tests/test_engine_gpu.py::test_texture_model_step_matches_oracle_512 -k "baseline or to_one or graph or compact or train_loop or two_steps or ce_mode or step_matches_oracle or b16" ) > $O/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02b_pytest.log
cp $O/parity.log $O/r02b_parity.log 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench.json 2> $O/r02b_bench.err
SN_NO_GRAPH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_nograph.json 2>> $O/r02b_bench.err
tail -40 $O/r02b_diag_512_b1.log; tail -15 $O/r02b_pytest.log; head -c 600 $O/r02b_bench.json; echo; head -c 300 $O/r02b_bench_nograph.json
