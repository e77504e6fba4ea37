#!/bin/bash
# round 2, GPU call P: dynamic tile schedule of the persistent tap GEMM — full suite, racecheck, A/B benches (static tiles,
# IN kernels compiled for 5 blocks/SM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $O/r02p_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02p_pytest.log
cp $O/parity.log $O/r02p_parity.log 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02p_bench.json 2> $O/r02p_bench.err
SN_TAP_STATIC_TILES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02p_bench_static_tiles.json 2>> $O/r02p_bench.err
SN_EW_MINB=5 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02p_bench_ew_minb5.json 2>> $O/r02p_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02p_bench_2.json 2>> $O/r02p_bench.err
( time timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "(test_conv_forward and nsplit-3 and (conv3r-2-128 or convT4s2-2-128 or conv4s2-2-64-128 or head-2-192 or conv4s1-2-128)) or fused_instance_norm" ) > $O/r02p_sanitizer_racecheck.log 2>&1
( time timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "(test_conv_forward and nsplit-3) or test_conv_backward" ) > $O/r02p_sanitizer_memcheck.log 2>&1
tail -4 $O/r02p_pytest.log; for f in bench bench_static_tiles bench_ew_minb5 bench_2; do head -c 230 $O/r02p_$f.json | tail -c 130; echo; done; grep -E "SUMMARY|passed" $O/r02p_sanitizer_racecheck.log $O/r02p_sanitizer_memcheck.log
