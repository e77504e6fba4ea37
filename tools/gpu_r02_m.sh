#!/bin/bash
# round 2, GPU call M: InstanceNorm statistics fused into the GEMM epilogue — tests, A/B bench, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/r02m_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02m_pytest.log
cp $O/parity.log $O/r02m_parity.log 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02m_bench.json 2> $O/r02m_bench.err
SN_NO_FUSED_STATS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02m_bench_no_fused_stats.json 2>> $O/r02m_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02m_bench_2.json 2>> $O/r02m_bench.err
SN_NO_GRAPH=1 SN_TRACE=1 timeout 300 python tools/profile_step.py > $O/r02m_plan_trace.txt 2>&1
SN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02m_launches_warp_step.csv python tools/profile_step.py > $O/r02m_profile_step.log 2>&1
( time timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "fused_instance_norm" ) > $O/r02m_sanitizer_memcheck_fused_stats.log 2>&1
( time timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "fused_instance_norm" ) > $O/r02m_sanitizer_racecheck_fused_stats.log 2>&1
tail -5 $O/r02m_pytest.log; for f in bench bench_no_fused_stats bench_2; do head -c 260 $O/r02m_$f.json | tail -c 120; echo; done; grep fused_in $O/r02m_parity.log; tail -3 $O/r02m_sanitizer_memcheck_fused_stats.log; tail -3 $O/r02m_sanitizer_racecheck_fused_stats.log
