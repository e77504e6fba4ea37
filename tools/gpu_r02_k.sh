#!/bin/bash
# round 2, GPU call K (2 GPUs): 2-rank CUDA DP equivalence test, 2-GPU bench lines (NCCL CTA cap A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
nvidia-smi -L > $O/r02k_gpus.txt 2>&1
( timeout 900 python -m pytest tests/test_dp_cuda.py -m gpu -q ) > $O/r02k_pytest_dp.log 2>&1
echo "pytest rc=$?" >> $O/r02k_pytest_dp.log
cp $O/parity.log $O/r02k_parity.log 2>/dev/null
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 10 --warmup 3"
timeout 400 $RUN > $O/r02k_bench_dp2.json 2> $O/r02k_bench_dp2.err
NCCL_MAX_CTAS=64 timeout 400 $RUN > $O/r02k_bench_dp2_nccl64.json 2>> $O/r02k_bench_dp2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02k_bench_n1_same_box.json 2>> $O/r02k_bench_dp2.err
tail -5 $O/r02k_pytest_dp.log; cat $O/r02k_parity.log; head -c 300 $O/r02k_bench_dp2.json; echo; head -c 300 $O/r02k_bench_dp2_nccl8.json; echo; head -c 300 $O/r02k_bench_n1_same_box.json; tail -5 $O/r02k_bench_dp2.err
