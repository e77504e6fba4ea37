#!/bin/bash
# round 2, GPU call O (8 GPUs): finer gradient buckets — 2-rank equivalence test, then the scaling command at N = 8, 2, 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_dp_cuda.py -m gpu -q ) > $O/r02o_pytest_dp.log 2>&1
echo "pytest rc=$?" >> $O/r02o_pytest_dp.log
cp $O/parity.log $O/r02o_parity.log 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02o_bench_dp8.json 2> $O/r02o_bench_dp8.err
echo "rc8=$?" >> $O/r02o_bench_dp8.err
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29762 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02o_bench_dp2.json 2> $O/r02o_bench_dp2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02o_bench_n1.json 2> $O/r02o_bench_n1.err
tail -3 $O/r02o_pytest_dp.log; cat $O/r02o_parity.log; for f in dp8 dp2 n1; do head -c 230 $O/r02o_bench_$f.json | tail -c 150; echo; done; tail -2 $O/r02o_bench_dp8.err
