#!/bin/bash
# round 2, GPU call N (8 GPUs): the driver's scaling command at N = 8 and N = 4 on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02n_gpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02n_bench_dp8.json 2> $O/r02n_bench_dp8.err
echo "rc8=$?" >> $O/r02n_bench_dp8.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29752 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r02n_bench_dp4.json 2> $O/r02n_bench_dp4.err
echo "rc4=$?" >> $O/r02n_bench_dp4.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02n_bench_n1.json 2> $O/r02n_bench_n1.err
head -c 330 $O/r02n_bench_dp8.json; echo; head -c 330 $O/r02n_bench_dp4.json; echo; head -c 330 $O/r02n_bench_n1.json; echo; tail -3 $O/r02n_bench_dp8.err
