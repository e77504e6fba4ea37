#!/bin/bash
# round 2, GPU call Q (8 GPUs): dynamic vs static tile schedule under NCCL contention at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench.py --gpus 8 --steps 10 --warmup 3"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29771 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02q_bench_dp8.json 2> $O/r02q_bench_dp8.err; echo "rc=$?" >> $O/r02q_bench_dp8.err
SN_TAP_STATIC_TILES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29772 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02q_bench_dp8_static_tiles.json 2> $O/r02q_bench_dp8_static.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02q_bench_n1.json 2> $O/r02q_bench_n1.err
for f in dp8 dp8_static_tiles n1; do head -c 230 $O/r02q_bench_$f.json | tail -c 150; echo; done; tail -2 $O/r02q_bench_dp8.err
