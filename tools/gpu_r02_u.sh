#!/bin/bash
# run U (single GPU, < 1 min): the augmentation GPU tests incl. the `--dataset warp_b200` batch format through set_input
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_augment_gpu.py -x -q > gpurun_out/r02_gpu_tests_augment_u.log 2>&1; echo "augment tests rc=$?"
tail -15 gpurun_out/r02_gpu_tests_augment_u.log
