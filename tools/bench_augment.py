#!/usr/bin/env python
"""Timing of the device-side cloth augmentation next to the reference's host pipeline.

    python tools/bench_augment.py [--batch 16] [--size 512] [--reps 20]

Prints one JSON line: device ms per batch (CUDA events, label maps resident, op table H2D inside), achieved GB/s against
the algorithmic bytes (every executed pass reads and writes one fp32 plane; pass 0 reads the uint8 map), the host draw
time per sample, and the reference's `per_channel_transform` through Pillow/torchvision on this box's CPU (one core, as
one DataLoader worker runs it) for the same draws.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swapnet_b200 import data as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cpu-samples", type=int, default=2)
    a = ap.parse_args()
    from PIL import Image
    from torchvision import transforms as T

    B, S, C = a.batch, a.size, 19
    tf = T.RandomOrder([T.RandomVerticalFlip(), T.RandomHorizontalFlip(),
                        T.RandomAffine(degrees=10, translate=(0.1, 0.1), scale=(0.8, 1.2), shear=20), T.RandomPerspective()])
    g = torch.Generator().manual_seed(1)
    lab = torch.randint(0, C, (B, S // 16, S // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2).to(torch.uint8)
    aug = D.ClothAugmenter(tf, C)
    random.seed(1); torch.manual_seed(1)
    t0 = time.perf_counter()
    sample_ops = [aug.draw(S, S) for _ in range(B)]
    draw_ms = (time.perf_counter() - t0) * 1e3 / B
    table = D.OpTable(sample_ops)
    passes = sum(max(len(o), 1) for s in sample_ops for o in s)
    alg_bytes = passes * S * S * 8 - B * C * S * S * 3          # pass 0 reads 1 byte instead of 4
    dev = torch.device("cuda:0")
    lab_d = lab.to(dev)
    for _ in range(3):
        out = aug.apply(lab_d, table)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = aug.apply(lab_d, table)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    # the reference's host pipeline on the same draws (one core)
    torch.set_num_threads(1)
    onehot = lambda l: ((l[None].numpy() == np.arange(C).reshape(-1, 1, 1)) & (np.arange(C).reshape(-1, 1, 1) > 0)).astype(np.float32)
    random.seed(1); torch.manual_seed(1)
    t0 = time.perf_counter()
    same = True
    for b in range(min(a.cpu_samples, B)):
        planes = onehot(lab[b])
        ref = np.stack([np.array(tf(Image.fromarray(planes[i]))) for i in range(C)])
        same = same and np.array_equal(ref, out[b].cpu().numpy())
    cpu_ms = (time.perf_counter() - t0) * 1e3 / max(min(a.cpu_samples, B), 1)
    print(json.dumps({
        "what": "per-channel cloth augmentation (datasets/data_utils.py:346-361), %d x %d x %dx%d" % (B, C, S, S),
        "device_ms_per_batch": ms, "device_images_per_s": B / (ms * 1e-3), "passes_executed": passes,
        "algorithmic_bytes": alg_bytes, "achieved_gb_s": alg_bytes / (ms * 1e-3) / 1e9,
        "op_table_bytes": table.nbytes, "label_map_bytes": lab.numel(), "host_draw_ms_per_sample": draw_ms,
        "cpu_reference_ms_per_sample": cpu_ms, "cpu_reference_images_per_s_one_core": 1e3 / cpu_ms,
        "cpu_reference": "Pillow/torchvision per_channel_transform incl. the one-hot planes, 1 thread, %d sample(s)" % min(a.cpu_samples, B),
        "device_equals_cpu_reference_bit_exact": bool(same)}))


if __name__ == "__main__":
    main()
