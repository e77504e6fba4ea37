"""TEST INFRASTRUCTURE ONLY — numpy restatement of the library's counter-based dropout mask
(swapnet_b200/csrc/elementwise.cu: sn_hash32 / sn_keep / drop_thresh) so that the CPU oracle can
apply exactly the masks the CUDA kernels apply ("shared masks")."""
from __future__ import annotations

import numpy as np
import torch

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def keep_mask(seed: int, p: float, count: int, offset: int = 0) -> np.ndarray:
    """-> bool[count]; element idx (+ offset) is kept iff hash32(seed, idx) >= floor(p * 2^32)."""
    with np.errstate(over="ignore"):
        idx = np.arange(count, dtype=np.uint64) + np.uint64(offset)
        z = idx + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    h = (z >> np.uint64(32)).astype(np.uint32)
    thresh = min(max(p * 4294967296.0, 0.0), 4294967295.0)
    return h >= np.uint32(int(thresh))


def make_drop(seeds: dict, p: float = 0.5, sample_base: int = 0):
    """seeds: stage name -> seed (Engine stage seeds).  Returns drop(name, x NCHW) applying the
    library's mask (indexed over the NHWC element order of the GLOBAL batch: local sample 0 is global sample
    `sample_base`, Engine.sample_base) with the 1/(1-p) scale."""

    def drop(name: str, x: torch.Tensor) -> torch.Tensor:
        n, c, h, w = x.shape
        m = keep_mask(seeds[name], p, n * h * w * c, sample_base * h * w * c).reshape(n, h, w, c)
        mask = torch.from_numpy(m).permute(0, 3, 1, 2).to(x.dtype)
        return x * mask * (1.0 / (1.0 - p))

    return drop
