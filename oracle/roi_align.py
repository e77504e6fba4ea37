"""TEST INFRASTRUCTURE ONLY — numpy restatement of the ROI path of TextureModule.

Follows, operation by operation in float32:
  * TextureModule.reshape_rois            (/root/reference/modules/swapnet_modules.py:209-229)
  * torchvision.ops.RoIAlign(output_size=(128,128), spatial_scale=1, sampling_ratio=1) as called at
    swapnet_modules.py:166-168,234 — third-party (torchvision 0.4.0 pinned in environment.yml:94;
    the op has no `aligned` arg there == today's aligned=False).  Algorithm = torchvision's CPU
    kernel (roi_align_kernel.cpp + pre_calc_for_bilinear_interpolate); pinned in
    tests/test_oracle_cpu.py against torchvision 0.26's CPU op on the notebook ROI fixture
    (test/Test TextureDataset Draw ROIs.ipynb) incl. degenerate / out-of-bounds ROIs: max |diff| = 0.
  * the .view() repack                    (swapnet_modules.py:237-240): channel = 3*roi + rgb
"""
from __future__ import annotations

import numpy as np

# first sample of the notebook fixture (256x256 image): xmin, ymin, xmax, ymax
NOTEBOOK_ROIS_256 = np.array(
    [[159, 0, 193, 14], [144, 15, 206, 89], [255, 0, 255, 0], [196, 20, 215, 94], [144, 151, 180, 229],
     [179, 151, 216, 226], [156, 1, 188, 24], [141, 83, 215, 155], [128, 20, 160, 82], [206, 92, 226, 158],
     [145, 220, 168, 255], [174, 217, 203, 255]], dtype=np.float32)
# degenerate rows seen in other samples of the same notebook
NOTEBOOK_EXTRA_256 = np.array([[0, 255, 0, 255], [147, 255, 183, 255]], dtype=np.float32)


def reshape_rois(rois: np.ndarray) -> np.ndarray:
    """[B, R, 4] -> [B*R, 5], column 0 = batch index (as the float dtype of rois), rows b-major."""
    b, r, _ = rois.shape
    idx = np.repeat(np.arange(b), r).astype(rois.dtype)[:, None]
    return np.concatenate([idx, rois.reshape(-1, 4)], axis=1)


def sample_table(rois5: np.ndarray, height: int, width: int, pool: int):
    """Per (roi, ph, pw): integer tap indices (y_low, x_low, y_high, x_high), weights w1..w4 and the
    `empty` flag, all computed in float32 exactly like the CPU kernel."""
    f = np.float32
    k = rois5.shape[0]
    ylo = np.zeros((k, pool, pool), np.int32); xlo = np.zeros_like(ylo)
    yhi = np.zeros_like(ylo); xhi = np.zeros_like(ylo)
    w = np.zeros((k, pool, pool, 4), np.float32)
    empty = np.zeros((k, pool, pool), bool)
    ph = np.arange(pool, dtype=np.float32)
    for i in range(k):
        x1, y1, x2, y2 = (f(v) for v in rois5[i, 1:5])
        roi_w = max(f(x2 - x1), f(1.0)); roi_h = max(f(y2 - y1), f(1.0))
        bin_h = f(roi_h / f(pool)); bin_w = f(roi_w / f(pool))
        ys = ((y1 + ph * bin_h).astype(np.float32) + f(f(0.5) * bin_h) / f(1.0)).astype(np.float32)
        xs = ((x1 + ph * bin_w).astype(np.float32) + f(f(0.5) * bin_w) / f(1.0)).astype(np.float32)
        Y, X = np.meshgrid(ys, xs, indexing="ij")
        e = (Y < -1.0) | (Y > height) | (X < -1.0) | (X > width)
        Y = np.where(Y <= 0, f(0), Y); X = np.where(X <= 0, f(0), X)
        yl = Y.astype(np.int32); xl = X.astype(np.int32)
        ycl = yl >= height - 1; xcl = xl >= width - 1
        yh = np.where(ycl, height - 1, yl + 1); yl = np.where(ycl, height - 1, yl)
        xh = np.where(xcl, width - 1, xl + 1); xl = np.where(xcl, width - 1, xl)
        Y = np.where(ycl, yl.astype(np.float32), Y); X = np.where(xcl, xl.astype(np.float32), X)
        ly = (Y - yl.astype(np.float32)).astype(np.float32); lx = (X - xl.astype(np.float32)).astype(np.float32)
        hy = (f(1.0) - ly).astype(np.float32); hx = (f(1.0) - lx).astype(np.float32)
        w[i, ..., 0] = hy * hx; w[i, ..., 1] = hy * lx; w[i, ..., 2] = ly * hx; w[i, ..., 3] = ly * lx
        w[i][e] = 0
        ylo[i], xlo[i], yhi[i], xhi[i] = yl, xl, yh, xh
        for a in (ylo, xlo, yhi, xhi):
            a[i][e] = 0
        empty[i] = e
    return ylo, xlo, yhi, xhi, w, empty


def roi_align_pack(tex: np.ndarray, rois: np.ndarray, pool: int = 128) -> np.ndarray:
    """tex [B, C, H, W] float32, rois [B, R, 4] -> [B, C*R, pool, pool] (channel = C*roi + c),
    i.e. roi_align(...).view(B, -1, pool, pool) of the reference."""
    b, c, h, wd = tex.shape
    r = rois.shape[1]
    rois5 = reshape_rois(rois.astype(np.float32))
    ylo, xlo, yhi, xhi, w, empty = sample_table(rois5, h, wd, pool)
    out = np.zeros((b * r, c, pool, pool), np.float32)
    for k in range(b * r):
        bi = int(rois5[k, 0])
        img = tex[bi]
        d1 = img[:, ylo[k], xlo[k]]; d2 = img[:, ylo[k], xhi[k]]
        d3 = img[:, yhi[k], xlo[k]]; d4 = img[:, yhi[k], xhi[k]]
        v = ((w[k, ..., 0] * d1 + w[k, ..., 1] * d2).astype(np.float32) + w[k, ..., 2] * d3).astype(np.float32)
        v = (v + w[k, ..., 3] * d4).astype(np.float32)
        v[:, empty[k]] = 0
        out[k] = v
    return out.reshape(b, r * c, pool, pool)
