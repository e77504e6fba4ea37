"""TEST INFRASTRUCTURE ONLY — numpy restatement of the per-channel cloth augmentation of the warp dataset.

What the reference does (the step BEFORE the hot path):
  * datasets/data_utils.py:346-361 `per_channel_transform(input_tensor, transform_function)`: every one of the 19
    channels of the one-hot cloth tensor becomes a PIL mode-"F" image (`Image.fromarray` of a float32 plane) and goes
    through `transform_function` on its own, i.e. with its own random draws;
  * datasets/__init__.py:88-110 `get_transforms(opt)`: `transform_function` is torchvision's
    `RandomOrder([RandomVerticalFlip(), RandomHorizontalFlip(), RandomAffine(degrees=10, translate=(.1,.1),
    scale=(.8,1.2), shear=20), RandomPerspective()])` (the subset named by `--input_transforms`;
    warp default: all four, datasets/warp_dataset.py:29-38);
  * datasets/warp_dataset.py:133-134 is the call site.

The pixel arithmetic lives in two third-party packages that are NOT under /root/reference:
  * Pillow (pinned 5.4.1, environment.yml:70; this container: 12.2.0) — `Image.transpose`, `Image.transform(AFFINE,
    NEAREST)`, `Image.transform(PERSPECTIVE, BILINEAR)` on mode "F" (libImaging/Geometry.c).  Restated here:
      - AFFINE + NEAREST with a rotation/shear term runs `affine_fixed`: 16.16 fixed point, coefficients
        FIX(v) = floor(v*65536 + 0.5), source pixel (xx >> 16, yy >> 16) with xx = FIX(a2 + a0/2 + a1/2) + x*a0 + y*a1;
        pixels whose source falls outside keep the fill colour (0);
      - PERSPECTIVE + BILINEAR: source coordinates in doubles at the pixel centre (x+.5, y+.5), rejected outside
        [0, size), then the 2x2 taps around (xs-.5, ys-.5) with clamped indices; the horizontal interpolation takes the
        tap DIFFERENCE in float32 (both taps are FLOAT32 in C, so `b - a` is a float operation) and everything else in
        doubles; the second row is dropped (v2 = v1) when y+1 is outside; the result is rounded to float32;
      - the flips are plain index reversals.
  * torchvision (pinned 0.4.0, environment.yml:94; here 0.26.0) — the random draws (`get_params`) and the matrices
    (`_get_inverse_affine_matrix`, `_get_perspective_coeffs`).  Those are NOT restated: the tests call torchvision's own
    functions for them, and `swapnet_b200.data.draw_channel_ops` (the product's host side) does the same.

Pinned (tests/test_augment_cpu.py): bit-identical to Pillow 12.2 / torchvision 0.26 in this container on seeded cases
(binary and real-valued planes, square and odd sizes), to the reference's own `per_channel_transform(get_transforms(opt))`
when /root/reference is importable, and to the golden fixture tests/golden/augment_64.npz made from the reference by
tests/tools/make_golden_augment.py.  torchvision 0.4's RandomPerspective defaulted to BICUBIC, 0.26's to BILINEAR: parity
is defined against the container's versions, the op table carries the interpolation explicitly.
"""
from __future__ import annotations

import math

import numpy as np

# op kinds (the same numbers as include/swapnet_b200.h SN_AUG_*)
NONE, HFLIP, VFLIP, AFFINE_NEAREST, PERSPECTIVE_BILINEAR = 0, 1, 2, 3, 4


def fix16(v: float) -> int:
    """Geometry.c FIX(): floor(v * 65536 + 0.5) with C's truncation for non-negative values."""
    v = v * 65536.0 + 0.5
    return math.floor(v) if v < 0.0 else int(v)


def affine_fixed_coeffs(a) -> tuple:
    """(a0, a1, a2, a3, a4, a5) of Geometry.c affine_fixed for the 6 floats of Image.transform(AFFINE)."""
    return (fix16(a[0]), fix16(a[1]), fix16(a[2] + a[0] * 0.5 + a[1] * 0.5),
            fix16(a[3]), fix16(a[4]), fix16(a[5] + a[3] * 0.5 + a[4] * 0.5))


def affine_takes_fixed_path(a, w: int, h: int) -> bool:
    """Pillow picks affine_fixed for NEAREST unless the matrix is a pure scale (a1 == a3 == 0: ImagingScaleAffine) or
    a corner leaves the 16.16 range (check_fixed)."""
    if a[1] == 0 and a[3] == 0:
        return False
    ok = lambda x, y: abs(x * a[0] + y * a[1] + a[2]) < 32768.0 and abs(x * a[3] + y * a[4] + a[5]) < 32768.0
    return ok(0, 0) and ok(w, h) and ok(0, h) and ok(w, 0)


def hflip(img: np.ndarray) -> np.ndarray:
    return img[:, ::-1].copy()


def vflip(img: np.ndarray) -> np.ndarray:
    return img[::-1, :].copy()


def affine_nearest(img: np.ndarray, a) -> np.ndarray:
    """Image.transform(size, AFFINE, a, NEAREST, fillcolor=0) on a mode-"F" image (float32 [h, w])."""
    h, w = img.shape
    assert affine_takes_fixed_path(a, w, h), "pure-scale / out-of-range matrices take another Pillow code path"
    return affine_nearest_fixed(img, affine_fixed_coeffs(a))


def affine_nearest_fixed(img: np.ndarray, coeffs) -> np.ndarray:
    """The pixel loop of Geometry.c affine_fixed for already FIX()ed coefficients (the form the op table carries)."""
    h, w = img.shape
    a0, a1, a2, a3, a4, a5 = (int(v) for v in coeffs[:6])
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    xin = (a2 + y * a1 + x * a0) >> 16
    yin = (a5 + y * a4 + x * a3) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def perspective_bilinear(img: np.ndarray, c) -> np.ndarray:
    """Image.transform(size, PERSPECTIVE, c, BILINEAR, fillcolor=0) on a mode-"F" image (float32 [h, w])."""
    h, w = img.shape
    y, x = np.mgrid[0:h, 0:w]
    xin, yin = x + 0.5, y + 0.5
    den = c[6] * xin + c[7] * yin + 1
    xs = (c[0] * xin + c[1] * yin + c[2]) / den
    ys = (c[3] * xin + c[4] * yin + c[5]) / den
    ok = ~((xs < 0.0) | (xs >= w) | (ys < 0.0) | (ys >= h))
    xs, ys = xs - 0.5, ys - 0.5
    x0, y0 = np.floor(xs).astype(np.int64), np.floor(ys).astype(np.int64)
    dx, dy = xs - x0, ys - y0
    xc = lambda v: np.clip(v, 0, w - 1)
    yc = lambda v: np.clip(v, 0, h - 1)

    def row(yy):          # a + (double)(float)(b - a) * dx
        p, q = img[yy, xc(x0)], img[yy, xc(x0 + 1)]
        return p.astype(np.float64) + (q - p).astype(np.float64) * dx

    v1 = row(yc(y0))
    v2 = np.where((y0 + 1 >= 0) & (y0 + 1 < h), row(yc(y0 + 1)), v1)
    v = v1 + (v2 - v1) * dy
    out = np.zeros_like(img)
    out[ok] = v[ok].astype(np.float32)
    return out


def apply_op(img: np.ndarray, kind: int, p) -> np.ndarray:
    if kind == NONE:
        return img
    if kind == HFLIP:
        return hflip(img)
    if kind == VFLIP:
        return vflip(img)
    if kind == AFFINE_NEAREST:          # p = the six 16.16 coefficients
        return affine_nearest_fixed(img, p)
    if kind == PERSPECTIVE_BILINEAR:
        return perspective_bilinear(img, p)
    raise ValueError(kind)


def onehot(labels: np.ndarray, channels: int) -> np.ndarray:
    """uint8 label map [h, w] -> float32 [channels, h, w]; label 0 is the all-zero vector
    (datasets/data_utils.py:330-343: the sparse matrix does not store zeros)."""
    ch = np.arange(channels).reshape(-1, 1, 1)
    return ((labels[None] == ch) & (ch > 0)).astype(np.float32)


def per_channel_transform(planes: np.ndarray, ops) -> np.ndarray:
    """planes float32 [c, h, w]; ops[c] = list of (kind, params) in application order
    (datasets/data_utils.py:346-361 with the draws already made; params as in `sn_aug_op.p`: the six 16.16
    coefficients for AFFINE_NEAREST, the eight Image.transform coefficients for PERSPECTIVE_BILINEAR)."""
    out = np.zeros_like(planes)
    for c in range(planes.shape[0]):
        img = planes[c]
        for kind, p in ops[c]:
            img = apply_op(img, kind, p)
        out[c] = img
    return out
