"""Device-side input pipeline of the warp stage: the cloth label map travels to the GPU as uint8 and the
per-channel augmentation of the reference's dataset runs there.

Reference (what a `WarpDataset.__getitem__` does on the host, per sample, for the INPUT cloth):
  datasets/data_utils.py:298-343   `decompress_cloth_segment`: scipy CSC .npz -> one-hot fp32 [19, H, W]
  datasets/warp_dataset.py:133-134 `per_channel_transform(cloth, self.cloth_transform)`
  datasets/data_utils.py:346-361   19 x (PIL mode-"F" image -> `RandomOrder([vflip, hflip, RandomAffine, RandomPerspective])`)
  datasets/__init__.py:88-110      `get_transforms(opt)` builds that torchvision transform

Split here:
  host   `load_label_map`   : the .npz as a uint8 label map [H, W] (0.26 MB at 512² instead of 19.9 MB of one-hot fp32;
                              the CSC arrays themselves — int64 data + int32 indices, 12 B per non-zero — are larger than
                              the dense uint8 map as soon as 8 % of the pixels are labelled, so the map is the wire format)
  host   `draw_channel_ops` : the random draws of every channel, made by torchvision's OWN `get_params` / matrix helpers
                              in the order `per_channel_transform` makes them, so the python `random` and torch RNG
                              streams advance exactly as in the reference (tests compare the generator states)
  device `ClothAugmenter.apply` / `ops.augment_channels` -> `sn_augment_channels` (csrc/augment.cu): one-hot expansion +
                              the Pillow resampling of every op, bit-exact (tests: vs Pillow/torchvision themselves).
The result is the fp32 [B, 19, H, W] device tensor `set_input` accepts as `input_cloths`; `target_cloths` goes in as the
uint8 label map itself (`ops.SegMap`).

Not covered (raises): transforms other than the four above, RandomAffine with an interpolation other than NEAREST or a
non-zero fill, RandomPerspective with an interpolation other than BILINEAR (torchvision 0.26's default; 0.4 defaulted to
BICUBIC), affine matrices that Pillow would route to its pure-scale or floating-point code path (rotation and shear both
exactly zero — probability zero under continuous draws).  The dataset's nearest resize + crop that FOLLOW the augmentation
(`--load_size` != stored size, `--crop_size` < `--load_size`) are index gathers: `resize_crop_indices` / `gather_rows_cols`
(the index rule is torch's own CPU op applied to an index ramp), applied on the device by `WarpModel.set_input`.
"""
from __future__ import annotations

import math
import random
from typing import List, Sequence, Tuple

import numpy as np
import torch

from ._lib import AUG_AFFINE_NEAREST, AUG_HFLIP, AUG_MAX_OPS, AUG_PERSPECTIVE_BILINEAR, AUG_VFLIP
from .ops import SegMap  # noqa: F401  (re-export: the compact form `target_cloths` travels in)

Op = Tuple[int, Sequence[float]]
# mirrors `sn_aug_op` (include/swapnet_b200.h): int kind, int nops, double p[8] -> 72 bytes
OP_DTYPE = np.dtype([("kind", "<i4"), ("nops", "<i4"), ("p", "<f8", (8,))])
assert OP_DTYPE.itemsize == 72


def load_label_map(fname: str, n_labels: int = 19) -> np.ndarray:
    """The cloth .npz of the reference's dataset (scipy CSC of the argmax labels, data_utils.py:311-327) as a uint8
    label map [H, W].  Same content as `decompress_cloth_segment(fname, n_labels)` (data_utils.py:298-343): channel L of
    the one-hot tensor is `label == L` for L > 0, channel 0 is empty (zeros are not stored)."""
    from scipy.sparse import load_npz

    m = load_npz(fname).toarray()
    if m.min() < 0 or m.max() >= n_labels:
        raise ValueError(f"{fname}: labels outside [0, {n_labels})")
    return np.ascontiguousarray(m.astype(np.uint8))


def _fix16(v: float) -> int:
    """Pillow's FIX(): 16.16 fixed point, floor(v * 65536 + 0.5) (libImaging/Geometry.c)."""
    v = v * 65536.0 + 0.5
    return math.floor(v) if v < 0.0 else int(v)


def _affine_op(matrix: Sequence[float], width: int, height: int) -> Op:
    a = [float(v) for v in matrix]
    in_range = lambda x, y: abs(x * a[0] + y * a[1] + a[2]) < 32768.0 and abs(x * a[3] + y * a[4] + a[5]) < 32768.0
    if (a[1] == 0 and a[3] == 0) or not (in_range(0, 0) and in_range(width, height) and in_range(0, height)
                                         and in_range(width, 0)):
        raise NotImplementedError("affine matrix without rotation/shear or outside the 16.16 range: Pillow takes a "
                                  "different code path (ImagingScaleAffine / floating point) that is not restated")
    return (AUG_AFFINE_NEAREST, (_fix16(a[0]), _fix16(a[1]), _fix16(a[2] + a[0] * 0.5 + a[1] * 0.5),
                                 _fix16(a[3]), _fix16(a[4]), _fix16(a[5] + a[3] * 0.5 + a[4] * 0.5)))


def _draw_one(t, width: int, height: int):
    """The draws of ONE torchvision transform applied to one PIL image of the given size -> Op or None."""
    from torchvision import transforms as T
    from torchvision.transforms import functional as TF

    nearest = ("nearest", 0)        # InterpolationMode.NEAREST / PIL.Image.NEAREST
    bilinear = ("bilinear", 2)
    mode = lambda v: getattr(v, "value", v)
    if isinstance(t, T.RandomVerticalFlip):
        return (AUG_VFLIP, ()) if torch.rand(1) < t.p else None
    if isinstance(t, T.RandomHorizontalFlip):
        return (AUG_HFLIP, ()) if torch.rand(1) < t.p else None
    if isinstance(t, T.RandomAffine):
        if mode(getattr(t, "interpolation", getattr(t, "resample", 0))) not in nearest or getattr(t, "fill", 0) not in (0, None):
            raise NotImplementedError("RandomAffine: only NEAREST with fill 0 (the reference's configuration)")
        if getattr(t, "center", None) is not None:
            raise NotImplementedError("RandomAffine(center=...)")
        angle, translate, scale, shear = t.get_params(t.degrees, t.translate, t.scale, t.shear, [width, height])
        shear = [float(s) for s in shear] if isinstance(shear, (tuple, list)) else [float(shear), 0.0]
        matrix = TF._get_inverse_affine_matrix([width * 0.5, height * 0.5], angle, list(translate), scale, shear)
        return _affine_op(matrix, width, height)
    if isinstance(t, T.RandomPerspective):
        if mode(t.interpolation) not in bilinear or getattr(t, "fill", 0) not in (0, None):
            raise NotImplementedError("RandomPerspective: only BILINEAR with fill 0")
        if not (torch.rand(1) < t.p):
            return None
        start, end = t.get_params(width, height, t.distortion_scale)
        return (AUG_PERSPECTIVE_BILINEAR, tuple(float(c) for c in TF._get_perspective_coeffs(start, end)))
    raise NotImplementedError(f"per-channel transform {type(t).__name__} is not available on the device")


def _draw_image(transform, width: int, height: int) -> List[Op]:
    from torchvision import transforms as T

    if transform is None:
        return []
    if isinstance(transform, T.RandomOrder):          # transforms.py RandomOrder.__call__: python `random`
        order = list(range(len(transform.transforms)))
        random.shuffle(order)
        seq = [transform.transforms[i] for i in order]
    elif isinstance(transform, T.Compose):
        seq = list(transform.transforms)
    else:
        seq = [transform]
    ops = []
    for t in seq:
        op = _draw_one(t, width, height)
        if op is not None:
            ops.append(op)
    return ops


def draw_channel_ops(transform, channels: int, width: int, height: int) -> List[List[Op]]:
    """The draws `per_channel_transform(cloth[channels, H, W], transform)` (data_utils.py:346-361) makes, channel by
    channel, WITHOUT touching pixels: ops[c] = the geometric ops of channel c in application order."""
    return [_draw_image(transform, width, height) for _ in range(channels)]


def encode_ops(ops_per_plane: Sequence[Sequence[Op]]):
    """[planes][ops] -> (table: np.ndarray of OP_DTYPE [planes, stride], max_ops)."""
    max_ops = max((len(o) for o in ops_per_plane), default=0)
    if max_ops > AUG_MAX_OPS:
        raise NotImplementedError(f"more than {AUG_MAX_OPS} ops on one plane")
    stride = max(max_ops, 1)
    table = np.zeros((len(ops_per_plane), stride), dtype=OP_DTYPE)
    for i, ops in enumerate(ops_per_plane):
        table["nops"][i, :] = len(ops)
        for j, (kind, p) in enumerate(ops):
            table["kind"][i, j] = kind
            table["p"][i, j, :len(p)] = np.asarray(p, dtype=np.float64)
    return table, max_ops


def resize_crop_indices(stored: int, load_size: int, crop=None) -> torch.Tensor:
    """Source index of every output row (or column) of the reference's post-augmentation glue
    (datasets/warp_dataset.py:150-174): `F.interpolate(x, size=load_size)` (nearest) followed by the crop
    `[lo:hi]` (datasets/data_utils.py:186-194).  Both are pure gathers, so they reduce to one index vector per axis.
    The nearest rule is not restated: torch's own CPU op resizes an index ramp."""
    ramp = torch.arange(stored, dtype=torch.float32).view(1, 1, stored, 1)
    idx = torch.nn.functional.interpolate(ramp, size=(int(load_size), 1)).view(-1).long()
    if crop is not None:
        idx = idx[crop[0]:crop[1]]
    return idx.contiguous()


def gather_rows_cols(t: torch.Tensor, iy: torch.Tensor, ix: torch.Tensor) -> torch.Tensor:
    """t[..., iy, :][..., :, ix] — resize (nearest) + crop of [..., H, W] tensors as two index_selects (any device)."""
    return t.index_select(-2, iy).index_select(-1, ix)


def encode_sample(ops_per_channel: Sequence[Sequence[Op]], slots: int) -> torch.Tensor:
    """One sample's draws as a fixed-size uint8 tensor [channels * slots * 72] (a DataLoader's default collate can
    stack it); `OpTable.from_collated` turns the stacked batch back into the device table."""
    table, max_ops = encode_ops(ops_per_channel)
    if max_ops > slots:
        raise NotImplementedError(f"{max_ops} ops on one channel, the sample format has {slots} slots")
    full = np.zeros((len(ops_per_channel), slots), dtype=OP_DTYPE)
    full[:, :table.shape[1]] = table
    full["nops"] = table["nops"][:, :1]
    return torch.from_numpy(full.view(np.uint8).reshape(-1).copy())


class OpTable:
    """The encoded draws of a batch: `sn_aug_op[planes][stride]` as pinned host bytes, ready for one async H2D."""

    def __init__(self, sample_ops: Sequence[Sequence[Sequence[Op]]], pin: bool = True):
        self.batch = len(sample_ops)
        self.channels = len(sample_ops[0]) if self.batch else 0
        assert all(len(s) == self.channels for s in sample_ops)
        table, self.max_ops = encode_ops([o for s in sample_ops for o in s])
        self.stride = table.shape[1]
        self.host = torch.from_numpy(table.view(np.uint8).reshape(-1).copy())
        if pin and torch.cuda.is_available():
            self.host = self.host.pin_memory()

    @classmethod
    def from_collated(cls, t: torch.Tensor, channels: int) -> "OpTable":
        """t: uint8 [B, channels * slots * 72], the stacked `encode_sample` tensors of a batch."""
        assert t.dtype == torch.uint8 and t.dim() == 2 and t.shape[1] % (channels * 72) == 0
        self = cls.__new__(cls)
        self.batch, self.channels, self.stride = t.shape[0], channels, t.shape[1] // (channels * 72)
        self.host = t.contiguous().reshape(-1)
        nops = self.host.numpy().view(OP_DTYPE)["nops"]
        self.max_ops = int(nops.max()) if nops.size else 0
        if self.max_ops > self.stride or (nops.size and int(nops.min()) < 0):
            raise ValueError("corrupt op table")
        return self

    @property
    def nbytes(self) -> int:
        return self.host.numel()


class ClothAugmenter:
    """`per_channel_transform` for a whole batch: draws on the host, pixels on the device.

        aug = ClothAugmenter(get_transforms(opt), channels=19)      # the reference's own transform object
        ops_b = aug.draw(w, h)                                       # per sample, e.g. in Dataset.__getitem__ (host, ~3 ms)
        table = OpTable([ops_0, ..., ops_{B-1}])                     # e.g. in the collate_fn: 72 bytes per (plane, op)
        input_cloths = aug.apply(label_maps_u8_cuda, table)          # fp32 [B, 19, H, W] on the device
        model.set_input({"input_cloths": input_cloths, "target_cloths": SegMap(label_maps_u8_cuda, 19), ...})
    """

    def __init__(self, transform, channels: int = 19):
        self.transform, self.channels = transform, channels
        self._tmp = None

    def draw(self, width: int, height: int) -> List[List[Op]]:
        return draw_channel_ops(self.transform, self.channels, width, height)

    def apply(self, labels: torch.Tensor, table) -> torch.Tensor:
        """labels: uint8 [B, H, W] on the GPU (or a dense fp32 [B, C, H, W] tensor); table: an OpTable, or the list
        [draw(...) of sample b for b in range(B)]."""
        from . import ops

        if not labels.is_cuda:
            raise RuntimeError("ClothAugmenter.apply runs on the GPU only (there is no CPU path)")
        if not isinstance(table, OpTable):
            table = OpTable(table)
        b, (h, w) = labels.shape[0], labels.shape[-2:]
        assert table.batch == b and table.channels == self.channels
        dev = table.host.to(labels.device, non_blocking=True)
        out = torch.empty(b, self.channels, h, w, dtype=torch.float32, device=labels.device)
        tmp = None
        if table.max_ops >= 2:
            if self._tmp is None or self._tmp.shape != out.shape or self._tmp.device != out.device:
                self._tmp = torch.empty_like(out)
            tmp = self._tmp
        ops.augment_channels(labels.contiguous(), self.channels, dev, table.stride, table.max_ops, out, tmp)
        return out
