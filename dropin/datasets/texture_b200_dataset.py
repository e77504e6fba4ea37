"""`--dataset texture_b200`: the reference's TextureDataset with the cloth segmentation kept as a uint8 label map
(0.26 MB per 512x512 sample on the wire instead of 19.9 MB of one-hot fp32; `TextureModel.set_input`
takes `cloths` as uint8 [B,H,W] and the kernels expand it on the device, ops.SegMap).

Subclass of the reference's own `datasets.texture_dataset.TextureDataset`; `__getitem__` IS the reference's
(texture loading, ROI scaling, the joint random flips, resize, crop — datasets/texture_dataset.py:87-165), run with one
substitution: while it executes, the module's `decompress_cloth_segment` returns the label plane itself (twice, so that
the reference's `.squeeze()` keeps a channel axis) instead of the 19-channel one-hot expansion.  The nearest-neighbour
resize and the crop that follow act on every channel alike, so resizing the label plane and expanding afterwards is the
same tensor as expanding and resizing (tests/test_dropin_launcher.py compares with the reference dataset).
"""
import torch

import datasets.texture_dataset as _ref                # the reference's module (see datasets/__init__.py of this overlay)
from swapnet_b200 import data as D


def _label_planes(fname, n_labels):
    lab = torch.from_numpy(D.load_label_map(fname, n_labels)).float()
    return lab.unsqueeze(0).expand(2, -1, -1).contiguous()


class TextureB200Dataset(_ref.TextureDataset):
    def __getitem__(self, index):
        original = _ref.decompress_cloth_segment
        _ref.decompress_cloth_segment = _label_planes
        try:
            item = super().__getitem__(index)
        finally:
            _ref.decompress_cloth_segment = original
        item["cloths"] = item["cloths"][0].to(torch.uint8)          # [H, W]
        return item
