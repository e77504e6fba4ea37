"""`--dataset warp_b200`: the reference's WarpDataset with the cloth work moved off the host.

Subclass of the reference's own `datasets.warp_dataset.WarpDataset` (file discovery, body loading, normalisation and
options are inherited, not restated).  What changes is `__getitem__` (datasets/warp_dataset.py:87-183):
  reference                                             here
  decompress .npz -> one-hot fp32 [19,H,W] (19.9 MB)    `data.load_label_map` -> uint8 label map [H,W] (0.26 MB)
  per_channel_transform: 19 x PIL RandomOrder(...)      `data.draw_channel_ops`: the same random draws, no pixels
  (data_utils.py:346-361, ~75-120 ms on a core)         (~2-3 ms); the op table travels with the sample
The pixel work (one-hot expansion + Pillow's resampling, bit-exact) runs on the GPU inside `WarpModel.set_input`, which
recognises the keys below.  A sample is
  bodys [3,H,W] f32, input_labels / target_labels uint8 [H,W], input_ops uint8 [19 * OP_SLOTS * 72], cloth_paths, body_paths
(+ resize_iy / resize_ix int64 index vectors when the stored size differs from --load_size or a crop is configured)
and the DataLoader's default collate stacks them.  With the same seeds the (labels, ops) pair expands to exactly the
`input_cloths` / `target_cloths` tensors the reference dataset yields (tests/test_dropin_launcher.py).

The reference resizes (nearest) and crops AFTER the augmentation (warp_dataset.py:150-174).  Both are gathers: the
augmentation runs at the stored size and `set_input` applies one index_select per axis with the vectors computed here by
`data.resize_crop_indices` (torch's own nearest rule on an index ramp).  The body image takes the reference's own host
path (bilinear resize + crop_tensors).
"""
import random

import torch
from torch import nn

from datasets.data_utils import crop_tensors       # the reference's (see datasets/__init__.py of this overlay)
from datasets.warp_dataset import WarpDataset
from swapnet_b200 import data as D


class WarpB200Dataset(WarpDataset):
    OP_SLOTS = 4        # ops per channel the table has room for (= the most transforms get_transforms() can list)

    def __init__(self, opt, cloth_dir=None, body_dir=None):
        super().__init__(opt, cloth_dir=cloth_dir, body_dir=body_dir)
        if not isinstance(self.opt.load_size, int):
            raise NotImplementedError("--dataset warp_b200: square --load_size only (the plugin's engines are square)")
        self._gather = {}
        n = len(getattr(self.cloth_transform, "transforms", [])) if self.cloth_transform else 0
        if n > self.OP_SLOTS:
            raise NotImplementedError(f"more than {self.OP_SLOTS} transforms per channel")

    def _labels(self, fname):
        return D.load_label_map(fname, self.opt.cloth_channels)

    def _indices(self, h, w):
        """(iy, ix) of the nearest resize to --load_size + the crop, or None when both are the identity."""
        if (h, w) not in self._gather:
            (x0, y0), (x1, y1) = self.crop_bounds if self.crop_bounds else ((None, None), (None, None))
            iy = D.resize_crop_indices(h, self.opt.load_size, (y0, y1) if self.crop_bounds else None)
            ix = D.resize_crop_indices(w, self.opt.load_size, (x0, x1) if self.crop_bounds else None)
            same = len(iy) == h and len(ix) == w and bool((iy == torch.arange(h)).all()) and bool((ix == torch.arange(w)).all())
            self._gather[(h, w)] = None if same else (iy, ix)
        return self._gather[(h, w)]

    def __getitem__(self, index):
        cloth_file = self.cloth_files[index]
        target = self._labels(cloth_file)
        source, ops = target, [[] for _ in range(self.opt.cloth_channels)]
        if self.is_train:
            if self.opt.dataset_mode == "video":      # warp_dataset.py:101-106: a random other frame is the input
                cloth_file = self.cloth_files[random.randint(0, len(self)) - 1]
                source = self._labels(cloth_file)
            elif self.opt.dataset_mode != "image":
                raise ValueError(self.opt.dataset_mode)
            if self.cloth_transform:
                if not self.opt.per_channel_transform:
                    raise NotImplementedError("Sorry, per_channel_transform must be true")
                h, w = source.shape
                ops = D.draw_channel_ops(self.cloth_transform, self.opt.cloth_channels, w, h)
        body_file, body = self._load_body(index)                        # the reference's own loader
        body = nn.functional.interpolate(body.unsqueeze(0), size=self.opt.load_size, mode="bilinear").squeeze()
        if self.crop_bounds:
            body = crop_tensors(body, crop_bounds=self.crop_bounds)
        if source.shape != target.shape:
            raise NotImplementedError("video mode with frames of different stored sizes")
        item = {"body_paths": body_file, "bodys": body, "cloth_paths": cloth_file,
                "input_labels": torch.from_numpy(source), "target_labels": torch.from_numpy(target),
                "input_ops": D.encode_sample(ops, self.OP_SLOTS)}
        idx = self._indices(*source.shape)
        if idx is not None:
            item["resize_iy"], item["resize_ix"] = idx
        return item
