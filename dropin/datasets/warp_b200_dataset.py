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
and the DataLoader's default collate stacks them.  With the same seeds the (labels, ops) pair expands to exactly the
`input_cloths` / `target_cloths` tensors the reference dataset yields (tests/test_dropin_launcher.py).

Not supported (raises): stored label maps whose size differs from --load_size, cropping (--crop_size < --load_size or
--crop_bounds): the reference resizes/crops AFTER the augmentation (warp_dataset.py:150-174), which is not on the device.
"""
import random

import torch
from torch import nn

from datasets.warp_dataset import WarpDataset      # the reference's (see datasets/__init__.py of this overlay)
from swapnet_b200 import data as D


class WarpB200Dataset(WarpDataset):
    OP_SLOTS = 4        # ops per channel the table has room for (= the most transforms get_transforms() can list)

    def __init__(self, opt, cloth_dir=None, body_dir=None):
        super().__init__(opt, cloth_dir=cloth_dir, body_dir=body_dir)
        if self.crop_bounds:
            raise NotImplementedError("--dataset warp_b200: cropping after the augmentation is not on the device; use "
                                      "--crop_size == --load_size (or the reference's --dataset warp)")
        n = len(getattr(self.cloth_transform, "transforms", [])) if self.cloth_transform else 0
        if n > self.OP_SLOTS:
            raise NotImplementedError(f"more than {self.OP_SLOTS} transforms per channel")

    def _labels(self, fname):
        lab = D.load_label_map(fname, self.opt.cloth_channels)
        size = self.opt.load_size if isinstance(self.opt.load_size, (tuple, list)) else (self.opt.load_size,) * 2
        if tuple(lab.shape) != tuple(size):
            raise NotImplementedError(f"{fname}: stored size {lab.shape} != --load_size {size}: the resize that follows "
                                      "the augmentation (warp_dataset.py:150-157) is not on the device")
        return lab

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
        return {"body_paths": body_file, "bodys": body, "cloth_paths": cloth_file,
                "input_labels": torch.from_numpy(source), "target_labels": torch.from_numpy(target),
                "input_ops": D.encode_sample(ops, self.OP_SLOTS)}
