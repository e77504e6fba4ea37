"""Texture-stage plugin (`--model texture`) on the B200 engines.

Options, attributes and step semantics of /root/reference/models/texture_model.py:17-180 for the
`--netG swapnet` generator (TextureModule): conditional PatchGAN on cat(cloths, texture) (cloth
FIRST, :138,142,164), loss_G = lambda_gan * GAN + lambda_l1 * L1 (+ perceptual terms).

The VGG16 perceptual terms (texture_model.py:171-178; weights `vgg16(pretrained=True)` are not
obtainable offline) are not on the CUDA path yet: lambda_content / lambda_style must be 0
— anything else raises instead of silently falling back.
"""
from __future__ import annotations

from argparse import ArgumentParser

import torch

from .. import engine as E
from .. import modules as M
from .. import ops
from ..ops import GradSrc
from .base_gan import BaseGAN
from .base_model import LazyLoss


class TextureModel(BaseGAN):
    @staticmethod
    def modify_commandline_options(parser: ArgumentParser, is_train):
        parser = super(TextureModel, TextureModel).modify_commandline_options(parser, is_train)
        parser.set_defaults(input_transforms=["hflip", "vflip", "affine", "perspective"])
        parser.add_argument("--netG", default="swapnet", choices=["swapnet", "unet_128"])
        if is_train:
            parser.add_argument("--lambda_l1", type=float, default=10, help="weight for L1 loss in final term")
            parser.add_argument("--lambda_content", type=float, default=20, help="weight for content loss in final term")
            parser.add_argument("--lambda_style", type=float, default=1e-8, help="weight for style loss in final term")
            parser.set_defaults(display_ncols=5)
        return parser

    def __init__(self, opt):
        if getattr(opt, "netG", "swapnet") != "swapnet":
            raise NotImplementedError("--netG unet_128 (the reference's debugging generator) is not provided")
        BaseGAN.__init__(self, opt)
        self.visual_names = ["textures_unnormalized", "cloths_decoded", "fakes", "fakes_scaled"]
        if self.is_train:
            self.visual_names.append("targets_unnormalized")
            if float(getattr(opt, "lambda_content", 0)) != 0 or float(getattr(opt, "lambda_style", 0)) != 0:
                raise NotImplementedError(
                    "VGG16 perceptual loss (lambda_content / lambda_style != 0) is not on the B200 path yet; "
                    "run with --lambda_content 0 --lambda_style 0")
            lam = float(opt.lambda_gan)
            self.loss_G_l1 = LazyLoss(lambda: self._acc[3].item())
            self.loss_G_content = 0.0
            self.loss_G_style = 0.0
            self.loss_G = LazyLoss(lambda: lam * self._acc[2].item() + self._acc[3].item())
            for loss in ("l1", "content", "style"):
                if getattr(opt, "lambda_" + loss, 0) != 0:
                    self.loss_names.append("G_" + loss)

    def compute_visuals(self):
        from datasets.data_utils import scale_tensor, unnormalize
        from util.decode_labels import decode_cloth_labels

        self.textures_unnormalized = unnormalize(self.textures, *self.opt.texture_norm_stats)
        self.cloths_decoded = decode_cloth_labels(self.cloths)
        self.fakes_scaled = scale_tensor(self.fakes, scale_each=True)
        if self.is_train:
            self.targets_unnormalized = unnormalize(self.targets, *self.opt.texture_norm_stats)

    def get_D_inchannels(self):
        return self.opt.texture_channels + self.opt.cloth_channels

    def define_G(self):
        return M.TextureModule(texture_channels=self.opt.texture_channels, cloth_channels=self.opt.cloth_channels,
                               num_roi=self.opt.body_channels, img_size=self.opt.crop_size,
                               norm_type=getattr(self.opt, "norm", "instance"))

    def build_generator_engine(self, batch, size):
        return E.TextureEngine(self.net_generator, batch, size, self.device, self.nsplit, train=self.is_train)

    def set_input(self, input):
        f32 = dict(device=self.device, dtype=torch.float32, non_blocking=True)
        self.textures = input["input_textures"].to(**f32).contiguous()
        self.rois = input["rois"].to(**f32).contiguous()
        self.cloths = input["cloths"].to(**f32).contiguous()
        self.targets = self.copy_late(input["target_textures"])
        self.image_paths = tuple(zip(input["cloth_paths"], input["texture_paths"]))

    def forward(self):
        B, _, S, S2 = self.textures.shape
        assert S == S2, "square inputs expected"
        self.ensure_engines(B, S)
        g = self._eng_G
        g.pack()
        out = g.forward(self.textures, self.rois, self.cloths, training=self.training and self.is_train,
                        seed=self.step_seed())
        self.fakes = out.permute(0, 3, 1, 2)
        self.wait_late_copies()

    def pack_D_inputs(self, din_fake, din_real):
        ops.pack_concat([(self.cloths, False), (self._eng_G.fakes, True)], din_fake)
        if din_real is not None:
            ops.pack_concat([(self.cloths, False), (self.targets, False)], din_real)

    def backward_G(self):
        g = self._eng_G
        B, S = self._eng_key
        ct = self.opt.texture_channels
        if not hasattr(self, "_dl1") or self._dl1.shape[0] != B or self._dl1.shape[1] != S:
            self._dl1 = torch.zeros(B, S, S, ct, device=self.device)
        ops.l1_loss_fwd_bwd(g.fakes, ct, self.targets, float(self.opt.lambda_l1), self._acc[3:4], self._dl1)
        dx = self.gan_backward_through_D()
        g.backward([GradSrc(self._dl1), GradSrc(dx, self.opt.cloth_channels)])
        self.allreduce_grads(g)
