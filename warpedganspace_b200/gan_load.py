"""Generator wrappers with the reference's uniform ``G(z, shift)`` API (models/gan_load.py:21-188).

Same class names, constructor arguments, attributes (``dim_z``, ``dim_w``, ``shift_in_w_space``,
``target_classes``) and builder functions, over the libwgs_b200 generators.  Outputs are logical NCHW
tensors backed by channels-last (NHWC) memory.
"""
import torch
from torch import nn

import json
import os

import numpy as np

from .stylegan2 import Generator as StyleGAN2Generator
from .generators import ProgGANGenerator, SNGANGenerator, BigGANGenerator


class StyleGAN2Wrapper(nn.Module):
    """models/gan_load.py:137-179."""

    def __init__(self, G, shift_in_w_space):
        super().__init__()
        self.G = G
        self.shift_in_w_space = shift_in_w_space
        self.dim_z = 512
        self.dim_w = self.G.style_dim if self.shift_in_w_space else self.dim_z

    def get_w(self, z):
        return self.G.get_latent(z)

    def forward(self, z, shift=None, latent_is_w=False):
        if self.shift_in_w_space:
            w = z if latent_is_w else self.G.get_latent(z)
            return self.G([w if shift is None else w + shift], input_is_latent=True)[0]
        return self.G([z if shift is None else z + shift], input_is_latent=False)[0]

    def forward_pair(self, z, shift):
        """(G(z), G(z, shift)) in one batched pass (fast path of lib/trainer.py:200,239)."""
        if self.shift_in_w_space:
            w = self.G.get_latent(z)
            return self.G.synthesize_pair(w, w + shift)
        lat = self.G.get_latent(torch.cat([z, z + shift], dim=0))
        b = z.shape[0]
        return self.G.synthesize_pair(lat[:b], lat[b:])


def build_stylegan2(pretrained_gan_weights, resolution, shift_in_w_space=False):
    """models/gan_load.py:182-188: ``torch.load(path)['g_ema']`` with strict=False."""
    G = StyleGAN2Generator(resolution, 512, 8)
    G.load_state_dict(torch.load(pretrained_gan_weights, map_location='cpu')['g_ema'], strict=False)
    return StyleGAN2Wrapper(G, shift_in_w_space=shift_in_w_space)


class SNGANWrapper(nn.Module):
    """models/gan_load.py:21-28."""

    def __init__(self, G):
        super().__init__()
        self.G = G
        self.dim_z = G.distribution.dim

    def forward(self, z, shift=None):
        return self.G.generate(z if shift is None else z + shift)


def build_sngan(pretrained_gan_weights, gan_type):
    """models/gan_load.py:31-57."""
    cfg = {'SNGAN_MNIST': ('sn_resnet32', 32, 1), 'SNGAN_AnimeFaces': ('sn_resnet64', 64, 3)}[gan_type]
    G = SNGANGenerator(cfg[0], img_size=cfg[1], channels=cfg[2], latent_dim=128)
    G.load_state_dict(torch.load(pretrained_gan_weights, map_location='cpu'), strict=False)
    return SNGANWrapper(G)


class BigGANWrapper(nn.Module):
    """models/gan_load.py:65-81 (class ids are drawn on the host exactly as there)."""

    def __init__(self, G, target_classes=(239,)):
        super().__init__()
        self.G = G
        self.target_classes = nn.Parameter(torch.tensor(target_classes, dtype=torch.int64), requires_grad=False)
        self.dim_z = self.G.dim_z

    def mixed_classes(self, batch_size):
        if self.target_classes.numel() == 1:         # one class: nothing to draw (stays on the device, CUDA-graph safe)
            return self.target_classes.reshape(1).repeat(batch_size).cuda()
        return torch.from_numpy(np.random.choice(self.target_classes.cpu(), [batch_size])).cuda()

    def forward(self, z, shift=None):
        classes = self.mixed_classes(z.shape[0]).to(z.device)
        return self.G(z if shift is None else z + shift, self.G.shared(classes))

    def forward_pair(self, z, shift):
        """(G(z), G(z, shift)) in one batched pass (fast path of lib/trainer.py:200,239).  The class ids are drawn exactly as
        the two separate calls would draw them: once for the un-shifted images, then once for the shifted ones."""
        c_plain = self.mixed_classes(z.shape[0]).to(z.device)
        c_shift = self.mixed_classes(z.shape[0]).to(z.device)
        return self.G.synthesize_pair(z, self.G.shared(c_plain), z + shift, self.G.shared(c_shift))


def build_biggan(pretrained_gan_weights, target_classes, config_path=None):
    """models/gan_load.py:84-101 (generator_config.json of the I128 model: G_ch 96, dim_z 120, hier, shared_dim 128)."""
    config = dict(G_ch=96, dim_z=120, resolution=128, G_attn='64', n_classes=1000, shared_dim=128, hier=True,
                  BN_eps=1e-5, SN_eps=1e-6)
    if config_path and os.path.exists(config_path):
        with open(config_path) as f:
            user = json.load(f)
        config.update({k: user[k] for k in config if k in user})
    G = BigGANGenerator(**config)
    G.load_state_dict(torch.load(pretrained_gan_weights, map_location='cpu'), strict=True)
    return BigGANWrapper(G, target_classes)


class ProgGANWrapper(nn.Module):
    """models/gan_load.py:109-120."""

    def __init__(self, G):
        super().__init__()
        self.G = G
        self.dim_z = 512

    def forward(self, z, shift=None):
        x = z if shift is None else z + shift
        return self.G(x.reshape(x.size(0), x.size(1), 1, 1))

    def forward_pair(self, z, shift):
        """(G(z), G(z, shift)) in one batched pass (fast path of lib/trainer.py:200,239)."""
        return self.G.synthesize_pair(z, z + shift)


def build_proggan(pretrained_gan_weights):
    """models/gan_load.py:123-129."""
    G = ProgGANGenerator()
    G.load_state_dict(torch.load(pretrained_gan_weights, map_location='cpu'))
    return ProgGANWrapper(G)
