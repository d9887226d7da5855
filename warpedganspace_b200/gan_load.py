"""Generator wrappers with the reference's uniform ``G(z, shift)`` API (models/gan_load.py:21-188).

Same class names, constructor arguments, attributes (``dim_z``, ``dim_w``, ``shift_in_w_space``,
``target_classes``) and builder functions, over the libwgs_b200 generators.  Outputs are logical NCHW
tensors backed by channels-last (NHWC) memory.
"""
import torch
from torch import nn

from .stylegan2 import Generator as StyleGAN2Generator


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
