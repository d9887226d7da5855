"""SupportSets: the RBF warper, backed by one sm_100a kernel per direction.

Drop-in for the reference ``lib.support_sets.SupportSets`` (lib/support_sets.py:6-101): same
constructor signature, attributes, parameter names/shapes (``SUPPORT_SETS [K, 2*D*d]``,
``ALPHAS [K, 2D]``, ``LOGGAMMA [K, 1]``) and ``forward(support_sets_mask, z) -> [B, d]``.

Deviation (documented in DESIGN.md): the mask must be one-hot, which is what every reference call
site builds (lib/trainer.py:227-231, traverse_latent_space.py:391-392); the kernel gathers the
selected row by index instead of multiplying the whole [K, 2*D*d] matrix by the mask.
"""
import math

import torch
from torch import nn

from . import _lib


class _RBFWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support_sets, alphas, loggamma, z, idx, mag, fixed_gamma, learn_gammas):
        B, d = z.shape
        K = support_sets.shape[0]
        n_vec = alphas.shape[1]
        z = z.contiguous()
        out = torch.empty_like(z)
        lg = loggamma.reshape(-1) if learn_gammas else None
        _lib.call('wgs_rbf_warp_forward', _lib.ptr(support_sets), _lib.ptr(alphas), _lib.ptr(lg), fixed_gamma,
                  _lib.ptr(idx), _lib.ptr(z), _lib.ptr(mag), _lib.ptr(out), B, K, n_vec, d, _lib.stream())
        ctx.save_for_backward(support_sets, alphas, loggamma, z, idx, mag if mag is not None else z.new_empty(0))
        ctx.has_mag = mag is not None
        ctx.fixed_gamma = fixed_gamma
        ctx.learn_gammas = learn_gammas
        return out

    @staticmethod
    def backward(ctx, dout):
        support_sets, alphas, loggamma, z, idx, mag = ctx.saved_tensors
        mag = mag if ctx.has_mag else None
        B, d = z.shape
        K = support_sets.shape[0]
        n_vec = alphas.shape[1]
        need_s, need_a, need_g, need_z = ctx.needs_input_grad[:4]
        d_s = torch.zeros_like(support_sets) if need_s else None
        d_a = torch.zeros_like(alphas) if need_a else None
        d_g = torch.zeros_like(loggamma) if (need_g and ctx.learn_gammas) else None
        d_z = torch.empty_like(z) if need_z else None
        lg = loggamma.reshape(-1) if ctx.learn_gammas else None
        dout = dout.contiguous()                    # (a named tensor: _lib.ptr() of a temporary would free it before the launch)
        _lib.call('wgs_rbf_warp_backward', _lib.ptr(support_sets), _lib.ptr(alphas), _lib.ptr(lg),
                  ctx.fixed_gamma, _lib.ptr(idx), _lib.ptr(z), _lib.ptr(mag), _lib.ptr(dout),
                  _lib.ptr(d_s), _lib.ptr(d_g), _lib.ptr(d_a), _lib.ptr(d_z), B, K, n_vec, d, _lib.stream())
        return d_s, d_a, d_g, d_z, None, None, None, None


class SupportSets(nn.Module):
    def __init__(self, num_support_sets, num_support_dipoles, support_vectors_dim,
                 learn_alphas=False, learn_gammas=False, gamma=None):
        super().__init__()
        self.num_support_sets = num_support_sets
        self.num_support_dipoles = num_support_dipoles
        self.support_vectors_dim = support_vectors_dim
        self.learn_alphas = learn_alphas
        self.learn_gammas = learn_gammas
        # the reference requires gamma to be given (train.py:158 passes 1/dim_z); default to that
        self.gamma = gamma if gamma is not None else 1.0 / support_vectors_dim
        self.loggamma = torch.log(torch.scalar_tensor(self.gamma))
        K, D, d = num_support_sets, num_support_dipoles, support_vectors_dim

        # K spheres of radius r_k in [r_min, r_max); D antipodal dipoles on each
        self.r_min, self.r_max = 1.0, 4.0
        self.radii = torch.arange(self.r_min, self.r_max, (self.r_max - self.r_min) / K)
        sv = torch.randn(K, D, d)
        sv = sv / sv.norm(dim=2, keepdim=True) * self.radii[:K].view(K, 1, 1)
        dipoles = torch.stack((sv, -sv), dim=2)                                  # [K, D, 2, d]
        self.SUPPORT_SETS = nn.Parameter(dipoles.reshape(K, 2 * D * d).contiguous(), requires_grad=True)
        signs = torch.tensor([1.0, -1.0]).repeat(D)
        self.ALPHAS = nn.Parameter(signs.expand(K, 2 * D).contiguous(), requires_grad=self.learn_alphas)
        self.LOGGAMMA = nn.Parameter(torch.full((K, 1), math.log(self.gamma)), requires_grad=self.learn_gammas)

    def warp(self, indices, z, magnitudes=None):
        """Fast path: path indices [B] (int64) instead of the one-hot mask; optionally fuses the
        ``target_shift_magnitudes.reshape(-1, 1) *`` of lib/trainer.py:235 into the kernel."""
        if indices.dtype != torch.int64:
            indices = indices.to(torch.int64)
        if magnitudes is not None:
            magnitudes = magnitudes.to(torch.float32).contiguous()
        return _RBFWarp.apply(self.SUPPORT_SETS, self.ALPHAS, self.LOGGAMMA, z, indices.contiguous(), magnitudes,
                              float(self.gamma), bool(self.learn_gammas))

    def forward(self, support_sets_mask, z):
        if not z.is_cuda:
            raise RuntimeError('SupportSets runs on CUDA tensors only (no CPU fallback); got %s' % z.device)
        indices = torch.argmax(support_sets_mask.to(z.device), dim=1)
        return self.warp(indices, z)

    @torch.no_grad()
    def traverse(self, start, paths, eps, shift_steps, shift_leap=1):
        """All traversal chains of traverse_latent_space.py:369-438 in one launch.

        start [C, d] latent (or w) codes, paths [C] int64.  Returns (codes, shifts), each
        [C, 2*(shift_steps // shift_leap)+1, d], most negative step first; with shift_leap > 1 only every
        shift_leap-th step of each direction is kept (:404, :434), the walk itself still takes shift_steps steps."""
        C, d = start.shape
        paths = paths.to(torch.int64).contiguous()
        start = start.contiguous()
        codes = start.new_empty(C, 2 * shift_steps + 1, d)
        shifts = torch.empty_like(codes)
        lg = self.LOGGAMMA.reshape(-1) if self.learn_gammas else None
        _lib.call('wgs_rbf_traverse', _lib.ptr(self.SUPPORT_SETS), _lib.ptr(self.ALPHAS), _lib.ptr(lg),
                  float(self.gamma), _lib.ptr(paths), _lib.ptr(start),
                  float(eps), int(shift_steps), _lib.ptr(codes), _lib.ptr(shifts), C, self.num_support_sets,
                  2 * self.num_support_dipoles, d, _lib.stream())
        if shift_leap > 1:
            n = shift_steps // shift_leap
            keep = [shift_steps - shift_leap * j for j in range(n, 0, -1)] + [shift_steps] + \
                   [shift_steps + shift_leap * j for j in range(1, n + 1)]
            keep = torch.tensor(keep, device=codes.device)
            codes, shifts = codes.index_select(1, keep), shifts.index_select(1, keep)
        return codes, shifts
