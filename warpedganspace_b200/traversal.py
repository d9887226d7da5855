"""Latent-space traversal, the generator-only variant of the hot path
(traverse_latent_space.py:333-463 of the reference; JPEG/GIF output is out of scope).

For every (latent, path) pair the reference walks ``shift_steps`` sequential RBF steps in each direction in a
Python loop (2 * steps launches of ~12 kernels each, per chain) and then renders ``G(code_t, shift_t)``.  Here all
chains run in ONE launch of the traversal kernel and the frames are rendered in batches on the inference path
(no activations kept).  The reference's look-ahead quirk is reproduced: frame t is rendered from
``code_t + shift_t`` because the wrapper adds the shift again (models/gan_load.py:172,179).
"""
import torch

from . import dist as wdist


@torch.no_grad()
def traverse_paths(G, S, z, paths=None, eps=0.15, shift_steps=16, batch_size=8, shift_in_w_space=None,
                   return_images=True, on_frames=None):
    """z [Z, dim_z]; paths: iterable of path indices (default: all).  Returns dict with
    ``codes`` [Z, K, 2*steps+1, d] (the reference's paths_latent_codes.pt per latent), ``shifts`` (same shape) and,
    when return_images, ``images`` [Z, K, 2*steps+1, C, H, W].  `on_frames(zi, ki, frames)` can consume frames
    instead of keeping them (2.2 M frames at config 5 do not fit in memory)."""
    if shift_in_w_space is None:
        shift_in_w_space = bool(getattr(G, 'shift_in_w_space', False))
    dev = z.device
    K = S.num_support_sets
    paths = torch.arange(K, device=dev) if paths is None else torch.as_tensor(list(paths), device=dev)
    Z, P = z.shape[0], paths.numel()
    start = G.get_w(z) if shift_in_w_space else z                             # traverse_latent_space.py:370
    chains_start = start.repeat_interleave(P, dim=0).contiguous()             # chain c = (latent c // P, path c % P)
    chains_path = paths.repeat(Z).contiguous()
    codes, shifts = S.traverse(chains_start, chains_path, eps, shift_steps)   # [Z*P, F, d]
    F_ = 2 * shift_steps + 1
    out = {'codes': codes.view(Z, P, F_, -1), 'shifts': shifts.view(Z, P, F_, -1)}
    if not (return_images or on_frames):
        return out
    flat_codes, flat_shifts = codes.view(Z * P * F_, -1), shifts.view(Z * P * F_, -1)
    frames = []
    for lo in range(0, flat_codes.shape[0], batch_size):
        hi = min(lo + batch_size, flat_codes.shape[0])
        if shift_in_w_space:
            img = G(flat_codes[lo:hi], shift=flat_shifts[lo:hi], latent_is_w=True)   # :455-458
        else:
            img = G(flat_codes[lo:hi], shift=flat_shifts[lo:hi])                      # :460-462
        if on_frames is not None:
            on_frames(lo, hi, img)
        if return_images:
            frames.append(img.contiguous())
    if return_images:
        imgs = torch.cat(frames)
        out['images'] = imgs.view(Z, P, F_, *imgs.shape[1:])
    return out


def shard_latents(z, rank=None, world=None):
    """Config 5 shards the latent codes over ranks; there is no collective on this path."""
    w, r, _ = wdist.env_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    lo, hi = wdist.shard_range(z.shape[0], rank, world)
    return z[lo:hi]
