"""Latent-space traversal, the generator-only variant of the hot path
(traverse_latent_space.py:333-490 of the reference; GIF strips are out of scope).

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
                   return_images=True, on_frames=None, shift_leap=1):
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
    codes, shifts = S.traverse(chains_start, chains_path, eps, shift_steps, shift_leap)   # [Z*P, F, d]
    F_ = codes.shape[1]                                                       # 2 * (shift_steps // shift_leap) + 1
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


@torch.no_grad()
def traverse_and_save(G, S, pool_dir, out_dir, eps=0.15, shift_steps=16, batch_size=8, paths=None, img_size=None,
                      img_quality=95, shift_in_w_space=None, shift_leap=1):
    """The reference traversal script's output tree (traverse_latent_space.py:333-490) for every code of a latent pool:

        <out_dir>/<hash>/paths_images/path_<dim:03d>/<frame:06d>.jpg      2*shift_steps+1 frames, most negative first
        <out_dir>/<hash>/original_image.jpg                               centre frame of path 0, quality 95
        <out_dir>/<hash>/paths_latent_codes.pt                            [num_paths, 2*shift_steps+1, dim]

    Chains come from one launch of the traversal kernel, frames from the generator's inference path, pixels from the
    device-side tensor2image; only uint8 crosses PCIe."""
    import os
    import os.path as osp
    from .image_out import images_to_uint8, save_jpeg, save_jpegs
    from .latent_pool import load_latent_pool
    hashes, zs = load_latent_pool(pool_dir)
    dev = next(S.parameters()).device
    for h, z in zip(hashes, zs):
        code_dir = osp.join(out_dir, h)
        os.makedirs(osp.join(code_dir, 'paths_images'), exist_ok=True)
        res = traverse_paths(G, S, z[None].to(dev), paths=paths, eps=eps, shift_steps=shift_steps, batch_size=batch_size,
                             shift_in_w_space=shift_in_w_space, return_images=False,
                             on_frames=None, shift_leap=shift_leap)
        codes, shifts = res['codes'][0], res['shifts'][0]                      # [P, F, d]
        P, F_ = codes.shape[0], codes.shape[1]
        wspace = bool(getattr(G, 'shift_in_w_space', False)) if shift_in_w_space is None else shift_in_w_space
        for pi in range(P):
            pdir = osp.join(code_dir, 'paths_images', 'path_%03d' % pi)
            os.makedirs(pdir, exist_ok=True)
            pix = []
            for lo in range(0, F_, batch_size):
                c, s_ = codes[pi, lo: lo + batch_size], shifts[pi, lo: lo + batch_size]
                img = G(c, shift=s_, latent_is_w=True) if wspace else G(c, shift=s_)
                pix.append(images_to_uint8(img, adaptive=True))
            pix = torch.cat(pix)
            save_jpegs(pix, [osp.join(pdir, '%06d.jpg' % t) for t in range(F_)], quality=img_quality, img_size=img_size)
            if pi == 0:
                save_jpeg(pix[F_ // 2].cpu(), osp.join(code_dir, 'original_image.jpg'), quality=95, img_size=img_size)
        torch.save(codes.cpu(), osp.join(code_dir, 'paths_latent_codes.pt'))
    return hashes
