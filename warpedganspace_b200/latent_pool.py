"""Latent-code pools: the on-disk format ``sample_gan.py`` writes and ``traverse_latent_space.py`` reads
(sample_gan.py:142-179, traverse_latent_space.py:300-331).

    <pool_dir>/<sha1 of the [1, dim_z] fp32 latent's bytes>/latent_code.pt      torch.Tensor [1, dim_z], CPU
                                                           /image.jpg            G(z), adaptive min-max, JPEG q95

The directory name is ``hashlib.sha1(z.cpu().numpy()).hexdigest()`` (sample_gan.py:159), so a pool can be verified
without any other metadata (some shipped pools prefix an order index, ``NNN_<hash>``); the 58 pools shipped with the reference are the known-answer test
(tests/golden/latent_pool.pt).
"""
import os
import os.path as osp
from hashlib import sha1

import torch


def latent_code_hash(z):
    """sha1 of the raw fp32 bytes of a [1, dim_z] latent code (sample_gan.py:159)."""
    z = z.detach().to('cpu', torch.float32).contiguous()
    if z.dim() != 2 or z.shape[0] != 1:
        raise ValueError('a pool entry is one latent code of shape [1, dim_z]; got %s' % (tuple(z.shape),))
    return sha1(z.numpy()).hexdigest()


def save_latent_pool(zs, out_dir, images_u8=None, jpeg_quality=95):
    """zs [n, dim_z] -> one directory per code; images_u8 (optional) [n, H, W, C] uint8 pixels of G(z) as produced by
    image_out.images_to_uint8(..., adaptive=True).  Returns the list of hashes in order."""
    os.makedirs(out_dir, exist_ok=True)
    hashes = []
    for i in range(zs.shape[0]):
        z = zs[i: i + 1].detach().to('cpu', torch.float32).contiguous()
        h = latent_code_hash(z)
        d = osp.join(out_dir, h)
        os.makedirs(d, exist_ok=True)
        torch.save(z.clone(), osp.join(d, 'latent_code.pt'))
        if images_u8 is not None:
            from .image_out import save_jpeg
            save_jpeg(images_u8[i], osp.join(d, 'image.jpg'), quality=jpeg_quality)
        hashes.append(h)
    return hashes


def load_latent_pool(pool_dir, verify=True):
    """-> (hashes, zs [n, dim_z]) in sorted directory order (traverse_latent_space.py:303-312); with ``verify`` every
    directory name is checked against the hash of the tensor it holds."""
    hashes, zs = [], []
    for name in sorted(os.listdir(pool_dir)):
        path = osp.join(pool_dir, name, 'latent_code.pt')
        if not osp.isfile(path):
            continue
        z = torch.load(path, map_location='cpu')
        if verify and latent_code_hash(z) != name.split('_')[-1]:      # the reference also ships ordered pools: NNN_<hash>
            raise ValueError('latent pool entry %s does not match the sha1 of its latent_code.pt' % name)
        hashes.append(name)
        zs.append(z)
    if not zs:
        raise FileNotFoundError('no latent_code.pt under %s' % pool_dir)
    return hashes, torch.cat(zs)
