"""GPU parity of the generator-only traversal path (config 5 shape, reduced) against the oracle chains + oracle G."""
import pytest
import torch

import oracle.support_sets as o_ss
import oracle.stylegan2 as o_sg2
import oracle.step as o_step

pytestmark = pytest.mark.gpu


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize('wspace', [False, True])
def test_traverse_paths_matches_oracle(wspace):
    from warpedganspace_b200 import SupportSets
    from warpedganspace_b200.stylegan2 import Generator
    from warpedganspace_b200.gan_load import StyleGAN2Wrapper
    from warpedganspace_b200.traversal import traverse_paths
    ch = {4: 64, 8: 64, 16: 32, 32: 32}
    size, K, D, steps, eps = 32, 8, 4, 3, 0.2
    g_sd = o_sg2.init_state(size=size, generator=gen(1), channels=ch)
    s_sd = o_ss.init_state(K, D, 512, generator=gen(2))
    G = Generator(size, 512, 8, channels=ch)
    G.load_state_dict(g_sd, strict=False)
    W = StyleGAN2Wrapper(G, shift_in_w_space=wspace).cuda().eval()
    S = SupportSets(K, D, 512, learn_gammas=True, gamma=1.0 / 512)
    S.load_state_dict(s_sd)
    S.cuda()
    z = torch.randn(2, 512, generator=gen(3))
    paths = [1, 6]
    out = traverse_paths(W, S, z.cuda(), paths=paths, eps=eps, shift_steps=steps, batch_size=5)
    assert tuple(out['images'].shape) == (2, 2, 2 * steps + 1, 3, size, size)
    for zi in range(2):
        start = o_sg2.mapping(g_sd, z[zi:zi + 1]) if wspace else z[zi:zi + 1]
        for pi, k in enumerate(paths):
            codes, shifts = o_step.traverse_chain(s_sd, start, k, eps, steps)
            assert rel(out['codes'][zi, pi], codes) < 1e-5
            with torch.no_grad():
                want = o_sg2.generate(g_sd, codes, shifts, size, shift_in_w_space=wspace, latent_is_w=wspace)
            assert rel(out['images'][zi, pi], want) < 2e-4
    # centre frame = the un-shifted latent
    with torch.no_grad():
        plain = W(z.cuda())
    assert rel(out['images'][:, 0, steps], plain) < 1e-6


@pytest.mark.parametrize('adaptive', [True, False])
@pytest.mark.parametrize('shape', [(3, 3, 33, 47), (1, 1, 32, 32), (2, 3, 256, 256)])
def test_images_to_uint8_is_bit_identical_to_tensor2image(shape, adaptive):
    """Device-side tensor2image (traverse_latent_space.py:26-41) vs the reference's host formula on the same fp32 image:
    same operations in the same order -> identical uint8 pixels (HWC)."""
    from warpedganspace_b200.image_out import images_to_uint8
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 1.3).contiguous(memory_format=torch.channels_last)
    want = []
    for t in x:                                                            # the reference's per-image host code
        t = (t - t.min()) / (t.max() - t.min()) if adaptive else (t + 1) / 2
        want.append((255 * t).to(torch.uint8).permute(1, 2, 0))
    got = images_to_uint8(x.cuda(), adaptive=adaptive)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (shape[0], shape[2], shape[3], shape[1])
    assert torch.equal(got.cpu(), torch.stack(want))


def test_traverse_and_save_writes_the_reference_tree(tmp_path):
    """traverse_latent_space.py:333-490: <hash>/paths_images/path_XXX/NNNNNN.jpg, original_image.jpg,
    paths_latent_codes.pt [paths, 2*steps+1, dim]; frames decode to the generator's images."""
    import os
    from PIL import Image
    from warpedganspace_b200 import SupportSets, latent_pool
    from warpedganspace_b200.generators import SNGANGenerator
    from warpedganspace_b200.gan_load import SNGANWrapper
    from warpedganspace_b200.traversal import traverse_and_save
    import oracle.sngan as o_sn
    import oracle.support_sets as o_ss
    gen = lambda s: torch.Generator().manual_seed(s)
    G = SNGANGenerator('sn_resnet32', 32, 1)
    G.load_state_dict({'model.' + k: v for k, v in o_sn.init_state('sn_resnet32', 1, generator=gen(1)).items()}, strict=False)
    W = SNGANWrapper(G).cuda().eval()
    S = SupportSets(4, 2, 128, learn_gammas=True, gamma=1.0 / 128)
    S.load_state_dict(o_ss.init_state(4, 2, 128, generator=gen(2)))
    S = S.cuda()
    zs = torch.randn(2, 128, generator=gen(3))
    hashes = latent_pool.save_latent_pool(zs, str(tmp_path / 'pool'))
    done = traverse_and_save(W, S, str(tmp_path / 'pool'), str(tmp_path / 'out'), eps=0.2, shift_steps=3, batch_size=4)
    assert sorted(done) == sorted(hashes)
    for h in hashes:
        root = tmp_path / 'out' / h
        assert sorted(os.listdir(root)) == ['original_image.jpg', 'paths_images', 'paths_latent_codes.pt']
        assert sorted(os.listdir(root / 'paths_images')) == ['path_%03d' % i for i in range(4)]
        assert sorted(os.listdir(root / 'paths_images' / 'path_002')) == ['%06d.jpg' % t for t in range(7)]
        codes = torch.load(root / 'paths_latent_codes.pt')
        assert tuple(codes.shape) == (4, 7, 128)
        z = zs[hashes.index(h)]
        assert torch.allclose(codes[:, 3], z.expand(4, -1), atol=1e-6)            # centre frame = the pool's code
        im = Image.open(root / 'paths_images' / 'path_000' / '000003.jpg')
        assert im.size == (32, 32) and im.mode == 'L'


@pytest.mark.parametrize('shape', [(3, 256, 320, 3), (2, 64, 64, 1), (1, 1024, 1024, 3)])
def test_nvjpeg_encode_decodes_to_the_frames(shape, tmp_path):
    """Output stage on the GPU (traverse_latent_space.py:466-483 writes PIL JPEGs: quality 95, optimised, progressive): the
    nvJPEG bitstreams must be valid JPEG files of the right size / mode, progressive, and decode to the source pixels as
    closely as the reference's own PIL encode of the same pixels does."""
    import io
    import numpy as np
    from PIL import Image
    from warpedganspace_b200.image_out import encode_jpegs, save_jpegs, nvjpeg_available
    if not nvjpeg_available():
        pytest.skip('nvJPEG is not installed on this box')
    n, h, w, c = shape
    g = torch.Generator().manual_seed(h + w)
    # smooth synthetic frames (JPEG is built for those): low-frequency pattern + mild noise
    yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h), torch.linspace(0, 6.28, w), indexing='ij')
    base = torch.stack([torch.sin(yy * (i + 1)) * torch.cos(xx * (c - i)) for i in range(c)], -1)
    pix = ((base[None] * 0.4 + 0.5 + 0.02 * torch.randn(n, h, w, c, generator=g)).clamp(0, 1) * 255).to(torch.uint8)
    streams = encode_jpegs(pix.cuda(), quality=95, progressive=True)
    assert len(streams) == n
    for i, data in enumerate(streams):
        assert data[:2] == b'\xff\xd8' and data[-2:] == b'\xff\xd9'                 # SOI ... EOI
        assert b'\xff\xc2' in data                                                      # SOF2: progressive DCT
        im = Image.open(io.BytesIO(data))
        assert im.size == (w, h) and im.mode == ('L' if c == 1 else 'RGB')
        dec = np.asarray(im).reshape(h, w, c).astype(np.float64)
        src = pix[i].numpy().astype(np.float64)
        err_nv = np.sqrt(((dec - src) ** 2).mean())
        ref = io.BytesIO()
        Image.fromarray(pix[i].numpy()[:, :, 0] if c == 1 else pix[i].numpy()).save(ref, 'JPEG', quality=95, optimize=True,
                                                                                    progressive=True)
        dec_ref = np.asarray(Image.open(io.BytesIO(ref.getvalue()))).reshape(h, w, c).astype(np.float64)
        err_pil = np.sqrt(((dec_ref - src) ** 2).mean())
        print('image %d: rmse nvJPEG %.3f, PIL %.3f (of 255); bytes %d vs %d' % (i, err_nv, err_pil, len(data), len(ref.getvalue())))
        assert err_nv < max(1.5 * err_pil, 1.0)
    paths = [str(tmp_path / ('%d.jpg' % i)) for i in range(n)]
    save_jpegs(pix.cuda(), paths, quality=95)
    assert all(Image.open(p).size == (w, h) for p in paths)
