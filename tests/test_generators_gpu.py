"""GPU parity: ProgGAN / SNGAN / BigGAN generators (convs on the tcgen05 kernel) against the reference
fixtures and, for the data gradient that training needs, against the oracle."""
import pytest
import torch

import oracle.proggan as o_pg
import oracle.sngan as o_sn
import oracle.biggan as o_bg

pytestmark = pytest.mark.gpu


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def test_proggan_fixture(golden):
    from warpedganspace_b200.generators import ProgGANGenerator
    from warpedganspace_b200.gan_load import ProgGANWrapper
    fx = golden('proggan_1024.pt')
    for tag, like in (('like', True), ('ctor', False)):
        c = fx[tag]
        G = ProgGANGenerator()
        G.load_state_dict(o_pg.init_state(generator=gen(c['seed']), pretrained_like=like))
        W = ProgGANWrapper(G).cuda().eval()
        with torch.no_grad():
            img = W(fx['z'].cuda(), fx['shift'].cuda())
        assert tuple(img.shape) == (1, 3, 1024, 1024)
        assert rel(img[:, :, ::fx['stride'], ::fx['stride']], c['img']) < 2e-4, tag
        assert abs(float(img.double().std()) - c['std']) < 1e-3 * c['std']


@pytest.mark.parametrize('size', [32, 64])
def test_sngan_fixture(golden, size):
    from warpedganspace_b200.generators import SNGANGenerator
    from warpedganspace_b200.gan_load import SNGANWrapper
    fx = golden('sngan_%d.pt' % size)
    sd = o_sn.init_state(fx['model'], fx['channels'], generator=gen(fx['seed']))
    G = SNGANGenerator(fx['model'], size, fx['channels'])
    res = G.load_state_dict({'model.' + k: v for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys and all('num_batches_tracked' in k for k in res.missing_keys)
    W = SNGANWrapper(G).cuda().eval()
    assert W.dim_z == 128
    with torch.no_grad():
        assert rel(W(fx['z'].cuda()), fx['img']) < 1e-4
        assert rel(W(fx['z'].cuda(), fx['shift'].cuda()), fx['img_shifted']) < 1e-4
    # data gradient w.r.t. the shift (what the SupportSets warp receives)
    cot = torch.randn(fx['img'].shape, generator=gen(1))
    so = fx['shift'].clone().requires_grad_(True)
    (o_sn.generate(sd, fx['z'], so, fx['model']) * cot).sum().backward()
    sc = fx['shift'].cuda().requires_grad_(True)
    (W(fx['z'].cuda(), sc) * cot.cuda()).sum().backward()
    cos = float(torch.nn.functional.cosine_similarity(sc.grad.flatten().double().cpu(), so.grad.flatten().double(), dim=0))
    assert cos > 0.9999 and rel(sc.grad, so.grad) < 1e-2


def test_biggan_fixture(golden):
    from warpedganspace_b200.generators import BigGANGenerator
    fx = golden('biggan_128.pt')
    G = BigGANGenerator()
    G.load_state_dict(o_bg.init_state(128, generator=gen(fx['seed'])), strict=True)
    G.cuda().eval()
    assert G.dim_z == 120
    with torch.no_grad():
        img = G(fx['z'].cuda() + fx['shift'].cuda(), G.shared(fx['classes'].cuda()))
    assert rel(img[:, :, ::fx['stride'], ::fx['stride']], fx['img']) < 2e-4
    assert abs(float(img.double().std()) - fx['std']) < 1e-3 * fx['std']


def test_biggan_wrapper_single_class():
    from warpedganspace_b200.generators import BigGANGenerator
    from warpedganspace_b200.gan_load import BigGANWrapper
    sd = o_bg.init_state(128, generator=gen(3))
    G = BigGANGenerator()
    G.load_state_dict(sd)
    W = BigGANWrapper(G, target_classes=(239,)).cuda().eval()
    z = torch.randn(2, 120, generator=gen(4))
    with torch.no_grad():
        got = W(z.cuda())
        want = o_bg.generate(sd, z, torch.tensor([239, 239]))
    assert rel(got, want) < 2e-4
