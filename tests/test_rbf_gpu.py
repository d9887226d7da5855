"""GPU parity: the RBF warp kernels (through the C ABI) against the oracle and the reference fixtures."""
import pytest
import torch

import oracle.support_sets as o_ss
import oracle.step as o_step

pytestmark = pytest.mark.gpu


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def make(K, D, d, seed, learn_gammas=True):
    from warpedganspace_b200 import SupportSets
    sd = o_ss.init_state(K, D, d, generator=gen(seed))
    S = SupportSets(K, D, d, learn_alphas=False, learn_gammas=learn_gammas, gamma=1.0 / d)
    S.load_state_dict(sd)
    return sd, S.cuda()


def test_fixture_tiny(golden):
    """K=6, D=3, d=10 (not a multiple of 4: the guarded scalar-load path), state written by the reference constructor
    itself: outputs and all three gradients of the unmodified reference module."""
    fx = golden('support_sets_tiny.pt')
    from warpedganspace_b200 import SupportSets
    S = SupportSets(fx['K'], fx['D'], fx['d'], learn_gammas=True, gamma=1.0 / fx['d'])
    S.load_state_dict(fx['state'])
    S.cuda()
    z = fx['z'].cuda().requires_grad_(True)
    out = S(o_ss.one_hot(fx['idx'], fx['K']).cuda(), z)
    assert rel(out, fx['out']) < 2e-6
    (out * fx['cot'].cuda()).sum().backward()
    assert rel(S.SUPPORT_SETS.grad, fx['d_support_sets']) < 2e-5
    assert rel(S.LOGGAMMA.grad, fx['d_loggamma']) < 2e-4
    assert rel(z.grad, fx['d_z']) < 1e-4


def test_fixture_benchmark_shape(golden):
    """K=128, D=32, d=512 against outputs and gradients of the unmodified reference."""
    fx = golden('support_sets_c3.pt')
    sd, S = make(fx['K'], fx['D'], fx['d'], fx['seed'])
    z = fx['z'].cuda().requires_grad_(True)
    out = S(o_ss.one_hot(fx['idx'], fx['K']).cuda(), z)
    assert rel(out, fx['out']) < 2e-6                           # fp32; tolerance 2e-6 relative L2
    (out * fx['cot'].cuda()).sum().backward()
    g = S.SUPPORT_SETS.grad
    assert rel(g[fx['rows'].cuda()], fx['d_support_sets_rows']) < 1e-5
    untouched = torch.ones(fx['K'], dtype=torch.bool)
    untouched[fx['rows']] = False
    assert float(g[untouched.cuda()].abs().max()) == 0.0        # only the selected rows get gradient
    assert rel(S.LOGGAMMA.grad, fx['d_loggamma']) < 1e-4
    assert rel(z.grad, fx["d_z"]) < 1e-4                             # sum of 2D alternating-sign terms: order-dependent in fp32
    assert S.ALPHAS.grad is None
    # fixed-gamma branch (learn_gammas=False, lib/support_sets.py:93)
    _, S2 = make(fx['K'], fx['D'], fx['d'], fx['seed'], learn_gammas=False)
    out2 = S2(o_ss.one_hot(fx['idx'], fx['K']).cuda(), fx['z'].cuda())
    assert rel(out2, fx['out_fixed_gamma']) < 2e-6


@pytest.mark.parametrize('K,D,d,B', [(32, 16, 128, 4), (120, 256, 120, 8), (200, 64, 512, 12), (7, 1, 4, 3),
                                     (16, 5, 1024, 2), (128, 32, 512, 64),
                                     (120, 256, 119, 8), (9, 3, 7, 3), (5, 2, 130, 2)])     # d % 4 != 0: BigGAN-256's dim_z = 119
def test_against_oracle(K, D, d, B):
    sd, S = make(K, D, d, 100 + K)
    g = gen(7 + d)
    z = torch.randn(B, d, generator=g)
    idx = torch.randint(0, K, (B,), generator=g)
    idx[-1] = idx[0]                                            # repeated path inside a batch (atomics)
    mag = o_step.sample_shift_magnitudes(B, 0.1, 0.2, generator=g)
    cot = torch.randn(B, d, generator=g)
    # oracle with autograd
    leaf = {k: v.clone().requires_grad_(k != 'ALPHAS') for k, v in sd.items()}
    zo = z.clone().requires_grad_(True)
    want = mag.reshape(-1, 1) * o_ss.forward(leaf, o_ss.one_hot(idx, K), zo)
    (want * cot).sum().backward()
    # product, fused-magnitude fast path
    zc = z.cuda().requires_grad_(True)
    got = S.warp(idx.cuda(), zc, mag.cuda())
    assert rel(got, want) < (3e-6 if D <= 64 else 6e-6)        # fp32 sum over 2D vectors
    (got * cot.cuda()).sum().backward()
    assert rel(S.SUPPORT_SETS.grad, leaf['SUPPORT_SETS'].grad) < 2e-5
    assert rel(S.LOGGAMMA.grad, leaf['LOGGAMMA'].grad) < 2e-4
    assert rel(zc.grad, zo.grad) < 1e-4
    # unit norm of the bare module output
    u = S(o_ss.one_hot(idx, K).cuda(), z.cuda())
    assert torch.allclose(u.norm(dim=1), torch.ones(B, device='cuda'), atol=1e-5)


def test_learn_alphas_gradient():
    from warpedganspace_b200 import SupportSets
    K, D, d, B = 8, 4, 64, 5
    sd = o_ss.init_state(K, D, d, generator=gen(3))
    S = SupportSets(K, D, d, learn_alphas=True, learn_gammas=True, gamma=1.0 / d)
    S.load_state_dict(sd)
    S.cuda()
    g = gen(4)
    z = torch.randn(B, d, generator=g)
    idx = torch.randint(0, K, (B,), generator=g)
    cot = torch.randn(B, d, generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (o_ss.forward(leaf, o_ss.one_hot(idx, K), z) * cot).sum().backward()
    (S(o_ss.one_hot(idx, K).cuda(), z.cuda()) * cot.cuda()).sum().backward()
    assert rel(S.ALPHAS.grad, leaf['ALPHAS'].grad) < 2e-5


def test_traversal_fixture(golden):
    fx = golden('traversal_chain.pt')
    sd, S = make(fx['K'], fx['D'], fx['d'], fx['seed'])
    codes, shifts = S.traverse(fx['z0'].cuda(), torch.tensor([fx['path']]).cuda(), fx['eps'], fx['steps'])
    assert rel(codes[0], fx['codes']) < 2e-6
    assert rel(shifts[0], fx['shifts']) < 1e-5


def test_traversal_many_chains():
    """Config-5 shaped slice: many (latent, path) chains in one launch, vs the oracle chain by chain;
    plus the size-independent property that every step has length eps."""
    K, D, d, steps, eps = 128, 32, 512, 16, 0.15
    sd, S = make(K, D, d, 55)
    g = gen(56)
    nz = 3
    z = torch.randn(nz, d, generator=g)
    paths = torch.tensor([0, 17, 127])
    start = z.repeat_interleave(len(paths), dim=0)
    pp = paths.repeat(nz)
    codes, shifts = S.traverse(start.cuda(), pp.cuda(), eps, steps)
    for c in (0, 4, 8):
        wc, ws = o_step.traverse_chain(sd, start[c:c + 1], int(pp[c]), eps, steps)
        assert rel(codes[c], wc) < 1e-5 and rel(shifts[c], ws) < 1e-4
    n = shifts.norm(dim=2)
    n = torch.cat([n[:, :steps], n[:, steps + 1:]], dim=1)
    assert torch.allclose(n, torch.full_like(n, eps), atol=1e-6)
    assert float(shifts[:, steps].abs().max()) == 0.0


@pytest.mark.parametrize('d,leap', [(512, 1), (512, 4), (119, 3)])
def test_traversal_shift_leap_and_unaligned_dims(d, leap):
    """--shift_leap (traverse_latent_space.py:404,434): only every leap-th step of each direction is kept; also the
    scalar-tail path for latent dimensions that are not a multiple of 4 (BigGAN-256: dim_z = 119)."""
    K, D, steps, eps = 12, 6, 12, 0.2
    sd, S = make(K, D, d, 70 + d)
    g = gen(71)
    start = torch.randn(5, d, generator=g)
    paths = torch.randint(0, K, (5,), generator=g)
    codes, shifts = S.traverse(start.cuda(), paths.cuda(), eps, steps, shift_leap=leap)
    assert codes.shape == (5, 2 * (steps // leap) + 1, d)
    for c in range(5):
        wc, ws = o_step.traverse_chain(sd, start[c:c + 1], int(paths[c]), eps, steps, shift_leap=leap)
        assert rel(codes[c], wc) < 1e-5 and rel(shifts[c], ws) < 1e-4
