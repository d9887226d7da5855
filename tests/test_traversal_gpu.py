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
