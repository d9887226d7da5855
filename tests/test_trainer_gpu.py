"""GPU: warpedganspace_b200.Trainer (the drop-in for lib/trainer.py's driver) against the fixture written by the
UNMODIFIED reference ``Trainer.train`` on config 1 (SNGAN-MNIST 32x32, K=32, D=16, LeNet, batch 4, 6 iterations, CPU,
global seed 2024): same host random stream, so the per-window statistics in stats.json, the files on disk and the
parameters after six Adam steps must agree."""
import argparse
import json
import os

import pytest
import torch

import oracle.support_sets as o_ss
import oracle.reconstructor as o_rec
import oracle.sngan as o_sn

pytestmark = pytest.mark.gpu


def gen(seed):
    return torch.Generator().manual_seed(seed)


def _build(fx):
    from warpedganspace_b200 import SupportSets
    from warpedganspace_b200.generators import SNGANGenerator
    from warpedganspace_b200.gan_load import SNGANWrapper
    from warpedganspace_b200.reconstructor import Reconstructor
    sg, ss, sr = fx['seeds']
    g_sd = o_sn.init_state('sn_resnet32', 1, generator=gen(sg))
    s_sd = o_ss.init_state(fx['K'], fx['D'], fx['d'], generator=gen(ss))
    r_sd = o_rec.init_state('LeNet', fx['K'], 1, generator=gen(sr))
    G = SNGANGenerator('sn_resnet32', 32, 1)
    G.load_state_dict({'model.' + k: v for k, v in g_sd.items()}, strict=False)
    S = SupportSets(fx['K'], fx['D'], fx['d'], learn_gammas=True, gamma=1.0 / fx['d'])
    S.load_state_dict(s_sd)
    R = Reconstructor('LeNet', fx['K'], 1)
    R.load_state_dict(r_sd)
    return SNGANWrapper(G), S, R, s_sd


@pytest.mark.parametrize('graph', [False, True])
def test_trainer_matches_reference_driver(golden, tmp_path, graph):
    from warpedganspace_b200 import aux
    from warpedganspace_b200.trainer import Trainer
    fx = golden('trainer_c1.pt')
    params = argparse.Namespace(**dict(fx['params'], quiet=True, cuda_graph=graph))
    root = str(tmp_path / 'experiments')
    exp_dir = aux.create_exp_dir(params, root=root)
    assert exp_dir == fx['exp_dir']
    G, S, R, s_sd = _build(fx)
    trn = Trainer(params=params, exp_dir=exp_dir, use_cuda=True, multi_gpu=False, root=root)
    torch.manual_seed(fx['train_seed'])
    trn.train(generator=G, support_sets=S, reconstructor=R)
    wip = os.path.join(root, 'wip', exp_dir)
    listing = lambda d: sorted(os.path.relpath(os.path.join(r, f), d) for r, _, fs in os.walk(d) for f in fs)
    assert listing(wip) == fx['files_wip']
    assert listing(os.path.join(root, 'complete', exp_dir)) == fx['files_complete']
    stats = json.load(open(os.path.join(wip, 'stats.json')))
    assert sorted(stats) == sorted(fx['stats'])
    for it, want in fx['stats'].items():
        for key, v in want.items():
            tol = 1e-6 if key == 'accuracy' else 2e-3 * abs(v) + 1e-5      # fp32 step parity, 1e-3 relative (north star)
            assert abs(stats[it][key] - v) <= tol, (it, key, stats[it][key], v)
    # parameters after six Adam steps: the rows the reference moved are the rows we moved, in the same direction
    final = torch.load(os.path.join(wip, 'models', 'support_sets.pt'))
    moved = (final['SUPPORT_SETS'] - s_sd['SUPPORT_SETS']).abs().amax(dim=1).nonzero().flatten()
    assert torch.equal(moved, fx['moved_rows'])
    d_got = (final['SUPPORT_SETS'][moved] - s_sd['SUPPORT_SETS'][moved]).double().flatten()
    d_want = (fx['final_support_sets_rows'] - s_sd['SUPPORT_SETS'][moved]).double().flatten()
    cos = float(torch.nn.functional.cosine_similarity(d_got, d_want, dim=0))
    print('support-set displacement after %d steps: cos %.5f, rel %.3e' % (fx['iters'], cos, float((d_got - d_want).norm() / d_want.norm())))
    assert cos > 0.95                          # Adam's first updates are ~lr*sign(g): near-zero gradient entries may flip
    assert float((final['LOGGAMMA'] - fx['final_loggamma']).abs().max()) < 4e-4          # |step| <= lr per Adam update, 6 updates
    rfin = torch.load(os.path.join(wip, 'models', 'reconstructor.pt'))
    assert float((rfin['path_indices.3.bias'] - fx['final_head_bias']).abs().max()) < 4e-4
    assert float((rfin['feature_extractor.0.weight'] - fx['final_conv0_weight']).abs().max()) < 6.5e-4
    ckpt = torch.load(os.path.join(wip, 'models', 'checkpoint.pt'))
    assert ckpt['iter'] == fx['checkpoint_iter']
    assert {k: sorted(v.keys()) if isinstance(v, dict) else None for k, v in ckpt.items()} == fx['checkpoint_keys']
