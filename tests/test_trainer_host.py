"""CPU: host logic of the training driver (warpedganspace_b200.Trainer / aux) against the fixture written by the
UNMODIFIED reference driver (tests/golden/trainer_c1.pt, oracle/gen_golden.py::pin_trainer_loop): experiment
directory name, host draw order, files written, checkpoint layout, stats.json windows and resume.  The CUDA engine is
replaced by a stub (no compute happens here); tests/test_trainer_gpu.py runs the real thing."""
import argparse
import json
import os

import pytest
import torch
from torch import nn

import oracle.step as o_step
from warpedganspace_b200 import aux
from warpedganspace_b200.trainer import Trainer


def _params(fx, **over):
    p = dict(fx['params'])
    p.update(over)
    return argparse.Namespace(**p)


def test_exp_dir_name_matches_reference(golden):
    fx = golden('trainer_c1.pt')
    assert aux.exp_dir_name(_params(fx)) == fx['exp_dir']
    p = _params(fx, gan_type='StyleGAN2', stylegan2_resolution=1024, shift_in_w_space=True, reconstructor_type='ResNet',
                num_support_sets=128, num_support_dipoles=32, learn_alphas=True, min_shift_magnitude=0.1,
                max_shift_magnitude=0.2)
    assert aux.exp_dir_name(p) == 'StyleGAN2-1024-W-ResNet-K128-D32-LearnAlphas-LearnGammas-eps0.1_0.2'
    p = _params(fx, gan_type='BigGAN', biggan_target_classes=[239, 7], learn_gammas=False)
    assert aux.exp_dir_name(p).startswith('BigGAN-2397-LeNet-K32-D16-eps')


def test_create_exp_dir_writes_args_and_command(golden, tmp_path):
    fx = golden('trainer_c1.pt')
    name = aux.create_exp_dir(_params(fx), root=str(tmp_path / 'experiments'))
    wip = tmp_path / 'experiments' / 'wip' / name
    assert json.load(open(wip / 'args.json'))['num_support_sets'] == fx['K']
    assert open(wip / 'command.sh').read().startswith('#!/usr/bin/bash\n')


def test_sample_z_branches():
    torch.manual_seed(5)
    a = aux.sample_z(3, 7)
    torch.manual_seed(5)
    assert torch.equal(a, torch.randn(3, 7))
    torch.manual_seed(5)
    assert torch.equal(aux.sample_z(3, 7, truncation=1.0), a)          # truncation 1.0 is the untruncated branch
    t = aux.sample_z(64, 16, truncation=0.7)
    assert t.dtype == torch.float32 and t.shape == (64, 16) and float(t.abs().max()) <= 0.7
    assert aux.sample_z(0, 4).shape == (0, 4)


def test_draw_order_is_the_reference_order(golden, tmp_path):
    """z, then indices, then the magnitude pool (+, -) and its multinomial pick — the same stream the oracle's
    generator-based restatement consumes (global seed == Generator seed on CPU)."""
    fx = golden('trainer_c1.pt')
    T = Trainer(_params(fx), fx['exp_dir'], use_cuda=True, root=str(tmp_path))
    torch.manual_seed(77)
    z, idx, mag = T.draw_batch(fx['d'])
    g = torch.Generator().manual_seed(77)
    assert torch.equal(z, torch.randn(fx['B'], fx['d'], generator=g))
    assert torch.equal(idx, torch.randint(0, fx['K'], (fx['B'],), generator=g))
    assert torch.equal(mag, o_step.sample_shift_magnitudes(fx['B'], 0.15, 0.25, generator=g))
    lo, hi = fx['params']['min_shift_magnitude'], fx['params']['max_shift_magnitude']
    assert bool(((mag.abs() >= lo) & (mag.abs() <= hi)).all())


class _StubEngine:
    """Stands in for PairedTrainer: nudges the parameters and reports deterministic statistics."""
    _graph = None

    def __init__(self, S, R):
        self.S, self.R, self.calls = S, R, 0

    def step(self, z, indices, magnitudes):
        self.calls += 1
        with torch.no_grad():
            self.S.SUPPORT_SETS.add_(1.0)
            self.R.w.add_(1.0)
        c = float(self.calls)
        return dict(accuracy=torch.tensor(0.25), cls=torch.tensor(c), reg=torch.tensor(2 * c), loss=torch.tensor(3 * c))


class _S(nn.Module):
    def __init__(self):
        super().__init__()
        self.SUPPORT_SETS = nn.Parameter(torch.zeros(2, 3))


class _R(nn.Module):
    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.zeros(4))


class _G(nn.Module):
    dim_z = 8


def _run(fx, root, max_iter, monkeypatch, S=None, R=None):
    p = _params(fx, max_iter=max_iter, quiet=True)
    T = Trainer(p, fx['exp_dir'], use_cuda=True, root=root)
    S, R = S or _S(), R or _R()
    eng = {}
    monkeypatch.setattr(T, '_device', lambda: torch.device('cpu'))
    monkeypatch.setattr(T, '_make_engine', lambda g, s, r: eng.setdefault('e', _StubEngine(s, r)))
    T.train(_G(), S, R)
    return T, eng['e'], S, R


def test_files_stats_and_checkpoint_layout_match_reference(golden, tmp_path, monkeypatch):
    fx = golden('trainer_c1.pt')
    root = str(tmp_path / 'experiments')
    aux.create_exp_dir(_params(fx), root=root)
    T, eng, S, R = _run(fx, root, fx['iters'], monkeypatch)
    wip = os.path.join(root, 'wip', fx['exp_dir'])
    done = os.path.join(root, 'complete', fx['exp_dir'])
    listing = lambda d: sorted(os.path.relpath(os.path.join(r, f), d) for r, _, fs in os.walk(d) for f in fs)
    assert listing(wip) == fx['files_wip']
    assert listing(done) == fx['files_complete'] and 'models/checkpoint.pt' not in listing(done)
    stats = json.load(open(os.path.join(wip, 'stats.json')))
    assert sorted(stats) == sorted(fx['stats']) and sorted(stats['2']) == sorted(fx['stats']['2'])
    assert stats['4']['classification_loss'] == pytest.approx(3.5) and stats['4']['total_loss'] == pytest.approx(10.5)
    ckpt = torch.load(os.path.join(wip, 'models', 'checkpoint.pt'))
    assert ckpt['iter'] == fx['checkpoint_iter'] == 6 and sorted(ckpt) == sorted(fx['checkpoint_keys'])
    assert float(ckpt['support_sets']['SUPPORT_SETS'][0, 0]) == 6.0
    assert float(torch.load(os.path.join(wip, 'models', 'support_sets_init.pt'))['SUPPORT_SETS'].abs().sum()) == 0.0
    assert float(torch.load(os.path.join(wip, 'models', 'support_sets.pt'))['SUPPORT_SETS'][1, 2]) == 6.0
    assert float(torch.load(os.path.join(wip, 'models', 'reconstructor.pt'))['w'][0]) == 6.0


def test_resume_from_checkpoint_and_completed_experiment(golden, tmp_path, monkeypatch):
    fx = golden('trainer_c1.pt')
    root = str(tmp_path / 'experiments')
    _run(fx, root, 3, monkeypatch)                          # checkpoint at iteration 3 (ckp_freq 3)
    T, eng, S, R = _run(fx, root, 5, monkeypatch)           # resumes AT iteration 3 (reference quirk: 3 runs again)
    assert eng.calls == 3 and float(S.SUPPORT_SETS[0, 0]) == 3.0 + 3.0
    with pytest.raises(SystemExit):                         # starting_iter == max_iter -> "already completed"
        _run(fx, root, 3, monkeypatch)


def test_no_cpu_path(golden, tmp_path):
    fx = golden('trainer_c1.pt')
    T = Trainer(_params(fx), fx['exp_dir'], use_cuda=False, root=str(tmp_path))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        T.train(_G(), _S(), _R())
    with pytest.raises(ValueError):
        Trainer(None, 'x')


def test_checkpoint_to_models_and_load_experiment(golden, tmp_path, monkeypatch):
    """checkpoint2model.py and the experiment loader of traverse_latent_space.py:252-297 on a driver-written tree."""
    from warpedganspace_b200 import SupportSets
    from warpedganspace_b200.checkpoint import checkpoint_to_models, load_experiment
    fx = golden('trainer_c1.pt')
    root = str(tmp_path / 'experiments')
    p = _params(fx, max_iter=3, quiet=True)
    aux.create_exp_dir(p, root=root)
    S = SupportSets(fx['K'], fx['D'], fx['d'], learn_gammas=True, gamma=1.0 / fx['d'])
    T = Trainer(p, fx['exp_dir'], use_cuda=True, root=root)
    monkeypatch.setattr(T, '_device', lambda: torch.device('cpu'))
    monkeypatch.setattr(T, '_make_engine', lambda g, s, r: _StubEngine(s, _R()))
    T.train(_G(), S, _R())
    wip = os.path.join(root, 'wip', fx['exp_dir'])
    assert checkpoint_to_models(wip) == 3
    ckpt = torch.load(os.path.join(wip, 'models', 'checkpoint.pt'))
    assert torch.equal(torch.load(os.path.join(wip, 'models', 'support_sets-3.pt'))['SUPPORT_SETS'],
                       ckpt['support_sets']['SUPPORT_SETS'])
    args, S2 = load_experiment(wip)
    assert args.num_support_sets == fx['K'] and S2.support_vectors_dim == fx['d'] and S2.learn_gammas
    assert torch.equal(S2.SUPPORT_SETS.detach(), torch.load(os.path.join(wip, 'models', 'support_sets.pt'))['SUPPORT_SETS'])
    _, S3 = load_experiment(wip, iteration=3)
    assert torch.equal(S3.SUPPORT_SETS.detach(), ckpt['support_sets']['SUPPORT_SETS'])
    with pytest.raises(NotADirectoryError):
        checkpoint_to_models(str(tmp_path / 'nope'))
    with pytest.raises(FileNotFoundError):
        os.makedirs(tmp_path / 'e2' / 'models')
        checkpoint_to_models(str(tmp_path / 'e2'))
