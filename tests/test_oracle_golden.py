"""CPU: the oracle restatement must reproduce the fixtures generated from the unmodified
reference by oracle/gen_golden.py (the reference has no golden vectors of its own)."""
import torch
import torch.nn.functional as F
import pytest

import oracle.support_sets as o_ss
import oracle.stylegan2 as o_sg2
import oracle.proggan as o_pg
import oracle.sngan as o_sn
import oracle.biggan as o_bg
import oracle.reconstructor as o_rec
import oracle.step as o_step


def gen(seed):
    return torch.Generator().manual_seed(seed)


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def test_support_sets_tiny(golden):
    fx = golden('support_sets_tiny.pt')
    z = fx['z'].clone().requires_grad_(True)
    sd = {k: v.clone().requires_grad_(k != 'ALPHAS') for k, v in fx['state'].items()}
    out = o_ss.forward(sd, o_ss.one_hot(fx['idx'], fx['K']), z)
    assert rel(out, fx['out']) < 1e-6
    (out * fx['cot']).sum().backward()
    assert rel(sd['SUPPORT_SETS'].grad, fx['d_support_sets']) < 1e-5
    assert rel(sd['LOGGAMMA'].grad, fx['d_loggamma']) < 1e-5
    assert rel(z.grad, fx['d_z']) < 1e-5
    assert torch.allclose(out.norm(dim=1), torch.ones(out.shape[0]), atol=1e-6)


def test_support_sets_benchmark_shape(golden):
    fx = golden('support_sets_c3.pt')
    sd = o_ss.init_state(fx['K'], fx['D'], fx['d'], generator=gen(fx['seed']))
    assert abs(checksum(sd) - fx['state_checksum']) < 1e-6 * fx['state_checksum']
    mask = o_ss.one_hot(fx['idx'], fx['K'])
    assert rel(o_ss.forward(sd, mask, fx['z']), fx['out']) < 1e-6
    assert rel(o_ss.forward(sd, mask, fx['z'], learn_gammas=False, gamma=1.0 / fx['d']), fx['out_fixed_gamma']) < 1e-6
    leaf = {k: v.clone().requires_grad_(k != 'ALPHAS') for k, v in sd.items()}
    z = fx['z'].clone().requires_grad_(True)
    (o_ss.forward(leaf, mask, z) * fx['cot']).sum().backward()
    g = leaf['SUPPORT_SETS'].grad
    assert rel(g[fx['rows']], fx['d_support_sets_rows']) < 1e-5
    # backward touches only the selected rows
    assert abs(float(g.double().abs().sum()) - fx['d_support_sets_abs_sum']) < 1e-5 * fx['d_support_sets_abs_sum']
    assert rel(z.grad, fx['d_z']) < 1e-5


def test_stylegan2_ops(golden):
    fx = golden('stylegan2_ops.pt')
    for name, case in fx['cases'].items():
        assert rel(o_sg2.upfirdn2d(fx['x'], fx['kernel'], **case['kw']), case['out']) < 1e-6, name
    assert rel(o_sg2.fused_leaky_relu(fx['x'], fx['bias']), fx['lrelu']) < 1e-7


@pytest.mark.parametrize('size', [32, 128])
def test_stylegan2_generator(golden, size):
    fx = golden('stylegan2_%d.pt' % size)
    sd = o_sg2.init_state(size=size, generator=gen(fx['seed']))
    assert abs(checksum(sd) - fx['state_checksum']) < 1e-6 * fx['state_checksum']
    st = fx.get('stride', 1)
    with torch.no_grad():
        assert rel(o_sg2.mapping(sd, fx['z']), fx['w']) < 1e-5
        for tag, wspace in (('z', False), ('w', True)):
            img = o_sg2.generate(sd, fx['z'], None, size, wspace)
            img_s = o_sg2.generate(sd, fx['z'], fx['shift'], size, wspace)
            assert rel(img[:, :, ::st, ::st], fx['img_' + tag]) < 2e-5
            assert rel(img_s[:, :, ::st, ::st], fx['img_shifted_' + tag]) < 2e-5
            if st > 1:
                assert abs(float(img.double().mean()) - fx['img_%s_mean' % tag]) < 1e-5
                assert abs(float(img_s.double().std()) - fx['img_shifted_%s_std' % tag]) < 1e-4 * fx['img_shifted_%s_std' % tag]


@pytest.mark.parametrize('size', [32, 64])
def test_sngan(golden, size):
    fx = golden('sngan_%d.pt' % size)
    sd = o_sn.init_state(fx['model'], fx['channels'], generator=gen(fx['seed']))
    assert abs(checksum(sd) - fx['state_checksum']) < 1e-6 * fx['state_checksum']
    with torch.no_grad():
        assert rel(o_sn.generate(sd, fx['z'], None, fx['model']), fx['img']) < 1e-5
        assert rel(o_sn.generate(sd, fx['z'], fx['shift'], fx['model']), fx['img_shifted']) < 1e-5


def test_proggan(golden):
    fx = golden('proggan_1024.pt')
    for tag, like in (('like', True), ('ctor', False)):
        c = fx[tag]
        sd = o_pg.init_state(generator=gen(c['seed']), pretrained_like=like)
        assert abs(checksum(sd) - c['state_checksum']) < 1e-6 * c['state_checksum']
        with torch.no_grad():
            img = o_pg.generate(sd, fx['z'], fx['shift'])
        assert tuple(img.shape) == (1, 3, 1024, 1024)
        assert rel(img[:, :, ::fx['stride'], ::fx['stride']], c['img']) < 2e-5
        assert abs(float(img.double().std()) - c['std']) < 1e-4 * c['std']


def test_biggan(golden):
    fx = golden('biggan_128.pt')
    sd = o_bg.init_state(128, generator=gen(fx['seed']))
    assert abs(checksum(sd) - fx['state_checksum']) < 1e-6 * fx['state_checksum']
    with torch.no_grad():
        img = o_bg.generate(sd, fx['z'], fx['classes'], fx['shift'])
    assert rel(img[:, :, ::fx['stride'], ::fx['stride']], fx['img']) < 2e-5
    assert abs(float(img.double().std()) - fx['std']) < 1e-4 * fx['std']


@pytest.mark.parametrize('name', ['resnet', 'lenet'])
def test_reconstructor(golden, name):
    fx = golden('reconstructor_%s.pt' % name)
    sd = o_rec.init_state(fx['type'], fx['dim'], fx['channels'], generator=gen(fx['seed']))
    assert abs(checksum(sd) - fx['state_checksum']) < 1e-6 * fx['state_checksum']
    keys = o_rec.trainable_keys(sd)
    leaf = dict(sd)
    for k in keys:
        leaf[k] = sd[k].clone().requires_grad_(True)
    x1 = fx['x1'].clone().requires_grad_(True)
    x2 = fx['x2'].clone().requires_grad_(True)
    running = {}
    logits, mag = o_rec.forward(leaf, x1, x2, fx['type'], True, running)
    assert rel(logits, fx['logits']) < 1e-5 and rel(mag, fx['mag']) < 1e-5
    assert torch.equal(logits.argmax(1), fx['logits'].argmax(1))
    loss = F.cross_entropy(logits, fx['idx']) + 0.25 * (mag - fx['tgt']).abs().mean()
    assert rel(loss, fx['loss']) < 1e-6
    loss.backward()
    assert rel(x1.grad, fx['dx1']) < 1e-4 and rel(x2.grad, fx['dx2']) < 1e-4
    for k, n in fx['grad_norms'].items():
        assert abs(float(leaf[k].grad.double().norm()) - n) <= 2e-4 * max(n, 1e-12), k
    for k, v in fx['running'].items():
        assert rel(running[k], v) < 1e-5, k
    assert not any('.fc.' in k for k in keys)


def test_paired_step_config1(golden):
    fx = golden('step_c1.pt')
    sg, ss, sr = fx['seeds']
    g_sd = o_sn.init_state('sn_resnet32', 1, generator=gen(sg))
    s_sd = o_ss.init_state(fx['K'], fx['D'], fx['d'], generator=gen(ss))
    r_sd = o_rec.init_state('LeNet', fx['K'], 1, generator=gen(sr))
    for sd, c in zip((g_sd, s_sd, r_sd), fx['checksums']):
        assert abs(checksum(sd) - c) < 1e-6 * c
    gen_fn, _ = o_step.make_generator('SNGAN', g_sd, model='sn_resnet32')
    res = o_step.paired_step(gen_fn, s_sd, r_sd, fx['z'], fx['idx'], fx['mag'], reconstructor_type='LeNet')
    for k in ('img', 'img_shifted', 'shift', 'logits', 'loss'):
        assert rel(res[k], fx[k]) < 1e-5, k
    assert rel(res['cls_loss'], fx['cls']) < 1e-6 and rel(res['reg_loss'], fx['reg']) < 1e-6
    assert rel(res['grads']['S']['SUPPORT_SETS'][fx['rows']], fx['d_support_sets_rows']) < 1e-4
    assert rel(res['grads']['S']['LOGGAMMA'], fx['d_loggamma']) < 1e-4
    for k, n in fx['r_grad_norms'].items():
        assert abs(float(res['grads']['R'][k].double().norm()) - n) <= 2e-4 * max(n, 1e-12), k
    p = s_sd['SUPPORT_SETS'].clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    o_step.adam_update(p, res['grads']['S']['SUPPORT_SETS'], m, v, 1)
    assert rel((p - s_sd['SUPPORT_SETS'])[fx['rows']],
               (fx['new_support_sets_rows'] - s_sd['SUPPORT_SETS'][fx['rows']])) < 1e-3


def test_traversal_chain(golden):
    fx = golden('traversal_chain.pt')
    sd = o_ss.init_state(fx['K'], fx['D'], fx['d'], generator=gen(fx['seed']))
    codes, shifts = o_step.traverse_chain(sd, fx['z0'], fx['path'], fx['eps'], fx['steps'])
    assert rel(codes, fx['codes']) < 1e-6 and rel(shifts, fx['shifts']) < 1e-5
    # every step has length eps (unit-norm direction, support_sets.py:101)
    n = shifts.norm(dim=1)
    n = torch.cat([n[:fx['steps']], n[fx['steps'] + 1:]])
    assert torch.allclose(n, torch.full_like(n, fx['eps']), atol=1e-6)


def test_shift_magnitude_sampler_quirk():
    """lib/trainer.py:218-221: multinomial weights are the pool indices, so pool[0] is never drawn."""
    g = gen(5)
    for _ in range(50):
        m = o_step.sample_shift_magnitudes(4, 0.1, 0.2, generator=g)
        assert m.shape == (4,) and bool(((m.abs() >= 0.1) & (m.abs() <= 0.2)).all())
