"""GPU parity at the FULL benchmark size (BASELINE config 3: StyleGAN2-1024, K=128, D=32, ResNet-18 R at 1024^2).

The fixture tests/golden/stylegan2_1024_step.pt was written by the UNMODIFIED reference modules
(models/StyleGAN2/model.py Generator(1024, 512, 8) with its two CUDA ops restated on the CPU, lib/support_sets.py,
lib/reconstructor.py) running the loop body of lib/trainer.py:190-250 on injected draws (oracle/gen_golden.py::
pin_stylegan2_1024_step).  It exercises the kernel variants that only exist at >= 512-pixel-wide maps: the 16-tiles-per-CTA
multi-tile halo kernel (32 -> 32 at 1024^2), the stacked-weight 64 -> 64 kernel at 512^2, the phase-packed up-convs to
1025^2 and the FIR at 1024^2.

Tolerances.  Forward quantities: 1e-4 (images), 1e-3 (logits) relative L2, the north-star bar, against the reference's
fp32 values.  Whole-graph gradients at this size cannot be pinned at 1e-3 by ANY implementation: the reference's own
fp32 gradients differ from the fp64 evaluation of the same graph by 6.8e-3 (dSUPPORT_SETS) to 1.1e-2 (stem conv) - tens
of millions of ReLU / max-pool inputs include a few within rounding of a kink, and train-mode BatchNorm amplifies any
forward perturbation ~1000x into the gradients (tests/test_gradients_gpu.py).  The fixture therefore also stores, per
tensor, how far the oracle's gradients move under the kernels' arithmetic model (bf16 hi+lo operand rounding in the
forward convolutions, oracle/emulate.py): 1.9e-2 (dSUPPORT_SETS) to 2.9e-2.  Gradients are compared with the fp64 values
and must be no further from them than three times that distance; tests/test_gradients_gpu.py pins the same graph at
1e-3 on kink-free, well-conditioned draws."""
import pytest
import torch

import oracle.support_sets as o_ss
import oracle.stylegan2 as o_sg2
import oracle.reconstructor as o_rec

pytestmark = pytest.mark.gpu


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


@pytest.fixture(scope='module')
def world(golden):
    from warpedganspace_b200 import SupportSets
    from warpedganspace_b200.stylegan2 import Generator
    from warpedganspace_b200.reconstructor import Reconstructor
    fx = golden('stylegan2_1024_step.pt')
    sg, ss, sr = fx['seeds']
    g_sd = o_sg2.init_state(size=1024, generator=gen(sg))
    s_sd = o_ss.init_state(fx['K'], fx['D'], fx['d'], generator=gen(ss))
    r_sd = o_rec.init_state('ResNet', fx['K'], 3, generator=gen(sr))
    for sd, want in zip((g_sd, s_sd, r_sd), fx['checksums']):
        assert abs(_checksum(sd) - want) <= 1e-9 * want            # RNG drift guard: same weights as the fixture run
    G = Generator(1024, 512, 8)
    G.load_state_dict(g_sd, strict=False)
    S = SupportSets(fx['K'], fx['D'], fx['d'], learn_gammas=True, gamma=1.0 / fx['d'])
    S.load_state_dict(s_sd)
    R = Reconstructor('ResNet', fx['K'], 3)
    R.load_state_dict(r_sd)
    return fx, G.cuda(), S.cuda(), R.cuda()


def _check_image(img, fx, key, std_key):
    st = fx['stride']
    assert tuple(img.shape) == (fx['B'], 3, 1024, 1024)
    e = rel(img[:, :, ::st, ::st], fx[key])
    print('%s strided rel err %.2e' % (key, e))
    assert e < 1e-4, (key, e)
    assert abs(float(img.double().std()) - fx[std_key]) < 1e-4 * fx[std_key]


def test_generator_1024_matches_reference_fixture(world):
    """G(z), G(z, shift) in Z space and W space, and forward_pair, against the reference's 1024^2 images."""
    from warpedganspace_b200.gan_load import StyleGAN2Wrapper
    fx, G, _, _ = world
    z, shift = fx['z'].cuda(), fx['shift'].cuda()
    Wz = StyleGAN2Wrapper(G, shift_in_w_space=False).eval()
    with torch.no_grad():
        img = Wz(z)
        img_s = Wz(z, shift)
        pair = Wz.forward_pair(z, shift)
    _check_image(img, fx, 'img', 'img_std')
    _check_image(img_s, fx, 'img_shifted', 'img_shifted_std')
    assert rel(img[:, :, 517, :], fx['img_row']) < 1e-4 and rel(img_s[:, :, 517, :], fx['img_shifted_row']) < 1e-4
    assert abs(float(img.double().mean()) - fx['img_mean']) < 1e-4 * fx['img_std']
    _check_image(pair[0], fx, 'img', 'img_std')
    _check_image(pair[1], fx, 'img_shifted', 'img_shifted_std')
    assert rel(pair[0], img) < 1e-6 and rel(pair[1], img_s) < 1e-6
    Ww = StyleGAN2Wrapper(G, shift_in_w_space=True).eval()
    with torch.no_grad():
        img_w = Ww(z, fx['wshift'].cuda())
    _check_image(img_w, fx, 'img_w', 'img_w_std')


def test_paired_step_1024_matches_reference_fixture(world):
    """One full-size paired step (B = 2) through PairedTrainer against the reference loop body."""
    from warpedganspace_b200.gan_load import StyleGAN2Wrapper
    from warpedganspace_b200.trainer import PairedTrainer
    fx, G, S, R = world
    T = PairedTrainer(StyleGAN2Wrapper(G, shift_in_w_space=False), S, R)
    got = T.forward_backward(fx['z'].cuda(), fx['idx'].cuda(), fx['mag'].cuda())
    torch.cuda.synchronize()
    assert rel(got['shift'], fx['shift']) < 1e-5
    _check_image(got['img'], fx, 'img', 'img_std')
    _check_image(got['img_shifted'], fx, 'img_shifted', 'img_shifted_std')
    e_logits = rel(got['logits'], fx['logits'])
    print('logits rel err %.2e (vs fp64 %.2e), loss %.6f vs %.6f' % (e_logits, rel(got['logits'], fx['logits64']),
                                                                    float(got['loss']), float(fx['loss'])))
    assert e_logits < 1e-3 and rel(got['logits'], fx['logits64']) < 1e-3
    assert torch.equal(got['logits'].argmax(1).cpu(), fx['logits'].argmax(1))
    assert rel(got['loss'], fx['loss']) < 1e-4 and rel(got['cls'], fx['cls']) < 1e-4 and rel(got['reg'], fx['reg']) < 1e-3
    assert rel(got['pred'], fx['pred']) < 1e-3
    # gradients vs the fp64 evaluation, in units of the reference-fp32 distance to it (module docstring)
    rows = fx['rows']
    yard, model = fx['yardstick'], fx['yardstick_split17']
    e = rel(S.SUPPORT_SETS.grad[rows.cuda()], fx['d_support_sets_rows64'])
    print('dSUPPORT_SETS vs fp64 %.2e (reference fp32: %.2e, arithmetic model: %.2e)' % (e, yard['SUPPORT_SETS'], model['SUPPORT_SETS']))
    assert e < max(1e-3, 3 * model['SUPPORT_SETS'])
    gs, ws = S.SUPPORT_SETS.grad[rows.cuda()].double().cpu().flatten(), fx['d_support_sets_rows64'].double().flatten()
    assert float(torch.nn.functional.cosine_similarity(gs, ws, dim=0)) > 0.999
    e = rel(S.LOGGAMMA.grad[rows.cuda()], fx['d_loggamma64'][rows])
    assert e < max(1e-3, 3 * model['LOGGAMMA'])
    untouched = torch.ones(fx['K'], dtype=torch.bool)
    untouched[rows] = False
    assert float(S.SUPPORT_SETS.grad[untouched.cuda()].abs().max()) == 0.0
    params = dict(R.named_parameters())
    worst = (0.0, None)
    worst_model = max(ent['split17_vs_fp64'] for ent in fx['r_grads64'].values())
    for k, ent in fx['r_grads64'].items():
        g = params[k].grad.reshape(-1)[::ent['stride']]
        e = rel(g, ent['values'])
        bound = max(1e-3, 3 * max(ent['split17_vs_fp64'], 0.3 * worst_model))
        worst = max(worst, (e / bound, k))
        assert e < bound, (k, e, ent['split17_vs_fp64'])
        n = float(params[k].grad.double().norm())
        assert abs(n - ent['norm']) <= bound * ent['norm'], k
    print('worst dR (in units of its bound): %.2f at %s' % worst)
    sd = R.state_dict()
    for k, v in fx['running_sample'].items():
        assert rel(sd[k], v) < 1e-4, k
