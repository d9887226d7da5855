"""Pins the oracle against the UNMODIFIED reference and writes tests/golden/*.pt.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden

For every component it (1) builds the oracle's seeded random state, (2) loads it into the reference
module with ``load_state_dict`` (which also pins the key names / shapes), (3) runs the reference on
seeded inputs, (4) asserts the oracle agrees, and (5) stores inputs + REFERENCE outputs as a small
fixture.  Large states are not stored: fixtures carry the seed, and ``state_checksum`` guards against
RNG drift.  The reference has no CPU implementation of its two StyleGAN2 CUDA ops; they are replaced
at import time by the reference's own pure-PyTorch ``upfirdn2d_native`` (op/upfirdn2d.py:152-186) and
by ``scale * leaky_relu(x + b)`` (op/fused_bias_act_kernel.cu:25-47) — SURVEY.md §8c.
"""
import os
import sys
import types
import math
import importlib

import torch
import torch.nn.functional as F

REF = os.environ.get('WGS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def import_reference():
    """Import the reference package tree without building its CUDA extensions."""
    sys.dont_write_bytecode = True
    for name in ('skimage', 'skimage.io', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import torch.utils.cpp_extension as cpp_ext
    real_load = cpp_ext.load
    cpp_ext.load = lambda *a, **k: types.SimpleNamespace()          # no JIT build on a GPU-less box
    try:
        model = importlib.import_module('models.StyleGAN2.model')
    finally:
        cpp_ext.load = real_load
    up_mod = sys.modules['models.StyleGAN2.op.upfirdn2d']
    up_mod.F = F                                                     # the reference forgot this import

    def upfirdn2d_cpu(input, kernel, up=1, down=1, pad=(0, 0)):
        n, c, h, w = input.shape
        out = up_mod.upfirdn2d_native(input.reshape(-1, h, w, 1), kernel, up, up, down, down,
                                      pad[0], pad[1], pad[0], pad[1])
        return out.view(-1, c, out.shape[1], out.shape[2])

    def fused_lrelu_cpu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
        shape = [1, -1] + [1] * (input.ndim - 2)
        return scale * F.leaky_relu(input + bias.view(*shape), negative_slope)

    model.upfirdn2d = upfirdn2d_cpu
    model.fused_leaky_relu = fused_lrelu_cpu
    model.FusedLeakyReLU.forward = lambda self, x: fused_lrelu_cpu(x, self.bias, self.negative_slope, self.scale)
    return model, up_mod


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


def gen(seed):
    return torch.Generator().manual_seed(seed)


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print('  wrote %s (%.1f KB)' % (name, os.path.getsize(path) / 1024))


def check(tag, got, want, tol):
    e = rel_err(got, want)
    print('  %-44s rel err %.3e' % (tag, e))
    assert e <= tol, (tag, e, tol)


# ----------------------------------------------------------------------------------------------
def pin_support_sets():
    import oracle.support_sets as o
    ref = importlib.import_module('lib.support_sets')
    print('[support sets]')
    # (a) tiny, state produced by the reference constructor itself and stored in full
    torch.manual_seed(11)
    K, D, d, B = 6, 3, 10, 5
    S = ref.SupportSets(K, D, d, learn_alphas=False, learn_gammas=True, gamma=1.0 / d)
    sd = {k: v.detach().clone() for k, v in S.state_dict().items()}
    g = gen(12)
    z = torch.randn(B, d, generator=g).requires_grad_(True)
    idx = torch.randint(0, K, (B,), generator=g)
    mask = o.one_hot(idx, K)
    out = S(mask, z)
    cot = torch.randn(B, d, generator=g)
    (out * cot).sum().backward()
    got = o.forward(sd, mask, z.detach(), learn_gammas=True)
    check('tiny forward', got, out.detach(), 1e-6)
    # the oracle's own init must have the reference's structure (radii, antipodal pairs, alphas)
    mine = o.init_state(K, D, d, generator=gen(1))
    assert torch.equal(mine['ALPHAS'], sd['ALPHAS']) and torch.equal(mine['LOGGAMMA'], sd['LOGGAMMA'])
    r_ref = sd['SUPPORT_SETS'].view(K, 2 * D, d).norm(dim=2)
    r_mine = mine['SUPPORT_SETS'].view(K, 2 * D, d).norm(dim=2)
    check('init radii', r_mine, r_ref, 1e-6)
    pm = mine['SUPPORT_SETS'].view(K, D, 2, d)
    assert torch.equal(pm[:, :, 0], -pm[:, :, 1])
    save('support_sets_tiny.pt', dict(K=K, D=D, d=d, state=sd, z=z.detach(), idx=idx, cot=cot, out=out.detach(),
                                      d_support_sets=S.SUPPORT_SETS.grad.clone(), d_loggamma=S.LOGGAMMA.grad.clone(),
                                      d_z=z.grad.clone()))
    # (b) the benchmark shape, seed-regenerated state; also the fixed-gamma branch
    K, D, d, B = 128, 32, 512, 8
    sd = o.init_state(K, D, d, generator=gen(21))
    S = ref.SupportSets(K, D, d, learn_alphas=False, learn_gammas=True, gamma=1.0 / d)
    S.load_state_dict(sd)
    g = gen(22)
    z = torch.randn(B, d, generator=g).requires_grad_(True)
    idx = torch.randint(0, K, (B,), generator=g)
    mask = o.one_hot(idx, K)
    out = S(mask, z)
    cot = torch.randn(B, d, generator=g)
    (out * cot).sum().backward()
    check('K128 D32 d512 forward', o.forward(sd, mask, z.detach()), out.detach(), 1e-6)
    S2 = ref.SupportSets(K, D, d, learn_alphas=False, learn_gammas=False, gamma=1.0 / d)
    S2.load_state_dict(sd)
    out_fixed = S2(mask, z.detach())
    check('fixed-gamma forward', o.forward(sd, mask, z.detach(), learn_gammas=False, gamma=1.0 / d),
          out_fixed.detach(), 1e-6)
    rows = torch.unique(idx)
    save('support_sets_c3.pt', dict(K=K, D=D, d=d, seed=21, state_checksum=checksum(sd), z=z.detach(), idx=idx,
                                    cot=cot, out=out.detach(), out_fixed_gamma=out_fixed.detach(), rows=rows,
                                    d_support_sets_rows=S.SUPPORT_SETS.grad[rows].clone(),
                                    d_support_sets_abs_sum=float(S.SUPPORT_SETS.grad.double().abs().sum()),
                                    d_loggamma=S.LOGGAMMA.grad.clone(), d_z=z.grad.clone()))


def pin_stylegan2(model, up_mod):
    import oracle.stylegan2 as o
    gan_load = importlib.import_module('models.gan_load')
    print('[stylegan2]')
    # ops
    g = gen(31)
    x = torch.randn(2, 3, 9, 9, generator=g)
    k4 = o.fir_kernel() * 4
    cases = {}
    for name, kw in (('blur_pad11', dict(up=1, down=1, pad=(1, 1))), ('up2_pad21', dict(up=2, down=1, pad=(2, 1))),
                     ('down2_pad11', dict(up=1, down=2, pad=(1, 1))), ('crop', dict(up=1, down=1, pad=(-1, 2)))):
        want = model.upfirdn2d(x, k4, **kw)
        check('upfirdn2d ' + name, o.upfirdn2d(x, k4, **kw), want, 1e-6)
        cases[name] = dict(kw=kw, out=want)
    b = torch.randn(3, generator=g)
    want = model.fused_leaky_relu(x, b)
    check('fused_leaky_relu', o.fused_leaky_relu(x, b), want, 1e-7)
    save('stylegan2_ops.pt', dict(x=x, kernel=k4, cases=cases, bias=b, lrelu=want))
    # whole generator, Z space and W space, at 32 px (512 channels everywhere) and 128 px (256 ch at 128)
    for size, B in ((32, 2), (128, 1)):
        sd = o.init_state(size=size, generator=gen(32 + size))
        G = model.Generator(size, 512, 8)
        missing = G.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing
        assert all(k.endswith('.kernel') for k in missing.missing_keys), missing       # FIR buffers only
        G.eval()
        g = gen(33 + size)
        z = torch.randn(B, 512, generator=g)
        shift = 0.15 * F.normalize(torch.randn(B, 512, generator=g), dim=1)
        fx = dict(size=size, seed=32 + size, state_checksum=checksum(sd), z=z, shift=shift)
        with torch.no_grad():
            for wspace in (False, True):
                W = gan_load.StyleGAN2Wrapper(G, shift_in_w_space=wspace)
                tag = 'w' if wspace else 'z'
                img = W(z)
                img_s = W(z, shift)
                check('G%d %s-space plain' % (size, tag), o.generate(sd, z, None, size, wspace), img, 2e-5)
                check('G%d %s-space shifted' % (size, tag), o.generate(sd, z, shift, size, wspace), img_s, 2e-5)
                fx['img_' + tag] = img
                fx['img_shifted_' + tag] = img_s
            wlat = W.get_w(z)
            check('G%d get_w' % size, o.mapping(sd, z), wlat, 1e-5)
            img_lw = W(wlat, shift, latent_is_w=True)
            check('G%d latent_is_w' % size, o.generate(sd, wlat, shift, size, True, latent_is_w=True), img_lw, 2e-5)
            fx['w'] = wlat
        if size == 128:                               # keep the fixture small: strided sample + moments
            for k in list(fx):
                if k.startswith('img_'):
                    fx[k + '_mean'] = float(fx[k].double().mean())
                    fx[k + '_std'] = float(fx[k].double().std())
                    fx[k] = fx[k][:, :, ::4, ::4].clone()
            fx['stride'] = 4
        save('stylegan2_%d.pt' % size, fx)



def _sample_grad(g, limit=20000):
    """Whole tensor when small, else a fixed strided sample (fixtures stay small)."""
    flat = g.reshape(-1)
    if flat.numel() <= limit:
        return dict(stride=1, values=flat.clone())
    stride = flat.numel() // limit + 1
    return dict(stride=stride, values=flat[::stride].clone())


def pin_stylegan2_1024_step(model):
    """BASELINE config 3 at FULL size (StyleGAN2-1024, K=128, D=32, ResNet-18 R at 1024^2), batch 2: the reference
    modules' own loop body (lib/trainer.py:190-250) with injected draws, in Z space; plus a W-space forward pair.
    Gradients are additionally computed by the oracle in fp64: ReLU / leaky-ReLU / max-pool kinks make fp32 gradients of
    ANY two implementations of this graph differ at the 1e-3 level (millions of activations, a few within rounding of
    a kink), so every stored gradient carries the reference-fp32-vs-fp64 distance as its yardstick."""
    import oracle.stylegan2 as o
    import oracle.support_sets as o_ss
    import oracle.reconstructor as o_rec
    import oracle.step as o_step
    ss_ref = importlib.import_module('lib.support_sets')
    rec_ref = importlib.import_module('lib.reconstructor')
    gan_load = importlib.import_module('models.gan_load')
    print('[stylegan2-1024 paired step, config 3 at full size]')
    K, D, d, B, size = 128, 32, 512, 2, 1024
    seeds = (111, 112, 113)
    g_sd = o.init_state(size=size, generator=gen(seeds[0]))
    s_sd = o_ss.init_state(K, D, d, generator=gen(seeds[1]))
    r_sd = o_rec.init_state('ResNet', K, 3, generator=gen(seeds[2]))
    G = model.Generator(size, 512, 8)
    res = G.load_state_dict(g_sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith('.kernel') for k in res.missing_keys), res
    G.eval()
    W = gan_load.StyleGAN2Wrapper(G, shift_in_w_space=False)
    S = ss_ref.SupportSets(K, D, d, learn_alphas=False, learn_gammas=True, gamma=1.0 / d)
    S.load_state_dict(s_sd)
    R = rec_ref.Reconstructor('ResNet', K, 3)
    R.load_state_dict(r_sd)
    S.train(); R.train()
    for p in G.parameters():
        p.requires_grad_(False)          # the reference leaves these on and discards the result (SURVEY mismatch 4)
    g = gen(114)
    z = torch.randn(B, d, generator=g)
    idx = torch.tensor([5, 77])
    mag = torch.tensor([0.15, -0.12])
    mask = o_ss.one_hot(idx, K)
    img = W(z)
    shift = mag.reshape(-1, 1) * S(mask, z)
    img_shifted = W(z, shift)
    logits, pred = R(img, img_shifted)
    cls = torch.nn.CrossEntropyLoss()(logits, idx)
    reg = torch.mean(torch.abs(pred - mag))
    loss = 1.0 * cls + 0.25 * reg
    loss.backward()
    d_ss = S.SUPPORT_SETS.grad.clone()
    d_lg = S.LOGGAMMA.grad.clone()
    r_grads = {k: p.grad.clone() for k, p in R.named_parameters() if p.grad is not None}
    running = {k: v.clone() for k, v in R.state_dict().items() if 'running' in k}
    print('  reference fp32 step done: loss %.6f' % float(loss))
    # oracle fp32 (pins the restatement at full size)
    gen_fn, _ = o_step.make_generator('StyleGAN2', g_sd, size=size)
    o32 = o_step.paired_step(gen_fn, s_sd, r_sd, z, idx, mag, reconstructor_type='ResNet')
    check('img', o32['img'], img.detach(), 2e-5)
    check('img_shifted', o32['img_shifted'], img_shifted.detach(), 2e-5)
    check('logits', o32['logits'], logits.detach(), 1e-4)
    check('loss', o32['loss'], loss.detach(), 1e-5)
    # oracle fp64: the gradient yardstick
    to64 = lambda sd: {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    g64, s64, r64 = to64(g_sd), to64(s_sd), to64(r_sd)
    gen64, _ = o_step.make_generator('StyleGAN2', g64, size=size)
    o64 = o_step.paired_step(gen64, s64, r64, z.double(), idx, mag.double(), reconstructor_type='ResNet')
    print('  oracle fp64 step done: loss %.6f' % float(o64['loss']))
    # oracle fp32 under the kernels' arithmetic model (bf16 hi+lo operands in every dense forward conv): how far the
    # gradients of this graph move for a ~1e-5 forward perturbation (oracle/emulate.py)
    import oracle.emulate as o_emul
    with o_emul.split17_convs(o, o_rec):
        oem = o_step.paired_step(gen_fn, s_sd, r_sd, z, idx, mag, reconstructor_type='ResNet')
    print('  oracle fp32 with split17 operand rounding done: img_shifted rel err %.2e' % rel_err(oem['img_shifted'], o64['img_shifted']))
    rows = torch.unique(idx)
    yard_emul = {'SUPPORT_SETS': rel_err(oem['grads']['S']['SUPPORT_SETS'][rows], o64['grads']['S']['SUPPORT_SETS'][rows]),
                 'LOGGAMMA': rel_err(oem['grads']['S']['LOGGAMMA'][rows], o64['grads']['S']['LOGGAMMA'][rows]),
                 'img_shifted': rel_err(oem['img_shifted'], o64['img_shifted']), 'logits': rel_err(oem['logits'], o64['logits'])}
    yard = {'SUPPORT_SETS': rel_err(d_ss[rows], o64['grads']['S']['SUPPORT_SETS'][rows]),
            'LOGGAMMA': rel_err(d_lg[rows], o64['grads']['S']['LOGGAMMA'][rows]),
            'img_shifted': rel_err(img_shifted.detach(), o64['img_shifted']), 'logits': rel_err(logits.detach(), o64['logits'])}
    rg = {}
    for k, v in r_grads.items():
        e = rel_err(v, o64['grads']['R'][k])
        rg[k] = dict(norm=float(o64['grads']['R'][k].norm()), fp32_vs_fp64=e,
                     split17_vs_fp64=rel_err(oem['grads']['R'][k], o64['grads']['R'][k]), **_sample_grad(o64['grads']['R'][k].float()))
    worst = max(v['fp32_vs_fp64'] for v in rg.values())
    print('  reference-fp32 vs oracle-fp64: dSUPPORT_SETS %.2e, dLOGGAMMA %.2e, worst dR %.2e, logits %.2e'
          % (yard['SUPPORT_SETS'], yard['LOGGAMMA'], worst, yard['logits']))
    print('  split17 model  vs oracle-fp64: dSUPPORT_SETS %.2e, dLOGGAMMA %.2e, worst dR %.2e, logits %.2e'
          % (yard_emul['SUPPORT_SETS'], yard_emul['LOGGAMMA'], max(v['split17_vs_fp64'] for v in rg.values()), yard_emul['logits']))
    # W-space pair (forward only)
    Ww = gan_load.StyleGAN2Wrapper(G, shift_in_w_space=True)
    with torch.no_grad():
        wshift = 0.15 * F.normalize(torch.randn(B, d, generator=g), dim=1)
        img_w = Ww(z, wshift)
        check('img W-space', o.generate(g_sd, z, wshift, size, True), img_w, 2e-5)
    st = 8
    samp = lambda t: t.detach()[:, :, ::st, ::st].clone()
    save('stylegan2_1024_step.pt', dict(
        K=K, D=D, d=d, B=B, size=size, seeds=seeds, checksums=(checksum(g_sd), checksum(s_sd), checksum(r_sd)),
        z=z, idx=idx, mag=mag, wshift=wshift, stride=st,
        img=samp(img), img_shifted=samp(img_shifted), img_w=samp(img_w),
        img_mean=float(img.double().mean()), img_std=float(img.double().std()),
        img_shifted_std=float(img_shifted.double().std()), img_w_std=float(img_w.double().std()),
        img_row=img.detach()[:, :, 517, :].clone(), img_shifted_row=img_shifted.detach()[:, :, 517, :].clone(),
        shift=shift.detach(), logits=logits.detach(), pred=pred.detach(), cls=cls.detach(), reg=reg.detach(),
        loss=loss.detach(), rows=rows,
        logits64=o64['logits'].float(), loss64=float(o64['loss']),
        d_support_sets_rows=d_ss[rows].clone(), d_support_sets_rows64=o64['grads']['S']['SUPPORT_SETS'][rows].float(),
        d_loggamma=d_lg.clone(), d_loggamma64=o64['grads']['S']['LOGGAMMA'].float(),
        r_grads64=rg, yardstick=yard, yardstick_split17=yard_emul,
        running_sample={k: running[k] for k in ('features_extractor.bn1.running_mean', 'features_extractor.bn1.running_var',
                                                'features_extractor.layer4.1.bn2.running_var')}))

def pin_proggan():
    import oracle.proggan as o
    ref = importlib.import_module('models.ProgGAN.model')
    gan_load = importlib.import_module('models.gan_load')
    print('[proggan]')
    fx = {}
    for like in (True, False):
        sd = o.init_state(generator=gen(41 + like), pretrained_like=like)
        G = ref.Generator()
        G.load_state_dict(sd)
        W = gan_load.ProgGANWrapper(G).eval()
        g = gen(43)
        z = torch.randn(1, 512, generator=g)
        shift = 0.15 * F.normalize(torch.randn(1, 512, generator=g), dim=1)
        with torch.no_grad():
            img = W(z, shift)
            got = o.generate(sd, z, shift)
        check('ProgGAN-1024 pretrained_like=%s' % like, got, img, 2e-5)
        tag = 'like' if like else 'ctor'
        fx[tag] = dict(seed=41 + like, state_checksum=checksum(sd), img=img[:, :, ::16, ::16].clone(),
                       mean=float(img.double().mean()), std=float(img.double().std()))
    fx.update(z=z, shift=shift, stride=16)
    save('proggan_1024.pt', fx)


def pin_sngan():
    import oracle.sngan as o
    ref = importlib.import_module('models.SNGAN.sn_gen_resnet')
    dist = importlib.import_module('models.SNGAN.distribution')
    gan_load = importlib.import_module('models.gan_load')
    print('[sngan]')
    for model_name, ch, size in (('sn_resnet32', 1, 32), ('sn_resnet64', 3, 64)):
        sd = o.init_state(model_name, image_channels=ch, generator=gen(51 + size))
        G = ref.make_resnet_generator(ref.SN_RES_GEN_CONFIGS[model_name], img_size=size, channels=ch,
                                      distribution=dist.NormalDistribution(128))
        full = dict(sd)
        for k, v in list(sd.items()):                  # conv1/conv2 are registered twice (also as model.3/6)
            for a, b in (('.conv1.', '.model.3.'), ('.conv2.', '.model.6.')):
                if a in k:
                    full[k.replace(a, b)] = v
        res = G.model.load_state_dict(full, strict=False)
        assert not res.unexpected_keys and all('num_batches_tracked' in k for k in res.missing_keys), res
        W = gan_load.SNGANWrapper(G).eval()
        g = gen(52 + size)
        z = torch.randn(4, 128, generator=g)
        shift = 0.2 * F.normalize(torch.randn(4, 128, generator=g), dim=1)
        with torch.no_grad():
            img, img_s = W(z), W(z, shift)
        check('SNGAN %s' % model_name, o.generate(sd, z, None, model_name), img, 1e-5)
        check('SNGAN %s shifted' % model_name, o.generate(sd, z, shift, model_name), img_s, 1e-5)
        save('sngan_%d.pt' % size, dict(model=model_name, channels=ch, seed=51 + size, state_checksum=checksum(sd),
                                        z=z, shift=shift, img=img, img_shifted=img_s))


def pin_biggan():
    import oracle.biggan as o
    import json
    ref = importlib.import_module('models.BigGAN.BigGAN')
    utils = importlib.import_module('models.BigGAN.utils')
    print('[biggan]')
    with open(os.path.join(REF, 'models/BigGAN/generator_config.json')) as f:
        config = json.load(f)
    config['resolution'] = utils.imsize_dict[config['dataset']]
    config['n_classes'] = utils.nclass_dict[config['dataset']]
    config['G_activation'] = utils.activation_dict[config['G_nl']]
    config['D_activation'] = utils.activation_dict[config['D_nl']]
    config['skip_init'] = True
    config['no_optim'] = True
    G = ref.Generator(**config).eval()
    sd = o.init_state(128, generator=gen(61))
    G.load_state_dict(sd, strict=True)
    g = gen(62)
    z = torch.randn(2, G.dim_z, generator=g)
    shift = 0.2 * F.normalize(torch.randn(2, G.dim_z, generator=g), dim=1)
    classes = torch.tensor([239, 100])
    with torch.no_grad():
        img = G(z + shift, G.shared(classes))
    check('BigGAN-128', o.generate(sd, z, classes, shift), img, 2e-5)
    save('biggan_128.pt', dict(seed=61, state_checksum=checksum(sd), z=z, shift=shift, classes=classes,
                               img=img[:, :, ::2, ::2].clone(), stride=2, mean=float(img.double().mean()),
                               std=float(img.double().std())))


def pin_reconstructor():
    import oracle.reconstructor as o
    ref = importlib.import_module('lib.reconstructor')
    print('[reconstructor]')
    for rtype, ch, size, dim, B in (('ResNet', 3, 64, 16, 4), ('LeNet', 1, 32, 8, 4)):
        sd = o.init_state(rtype, dim, ch, generator=gen(71 + size))
        R = ref.Reconstructor(rtype, dim, ch)
        R.load_state_dict(sd, strict=True)
        R.train()
        g = gen(72 + size)
        x1 = torch.randn(B, ch, size, size, generator=g)
        x2 = x1 + 0.3 * torch.randn(B, ch, size, size, generator=g)
        x1.requires_grad_(True)
        x2.requires_grad_(True)
        logits, mag = R(x1, x2)
        idx = torch.randint(0, dim, (B,), generator=g)
        tgt = torch.rand(B, generator=g) * 0.2
        loss = F.cross_entropy(logits, idx) + 0.25 * (mag - tgt).abs().mean()
        loss.backward()
        # oracle, same thing under autograd
        leaf = dict(sd)
        keys = o.trainable_keys(sd)
        for k in keys:
            leaf[k] = sd[k].clone().requires_grad_(True)
        y1 = x1.detach().clone().requires_grad_(True)
        y2 = x2.detach().clone().requires_grad_(True)
        running = {}
        lo, mo = o.forward(leaf, y1, y2, rtype, True, running)
        lo_loss = F.cross_entropy(lo, idx) + 0.25 * (mo - tgt).abs().mean()
        lo_loss.backward()
        check('%s logits' % rtype, lo.detach(), logits.detach(), 1e-5)
        check('%s magnitudes' % rtype, mo.detach(), mag.detach(), 1e-5)
        check('%s dx1' % rtype, y1.grad, x1.grad, 1e-4)
        ref_params = dict(R.named_parameters())
        gnorms = {}
        for k in keys:
            assert ref_params[k].grad is not None, k
            check('%s d%s' % (rtype, k), leaf[k].grad, ref_params[k].grad, 2e-4)
            gnorms[k] = float(ref_params[k].grad.double().norm())
        for k in ref_params:
            if k not in keys:
                assert ref_params[k].grad is None, k                 # the dead fc gets no gradient
        ref_sd = R.state_dict()
        for k, v in running.items():
            check('%s %s' % (rtype, k), v, ref_sd[k], 1e-5)
        first = 'features_extractor.conv1.weight' if rtype == 'ResNet' else 'feature_extractor.0.weight'
        save('reconstructor_%s.pt' % rtype.lower(), dict(
            type=rtype, dim=dim, channels=ch, seed=71 + size, state_checksum=checksum(sd),
            x1=x1.detach(), x2=x2.detach(), idx=idx, tgt=tgt, logits=logits.detach(), mag=mag.detach(),
            loss=loss.detach(), dx1=x1.grad.clone(), dx2=x2.grad.clone(), grad_norms=gnorms,
            d_first_conv=ref_params[first].grad.clone(),
            d_head_w=ref_params['path_indices.weight' if rtype == 'ResNet' else 'path_indices.3.weight'].grad.clone(),
            running={k: ref_sd[k].clone() for k in running}))


def pin_step():
    """Config 1 (SNGAN-MNIST, K=32, D=16, LeNet, B=4): the trainer's loop body on reference modules
    with injected randoms, then one Adam step of both optimisers."""
    import oracle.support_sets as o_ss
    import oracle.sngan as o_sn
    import oracle.reconstructor as o_rec
    import oracle.step as o_step
    ss_ref = importlib.import_module('lib.support_sets')
    rec_ref = importlib.import_module('lib.reconstructor')
    sn_ref = importlib.import_module('models.SNGAN.sn_gen_resnet')
    dist = importlib.import_module('models.SNGAN.distribution')
    gan_load = importlib.import_module('models.gan_load')
    print('[paired step, config 1]')
    K, D, d, B = 32, 16, 128, 4
    g_sd = o_sn.init_state('sn_resnet32', 1, generator=gen(81))
    s_sd = o_ss.init_state(K, D, d, generator=gen(82))
    r_sd = o_rec.init_state('LeNet', K, 1, generator=gen(83))
    Gm = sn_ref.make_resnet_generator(sn_ref.SN_RES_GEN_CONFIGS['sn_resnet32'], img_size=32, channels=1,
                                      distribution=dist.NormalDistribution(128))
    full = dict(g_sd)
    for k, v in list(g_sd.items()):
        for a, b in (('.conv1.', '.model.3.'), ('.conv2.', '.model.6.')):
            if a in k:
                full[k.replace(a, b)] = v
    Gm.model.load_state_dict(full, strict=False)
    G = gan_load.SNGANWrapper(Gm).eval()
    S = ss_ref.SupportSets(K, D, d, learn_alphas=False, learn_gammas=True, gamma=1.0 / d)
    S.load_state_dict(s_sd)
    R = rec_ref.Reconstructor('LeNet', K, 1)
    R.load_state_dict(r_sd)
    S.train()
    R.train()
    s_opt = torch.optim.Adam(S.parameters(), lr=1e-4)
    r_opt = torch.optim.Adam(R.parameters(), lr=1e-4)
    g = gen(84)
    z = torch.randn(B, d, generator=g)
    idx = torch.randint(0, K, (B,), generator=g)
    mag = o_step.sample_shift_magnitudes(B, 0.15, 0.25, generator=g)
    mask = o_ss.one_hot(idx, K)
    # --- reference loop body, lib/trainer.py:190-254
    G.zero_grad(); S.zero_grad(); R.zero_grad()
    img = G(z)
    shift = mag.reshape(-1, 1) * S(mask, z)
    img_shifted = G(z, shift)
    logits, pred = R(img, img_shifted)
    cls = torch.nn.CrossEntropyLoss()(logits, idx)
    reg = torch.mean(torch.abs(pred - mag))
    loss = 1.0 * cls + 0.25 * reg
    loss.backward()
    d_ss = S.SUPPORT_SETS.grad.clone()
    d_lg = S.LOGGAMMA.grad.clone()
    r_grads = {k: p.grad.clone() for k, p in R.named_parameters() if p.grad is not None}
    s_opt.step(); r_opt.step()
    # --- oracle
    gen_fn, _ = o_step.make_generator('SNGAN', g_sd, model='sn_resnet32')
    res = o_step.paired_step(gen_fn, s_sd, r_sd, z, idx, mag, reconstructor_type='LeNet')
    check('img', res['img'], img.detach(), 1e-5)
    check('img_shifted', res['img_shifted'], img_shifted.detach(), 1e-5)
    check('shift', res['shift'], shift.detach(), 1e-6)
    check('logits', res['logits'], logits.detach(), 1e-5)
    check('loss', res['loss'], loss.detach(), 1e-6)
    check('dSUPPORT_SETS', res['grads']['S']['SUPPORT_SETS'], d_ss, 1e-4)
    check('dLOGGAMMA', res['grads']['S']['LOGGAMMA'], d_lg, 1e-4)
    for k, v in r_grads.items():
        check('dR ' + k, res['grads']['R'][k], v, 2e-4)
    # Adam: oracle update from the oracle grads must land on the reference's post-step parameters
    new_ss = s_sd['SUPPORT_SETS'].clone()
    m, v = torch.zeros_like(new_ss), torch.zeros_like(new_ss)
    o_step.adam_update(new_ss, res['grads']['S']['SUPPORT_SETS'], m, v, 1)
    touched = torch.unique(idx)
    check('Adam SUPPORT_SETS delta', (new_ss - s_sd['SUPPORT_SETS'])[touched],
          (S.SUPPORT_SETS.detach() - s_sd['SUPPORT_SETS'])[touched], 1e-3)
    save('step_c1.pt', dict(K=K, D=D, d=d, seeds=(81, 82, 83), z=z, idx=idx, mag=mag,
                            checksums=(checksum(g_sd), checksum(s_sd), checksum(r_sd)),
                            img=img.detach(), img_shifted=img_shifted.detach(), shift=shift.detach(),
                            logits=logits.detach(), pred=pred.detach(), cls=cls.detach(), reg=reg.detach(),
                            loss=loss.detach(), rows=touched, d_support_sets_rows=d_ss[touched],
                            d_loggamma=d_lg, r_grad_norms={k: float(v.double().norm()) for k, v in r_grads.items()},
                            new_support_sets_rows=S.SUPPORT_SETS.detach()[touched].clone(),
                            new_head_bias=R.path_indices[3].bias.detach().clone()))


def pin_traversal():
    import oracle.support_sets as o_ss
    import oracle.step as o_step
    ss_ref = importlib.import_module('lib.support_sets')
    print('[traversal chain]')
    K, D, d = 16, 8, 64
    sd = o_ss.init_state(K, D, d, generator=gen(91))
    S = ss_ref.SupportSets(K, D, d, learn_alphas=False, learn_gammas=True, gamma=1.0 / d)
    S.load_state_dict(sd)
    z0 = torch.randn(1, d, generator=gen(92))
    eps, steps, path = 0.2, 8, 5
    # reference loop, traverse_latent_space.py:369-438 (Z space)
    codes, shifts = [z0], [torch.zeros_like(z0)]
    z = z0.clone()
    for _ in range(steps):
        mask = torch.zeros(1, K); mask[0, path] = 1.0
        with torch.no_grad():
            sh = eps * S(mask, z)
        z = z + sh
        shifts.append(sh); codes.append(z)
    z = z0.clone()
    for _ in range(steps):
        mask = torch.zeros(1, K); mask[0, path] = 1.0
        with torch.no_grad():
            sh = -eps * S(mask, z)
        z = z + sh
        shifts = [sh] + shifts; codes = [z] + codes
    codes, shifts = torch.cat(codes), torch.cat(shifts)
    c2, s2 = o_step.traverse_chain(sd, z0, path, eps, steps)
    check('codes', c2, codes, 1e-6)
    check('shifts', s2, shifts, 1e-5)
    save('traversal_chain.pt', dict(K=K, D=D, d=d, seed=91, state_checksum=checksum(sd), z0=z0, eps=eps,
                                    steps=steps, path=path, codes=codes, shifts=shifts))


def pin_trainer_loop():
    """The UNMODIFIED reference driver (lib/trainer.py:24-319) on config 1 for a few iterations under a fixed global
    seed, on the CPU: pins the host draw order, the statistics written to stats.json, the checkpoint layout and the
    parameters after several Adam steps.  tests/test_trainer_gpu.py runs warpedganspace_b200.Trainer on the same
    seed and compares."""
    import argparse
    import json
    import tempfile
    import oracle.support_sets as o_ss
    import oracle.sngan as o_sn
    import oracle.reconstructor as o_rec
    lib = importlib.import_module('lib')
    sn_ref = importlib.import_module('models.SNGAN.sn_gen_resnet')
    dist = importlib.import_module('models.SNGAN.distribution')
    gan_load = importlib.import_module('models.gan_load')
    print('[reference Trainer.train, config 1]')
    K, D, d, B, iters = 32, 16, 128, 4, 6
    g_sd = o_sn.init_state('sn_resnet32', 1, generator=gen(101))
    s_sd = o_ss.init_state(K, D, d, generator=gen(102))
    r_sd = o_rec.init_state('LeNet', K, 1, generator=gen(103))
    Gm = sn_ref.make_resnet_generator(sn_ref.SN_RES_GEN_CONFIGS['sn_resnet32'], img_size=32, channels=1,
                                      distribution=dist.NormalDistribution(128))
    full = dict(g_sd)
    for k, v in list(g_sd.items()):
        for a, b in (('.conv1.', '.model.3.'), ('.conv2.', '.model.6.')):
            if a in k:
                full[k.replace(a, b)] = v
    Gm.model.load_state_dict(full, strict=False)
    G = gan_load.SNGANWrapper(Gm)
    S = lib.SupportSets(K, D, d, learn_alphas=False, learn_gammas=True, gamma=1.0 / d)
    S.load_state_dict(s_sd)
    R = lib.Reconstructor('LeNet', K, 1)
    R.load_state_dict(r_sd)
    params = argparse.Namespace(
        gan_type='SNGAN_MNIST', z_truncation=None, biggan_target_classes=None, stylegan2_resolution=1024,
        shift_in_w_space=False, num_support_sets=K, num_support_dipoles=D, learn_alphas=False, learn_gammas=True,
        gamma=None, support_set_lr=1e-4, reconstructor_type='LeNet', min_shift_magnitude=0.15,
        max_shift_magnitude=0.25, reconstructor_lr=1e-4, max_iter=iters, batch_size=B, lambda_cls=1.0,
        lambda_reg=0.25, log_freq=2, ckp_freq=3, tensorboard=False, cuda=False)
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix='wgs_ref_trainer_')
    os.chdir(work)
    try:
        exp_dir = lib.create_exp_dir(params)
        trn = lib.Trainer(params=params, exp_dir=exp_dir, use_cuda=False, multi_gpu=False)
        torch.manual_seed(2024)
        trn.train(generator=G, support_sets=S, reconstructor=R)
        wip = os.path.join(work, 'experiments', 'wip', exp_dir)
        done = os.path.join(work, 'experiments', 'complete', exp_dir)
        with open(os.path.join(wip, 'stats.json')) as f:
            stats = json.load(f)
        ckpt = torch.load(os.path.join(wip, 'models', 'checkpoint.pt'))
        files_wip = sorted(os.path.relpath(os.path.join(r, f), wip) for r, _, fs in os.walk(wip) for f in fs)
        files_done = sorted(os.path.relpath(os.path.join(r, f), done) for r, _, fs in os.walk(done) for f in fs)
        final_s = torch.load(os.path.join(wip, 'models', 'support_sets.pt'))
        final_r = torch.load(os.path.join(wip, 'models', 'reconstructor.pt'))
        init_s = torch.load(os.path.join(wip, 'models', 'support_sets_init.pt'))
    finally:
        os.chdir(cwd)
    assert torch.equal(init_s['SUPPORT_SETS'], s_sd['SUPPORT_SETS'])
    moved = (final_s['SUPPORT_SETS'] - s_sd['SUPPORT_SETS']).abs().amax(dim=1).nonzero().flatten()
    print('  exp_dir %s; stats keys %s; %d support-set rows moved' % (exp_dir, sorted(stats), moved.numel()))
    save('trainer_c1.pt', dict(
        K=K, D=D, d=d, B=B, iters=iters, seeds=(101, 102, 103), train_seed=2024, params=vars(params),
        checksums=(checksum(g_sd), checksum(s_sd), checksum(r_sd)), exp_dir=exp_dir, stats=stats,
        files_wip=files_wip, files_complete=files_done, checkpoint_iter=ckpt['iter'],
        checkpoint_keys={k: sorted(v.keys()) if isinstance(v, dict) else None for k, v in ckpt.items()},
        moved_rows=moved, final_support_sets_rows=final_s['SUPPORT_SETS'][moved].clone(),
        final_loggamma=final_s['LOGGAMMA'].clone(),
        final_head_bias=final_r['path_indices.3.bias'].clone(),
        final_conv0_weight=final_r['feature_extractor.0.weight'].clone()))


def pin_latent_pool():
    """Known-answer test for the latent-pool format (sample_gan.py:156-179): the 58 pools shipped with the reference,
    directory name = sha1 of the [1, dim_z] fp32 tensor bytes."""
    import glob
    print('[latent pools shipped with the reference]')
    entries = {}
    for path in sorted(glob.glob(os.path.join(REF, 'experiments', 'latent_codes', '*', '*', '*', 'latent_code.pt'))):
        parts = path.split(os.sep)
        entries['/'.join(parts[-4:-1])] = torch.load(path, map_location='cpu')
    assert len(entries) >= 50, len(entries)
    print('  %d latent codes, dims %s' % (len(entries), sorted({tuple(v.shape) for v in entries.values()})))
    save('latent_pool.pt', entries)


def pin_eval_nets():
    """The two ResNet-based attribute predictors: oracle.eval_nets against torchvision's resnet34 and the
    reference's own Hopenet class, seeded weights, randomised BatchNorm statistics, one 224 x 224 crop pair."""
    import torchvision
    import oracle.eval_nets as o_en
    print('[attribute predictors: FairFace resnet34 / Hopenet resnet50]')
    hop = importlib.import_module('lib.evaluation.hopenet.hopenet')
    x = torch.randn(2, 3, 224, 224, generator=gen(811))
    fx = {'seed_x': 811}
    sd = o_en.init_state('basic', {'fc': 18}, gen(812))
    ref = torchvision.models.resnet34()
    ref.fc = torch.nn.Linear(ref.fc.in_features, 18)                       # traverse_attribute_space.py:179
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    with torch.no_grad():
        want = ref(x)
        got = o_en.resnet_forward(sd, x, 'basic', ('fc',))[0]
    check('FairFace resnet34 logits', got, want, 1e-5)
    fx['fairface'] = dict(seed=812, checksum=checksum(sd), out=want)
    heads = {'fc_yaw': 66, 'fc_pitch': 66, 'fc_roll': 66}
    sd = o_en.init_state('bottleneck', heads, gen(813))
    ref = hop.Hopenet(torchvision.models.resnet.Bottleneck, [3, 4, 6, 3], 66)    # traverse_attribute_space.py:186
    missing = ref.load_state_dict(sd, strict=False)
    assert set(missing.missing_keys) == {'fc_finetune.weight', 'fc_finetune.bias'} and not missing.unexpected_keys
    ref.eval()
    with torch.no_grad():
        want = ref(x)
        got = o_en.resnet_forward(sd, x, 'bottleneck', tuple(heads))
    for name, g, w in zip(heads, got, want):
        check('Hopenet %s' % name, g, w, 1e-5)
    fx['hopenet'] = dict(seed=813, checksum=checksum(sd), out=[w.clone() for w in want])
    cel = importlib.import_module('lib.evaluation.celeba_attributes.celeba_attr_predictor')
    sd = o_en.init_celeba_state(gen(814))
    ref = cel.ResNet(cel.Bottleneck, [3, 4, 6, 3],                               # celeba_attr_predictor.py:185-186, no download
                     attr_file=os.path.join(REF, 'lib', 'evaluation', 'celeba_attributes', 'attributes_5.json'))
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    with torch.no_grad():
        want = ref(x)
        got = o_en.celeba_forward(sd, x)
    assert list(want) == list(got)
    for name in want:
        check('CelebA predictor %s' % name, got[name], want[name], 1e-5)
    fx['celeba'] = dict(seed=814, checksum=checksum(sd), out={k: v.clone() for k, v in want.items()})
    # S3FD face detector: network against the reference class, post-processing against the reference's batch_detect + nms
    import numpy as np
    net_mod = importlib.import_module('lib.evaluation.sfd.net_s3fd')
    det_mod = importlib.import_module('lib.evaluation.sfd.detect')
    sd = o_en.init_s3fd_state(gen(815))
    ref = net_mod.s3fd()
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    xf = 255.0 * torch.rand(2, 3, 128, 128, generator=gen(816))
    with torch.no_grad():
        want = ref(xf)
        got = o_en.s3fd_forward(sd, xf)
    for i, (g_, w_) in enumerate(zip(got, want)):
        check('S3FD output %d %s' % (i, tuple(w_.shape)), g_, w_, 1e-5)
    boxes_ref = det_mod.batch_detect(ref, xf, 'cpu')                      # [B, M, 5] (positions gathered over the batch)
    dets_ref = []
    for i in range(boxes_ref.shape[0]):                                   # sfd_detector.py:29-36
        bl = boxes_ref[i].astype(np.float64)
        keep = det_mod.nms(bl, 0.3)
        dets_ref.append(np.array([b for b in bl[keep, :] if b[-1] > 0.5]).reshape(-1, 5))
    dets = o_en.sfd_detect_from_batch(got)
    for j in range(2):
        print('  image %d: %d candidate boxes in the reference list, %d detections' % (j, boxes_ref.shape[1], len(dets_ref[j])))
        assert dets[j].shape == dets_ref[j].shape and len(dets[j]) > 0, (dets[j].shape, dets_ref[j].shape)
        a = dets[j][np.lexsort(dets[j].T)]
        b = dets_ref[j][np.lexsort(dets_ref[j].T)]
        assert np.allclose(a, b, rtol=1e-4, atol=1e-3), float(np.abs(a - b).max())
    fx['s3fd'] = dict(seed=815, seed_x=816, checksum=checksum(sd), out=[w_[:, :, ::2, ::2].clone() for w_ in want],
                      detections=[torch.from_numpy(d) for d in dets_ref])
    # ArcFace identity comparator: the reference IDComparator (its constructor reads the checkpoint: torch.load is patched
    # to hand it the seeded state) on two 256 x 256 frames
    from unittest import mock
    arc = importlib.import_module('lib.evaluation.archface.arcface')
    sd = o_en.init_arcface_state(gen(817))
    with mock.patch('torch.load', return_value=sd):
        idc = arc.IDComparator()
    idc.eval()
    assert set(idc.backbone.state_dict()) == set(sd)
    xa = torch.rand(3, 3, 256, 256, generator=gen(818)) * 2 - 1
    xa[2] = 0.7 * xa[0] + 0.3 * xa[1]                                     # a related frame: similarity away from 0
    with torch.no_grad():
        want_f = idc.extract_feats(xa)
        got_f = o_en.arcface_extract_feats(sd, xa)
        want_s = torch.stack([idc(xa[0:1], xa[t: t + 1]) for t in range(3)])
        got_s = torch.stack([o_en.id_similarity(sd, xa[0:1], xa[t: t + 1]) for t in range(3)])
    check('ArcFace embeddings', got_f, want_f, 1e-5)
    check('ArcFace similarity to frame 0 %s' % [round(float(v), 4) for v in want_s], got_s, want_s, 1e-5)
    fx['arcface'] = dict(seed=817, seed_x=818, checksum=checksum(sd), feats=want_f.clone(), sim=want_s.clone())
    # Action-unit detector: the reference AUdetector (same patch for its checkpoint read) on two 256 x 256 crops
    aud = importlib.import_module('lib.evaluation.au_detector.AU_detector')
    sd = o_en.init_au_state(gen(819))
    with mock.patch('torch.load', return_value={'state_dict': sd}):
        ref = aud.AUdetector(use_cuda=False)
    assert set(ref.AUdetector.FAN.state_dict()) == set(sd)
    xu = 255.0 * torch.rand(2, 3, 256, 256, generator=gen(820))
    with torch.no_grad():
        want_i = ref.detect_AU(xu)
        want_h = ref.AUdetector.forward_FAN((xu - xu.min()) / (xu.max() - xu.min()))
        got_i = o_en.detect_au(sd, xu)
        got_h = o_en.au_heatmaps(sd, (xu - xu.min()) / (xu.max() - xu.min()))
    check('AU heat maps %s' % (tuple(want_h.shape),), got_h, want_h, 1e-5)
    check('AU intensities', got_i, want_i, 1e-5)
    fx['au'] = dict(seed=819, seed_x=820, checksum=checksum(sd), heat=want_h[:, :, ::4, ::4].clone(), intensities=want_i.clone())
    save('eval_nets.pt', fx)


def main():
    torch.set_num_threads(os.cpu_count())
    os.chdir('/tmp')
    model, up_mod = import_reference()
    only = sys.argv[1:]
    if only:                                     # e.g. python -m oracle.gen_golden stylegan2_1024_step
        for name in only:
            fn = globals()['pin_' + name]
            fn(*((model, up_mod) if name == 'stylegan2' else (model,) if name == 'stylegan2_1024_step' else ()))
        return
    pin_support_sets()
    pin_stylegan2(model, up_mod)
    pin_sngan()
    pin_reconstructor()
    pin_step()
    pin_traversal()
    pin_trainer_loop()
    pin_latent_pool()
    pin_proggan()
    pin_biggan()
    pin_stylegan2_1024_step(model)
    pin_eval_nets()
    print('oracle pinned against the reference; fixtures in', OUT)


if __name__ == '__main__':
    main()
