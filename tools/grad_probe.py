"""GPU probe behind tests/test_gradients_gpu.py: whole-graph gradients of the paired step (small StyleGAN2 + ResNet R) from
libwgs_b200 vs the fp64 oracle, next to the fp32 oracle's own distance from fp64, per seed and per tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle.support_sets as o_ss, oracle.stylegan2 as o_sg2, oracle.reconstructor as o_rec, oracle.step as o_step

gen = lambda s: torch.Generator().manual_seed(s)
rel = lambda a, b: float((a.detach().double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
to64 = lambda sd: {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
size = int(sys.argv[3]) if len(sys.argv) > 3 else 32
ch = {4: 64, 8: 64, 16: 32, 32: 32, 64: 32, 128: 32}
K, D, B = 16, 4, 4
torch.backends.cudnn.allow_tf32 = False


def run(g_sd, s_sd, r_sd, z, idx, mag):
    gen_fn, _ = o_step.make_generator('StyleGAN2', g_sd, size=size)
    return o_step.paired_step(gen_fn, s_sd, r_sd, z, idx, mag, reconstructor_type='ResNet')


for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    from warpedganspace_b200 import SupportSets
    from warpedganspace_b200.stylegan2 import Generator
    from warpedganspace_b200.gan_load import StyleGAN2Wrapper
    from warpedganspace_b200.reconstructor import Reconstructor
    from warpedganspace_b200.trainer import PairedTrainer
    g_sd = o_sg2.init_state(size=size, generator=gen(seed), channels=ch)
    s_sd = o_ss.init_state(K, D, 512, generator=gen(seed + 1))
    r_sd = o_rec.init_state('ResNet', K, 3, generator=gen(seed + 2))
    g = gen(seed + 3)
    z = torch.randn(B, 512, generator=g)
    idx = torch.randint(0, K, (B,), generator=g)
    mag = o_step.sample_shift_magnitudes(B, 0.1, 0.2, generator=g)
    rows = torch.unique(idx)
    a32 = run(g_sd, s_sd, r_sd, z, idx, mag)
    a64 = run(to64(g_sd), to64(s_sd), to64(r_sd), z.double(), idx, mag.double())
    G = Generator(size, 512, 8, channels=ch); G.load_state_dict(g_sd, strict=False)
    S = SupportSets(K, D, 512, learn_gammas=True, gamma=1.0 / 512); S.load_state_dict(s_sd)
    R = Reconstructor('ResNet', K, 3); R.load_state_dict(r_sd)
    T = PairedTrainer(StyleGAN2Wrapper(G, False).cuda(), S.cuda(), R.cuda())
    got = T.forward_backward(z.cuda(), idx.cuda(), mag.cuda())
    params = dict(R.named_parameters())
    ours = {'SS': rel(S.SUPPORT_SETS.grad[rows.cuda()], a64['grads']['S']['SUPPORT_SETS'][rows])}
    ref = {'SS': rel(a32['grads']['S']['SUPPORT_SETS'][rows], a64['grads']['S']['SUPPORT_SETS'][rows])}
    for k in a64['grads']['R']:
        ours[k] = rel(params[k].grad, a64['grads']['R'][k])
        ref[k] = rel(a32['grads']['R'][k], a64['grads']['R'][k])
    wk = max(ours, key=ours.get)
    print('seed %d size %d: img %.1e logits %.1e | ours-vs-fp64: SS %.2e worst %.2e (%s) | fp32-oracle-vs-fp64: SS %.2e worst %.2e'
          % (seed, size, rel(got['img_shifted'], a64['img_shifted']), rel(got['logits'], a64['logits']), ours['SS'], ours[wk], wk,
             ref['SS'], max(ref.values())), flush=True)
