"""Layer-by-layer comparison of the CUDA StyleGAN2 forward against the oracle (debug aid)."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle.stylegan2 as o
from warpedganspace_b200.stylegan2 import Generator, synthesis, styles_and_demod

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))

size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator().manual_seed(1)
sd = o.init_state(size=size, generator=g)
G = Generator(size, 512, 8); G.load_state_dict(sd, strict=False); G.cuda().eval()
z = torch.randn(2, 512, generator=g)
w = o.mapping(sd, z)
wc = G.get_latent(z.cuda())
print('w', rel(wc, w))
P = G.plan()
s_all, demod = styles_and_demod(G, wc)
names = ['conv1'] + ['convs.%d' % i for i in range(G.num_layers - 1)]
for li, (e, name) in enumerate(zip(P['styled'], names)):
    s_ref = o.equal_linear(w, sd[name + '.conv.modulation.weight'], sd[name + '.conv.modulation.bias'])
    print(name, 'style', rel(s_all[:, e['s_off']:e['s_off'] + e['ci']], s_ref), end=' ')
    wt = sd[name + '.conv.weight']
    wm = (1 / math.sqrt(e['ci'] * 9)) * wt * s_ref.view(2, 1, -1, 1, 1)
    d_ref = torch.rsqrt(wm.pow(2).sum([2, 3, 4]) + 1e-8)
    print('demod', rel(demod[li], d_ref))
tape = {}
img = synthesis(G, wc, tape)
x = sd['input.input'].repeat(2, 1, 1, 1)
x = o.styled_conv(sd, 'conv1', x, w, sd['noises.noise_0'])
print('conv1 act', rel(tape['acts'][1].permute(0, 3, 1, 2), x))
skip = o.to_rgb(sd, 'to_rgb1', x, w)

for i in range(G.log_size - 2):
    x = o.styled_conv(sd, 'convs.%d' % (2 * i), x, w, sd['noises.noise_%d' % (2 * i + 1)], upsample=True)
    print('convs.%d act' % (2 * i), rel(tape['acts'][2 * i + 2].permute(0, 3, 1, 2), x))
    x = o.styled_conv(sd, 'convs.%d' % (2 * i + 1), x, w, sd['noises.noise_%d' % (2 * i + 2)])
    print('convs.%d act' % (2 * i + 1), rel(tape['acts'][2 * i + 3].permute(0, 3, 1, 2), x))
    skip = o.to_rgb(sd, 'to_rgbs.%d' % i, x, w, skip)
print('image', rel(img.permute(0, 3, 1, 2), skip))
