import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import oracle.stylegan2 as o
from warpedganspace_b200.stylegan2 import Generator
from warpedganspace_b200 import conv as C
torch.backends.cudnn.allow_tf32 = False

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))

# (b) strided conv, small spatial, several images per tile
g = torch.Generator().manual_seed(0)
for (N, Ci, H, Co, k, s, p) in [(3, 64, 9, 64, 3, 2, 0), (3, 32, 17, 32, 3, 2, 0), (3, 32, 33, 32, 3, 2, 0), (3, 32, 65, 32, 3, 2, 0), (2, 64, 5, 32, 3, 2, 0)]:
    x = torch.randn(N, Ci, H, H, generator=g).cuda(); w = torch.randn(Co, Ci, k, k, generator=g).cuda() / (Ci * k * k) ** .5
    want = F.conv2d(x, w, stride=s, padding=p).permute(0, 2, 3, 1)
    got = C.conv2d(C.pack_split32(x.permute(0, 2, 3, 1).contiguous()), C.pack_weights(w), k, k, stride=s, padding=p)
    print('strided conv', (N, Ci, H, Co), rel(got, want), [rel(got[i], want[i]) for i in range(N)])

ch = {4: 64, 8: 64, 16: 32, 32: 32, 64: 32}
size = 64
sd = o.init_state(size=size, generator=torch.Generator().manual_seed(11), channels=ch)
G = Generator(size, 512, 8, channels=ch); G.load_state_dict(sd, strict=False); G.cuda().eval()
g = torch.Generator().manual_seed(12)
z = torch.randn(3, 512, generator=g)
cot = torch.randn(3, 3, size, size, generator=g)
w = o.mapping(sd, z)
wo = w.clone().requires_grad_(True)
img_o = o.synthesis(sd, wo, size)
(img_o * cot).sum().backward()
def run(idx):
    wc = w[idx].cuda().requires_grad_(True)
    img = G([wc], input_is_latent=True)[0]
    (img * cot[idx].cuda()).sum().backward()
    return wc.grad.cpu()
full = run([0, 1, 2])
print('full batch per-sample err', [rel(full[i], wo.grad[i]) for i in range(3)])
for i in range(3):
    single = run([i])
    print('single', i, rel(single[0], wo.grad[i]))
perm = run([2, 0, 1])
print('perm batch per-sample err', [rel(perm[j], wo.grad[i]) for j, i in enumerate([2, 0, 1])])
