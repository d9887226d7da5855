"""GPU parity of the op-level boundary (warpedganspace_b200.op = the reference's models/StyleGAN2/op package):
``upfirdn2d``, ``fused_leaky_relu`` / ``FusedLeakyReLU`` and the two raw extension functions, against the fixture written
from the reference's own ``upfirdn2d_native`` (op/upfirdn2d.py:152-186) and against the oracle under autograd."""
import pytest
import torch

import oracle.stylegan2 as o_sg2

pytestmark = pytest.mark.gpu


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def test_ops_match_reference_fixture(golden):
    from warpedganspace_b200.op import upfirdn2d, fused_leaky_relu
    fx = golden('stylegan2_ops.pt')
    x, k = fx['x'].cuda(), fx['kernel'].cuda()
    for name, c in fx['cases'].items():
        got = upfirdn2d(x, k, **c['kw'])
        assert got.shape == c['out'].shape, name
        assert rel(got, c['out']) < 1e-6, name
    assert rel(fused_leaky_relu(x, fx['bias'].cuda()), fx['lrelu']) < 1e-6


@pytest.mark.parametrize('shape,ksize,up,down,pad', [
    ((2, 5, 17, 13), (4, 4), 1, 1, (1, 1)),        # Blur after an up-conv (model.py:160-165, kernel mode 1)
    ((2, 3, 16, 16), (4, 4), 2, 1, (2, 1)),        # ToRGB skip Upsample (model.py:29-45, mode 3)
    ((1, 4, 32, 20), (4, 4), 1, 2, (1, 1)),        # Downsample / the adjoint of the skip up-sample (mode 5)
    ((3, 2, 9, 11), (3, 5), 3, 2, (2, 4)),         # no reference tile mode exists for this one (App. B.11)
    ((1, 1, 8, 8), (2, 2), 1, 1, (-1, 2)),         # negative pad = crop
    ((2, 32, 64, 64), (4, 4), 2, 1, (2, 1)),
])
def test_upfirdn2d_forward_and_gradient(shape, ksize, up, down, pad):
    from warpedganspace_b200.op import upfirdn2d
    g = gen(sum(shape) + up + down)
    x = torch.randn(*shape, generator=g)
    k = torch.randn(*ksize, generator=g)
    xo = x.clone().requires_grad_(True)
    want = o_sg2.upfirdn2d(xo, k, up=up, down=down, pad=pad)
    cot = torch.randn(want.shape, generator=g)
    (want * cot).sum().backward()
    xc = x.cuda().requires_grad_(True)
    got = upfirdn2d(xc, k.cuda(), up=up, down=down, pad=pad)
    assert got.shape == want.shape
    assert rel(got, want) < 1e-6
    (got * cot.cuda()).sum().backward()
    assert rel(xc.grad, xo.grad) < 1e-6


@pytest.mark.parametrize('shape', [(3, 7), (2, 5, 9), (2, 6, 8, 12), (1, 32, 33, 31)])
def test_fused_leaky_relu_forward_and_gradients(shape):
    from warpedganspace_b200.op import fused_leaky_relu, FusedLeakyReLU
    g = gen(sum(shape))
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    xo, bo = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    want = o_sg2.fused_leaky_relu(xo, bo, 0.1, 1.7)
    cot = torch.randn(want.shape, generator=g)
    (want * cot).sum().backward()
    xc, bc = x.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    got = fused_leaky_relu(xc, bc, 0.1, 1.7)
    assert rel(got, want) < 1e-6
    (got * cot.cuda()).sum().backward()
    assert rel(xc.grad, xo.grad) < 1e-6 and rel(bc.grad, bo.grad) < 1e-5
    m = FusedLeakyReLU(shape[1]).cuda()
    assert set(dict(m.named_parameters())) == {'bias'} and m.negative_slope == 0.2 and abs(m.scale - 2 ** 0.5) < 1e-12
    with torch.no_grad():
        m.bias.copy_(b)
    assert rel(m(x.cuda()), o_sg2.fused_leaky_relu(x, b)) < 1e-6


def test_raw_extension_functions_and_errors():
    """The pybind-level signatures (op/fused_bias_act.cpp:11-20, op/upfirdn2d.cpp:12-22) and their error behaviour."""
    from warpedganspace_b200.op import fused_bias_act, upfirdn2d_op
    g = gen(5)
    x = torch.randn(2, 4, 6, 6, generator=g).cuda()
    b = torch.randn(4, generator=g).cuda()
    empty = x.new_empty(0)
    lin = fused_bias_act(x, b, empty, 1, 0, 0.2, 3.0)                          # act 1 = linear
    assert rel(lin, (x + b.view(1, -1, 1, 1)) * 3.0) < 1e-6
    out = fused_bias_act(x, b, empty, 3, 0, 0.2, 2 ** 0.5)
    d1 = fused_bias_act(x, empty, out, 3, 1, 0.2, 2 ** 0.5)                    # derivative selected by the sign of `out`
    assert rel(d1, x * torch.where(out > 0, 1.0, 0.2) * 2 ** 0.5) < 1e-6
    assert float(fused_bias_act(x, empty, out, 3, 2, 0.2, 1.0).abs().max()) == 0.0
    k = torch.ones(2, 2).cuda()
    y = upfirdn2d_op(x.reshape(-1, 6, 6, 1), k, 1, 1, 1, 1, 0, 0, 0, 0)
    assert tuple(y.shape) == (8, 5, 5, 1)
    with pytest.raises(RuntimeError):
        fused_bias_act(x.cpu(), b.cpu(), empty.cpu(), 3, 0, 0.2, 1.0)         # the reference CHECK_CUDAs as well
    with pytest.raises(RuntimeError):
        fused_bias_act(x, b, empty, 7, 0, 0.2, 1.0)
