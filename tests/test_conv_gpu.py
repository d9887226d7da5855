"""GPU: the tcgen05 tap-list convolution (through the C ABI) against fp32 torch convolutions.
Tolerance: 2e-5 relative L2 (3-term bf16 split, fp32 accumulation); the north-star bar is 1e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.fixture(autouse=True)
def _fp32_convs():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def test_pack_split32_roundtrip():
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(0)
    for C in (32, 64, 6, 3, 96, 17):
        x = torch.randn(5, 7, C, generator=g).cuda() * 3
        s = conv.pack_split32(x)
        assert s.shape == (5, 7, (C + 31) // 32, 64)
        hi = s[..., :32].float().reshape(5, 7, -1)[..., :C]
        lo = s[..., 32:].float().reshape(5, 7, -1)[..., :C]
        assert rel(hi + lo, x) < 1e-5
        assert float((hi + lo - x).abs().max()) <= 2 ** -16 * float(x.abs().max())
        pad = s.float().reshape(5, 7, -1, 2, 32).permute(0, 1, 3, 2, 4).reshape(5, 7, 2, -1)[..., C:]
        assert float(pad.abs().max()) == 0.0 if pad.numel() else True
    # per-group channel scale
    x = torch.randn(4, 6, 40, generator=g).cuda()
    sc = torch.randn(4, 40, generator=g).cuda()
    s = conv.pack_split32(x, scale=sc, rows_per_group=6)
    got = (s[..., :32].float() + s[..., 32:].float()).reshape(4, 6, -1)[..., :40]
    assert rel(got, x * sc[:, None, :]) < 1e-5


CASES = [
    # N, Ci, H, W, Co, k, stride, pad
    (2, 32, 16, 16, 32, 3, 1, 1),
    (1, 64, 32, 32, 48, 3, 1, 1),
    (8, 512, 4, 4, 512, 3, 1, 1),
    (3, 128, 8, 8, 256, 3, 1, 1),
    (2, 96, 8, 8, 3, 1, 1, 0),
    (1, 32, 20, 24, 64, 3, 2, 1),
    (2, 6, 32, 32, 64, 7, 2, 3),
    (1, 512, 16, 16, 512, 3, 1, 1),
    (2, 64, 16, 16, 128, 1, 2, 0),
    (1, 16, 64, 64, 16, 3, 1, 1),
    (1, 32, 128, 128, 32, 3, 1, 1),
    (5, 64, 5, 9, 32, 3, 1, 1),
]


@pytest.mark.parametrize('case', CASES)
def test_conv2d_matches_torch(case):
    from warpedganspace_b200 import conv
    N, Ci, H, W, Co, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = torch.randn(Co, Ci, k, k, generator=g).cuda() / (Ci * k * k) ** 0.5
    want = F.conv2d(x, w, stride=stride, padding=pad)
    got = conv.conv2d(conv.pack_split32(nhwc(x)), conv.pack_weights(w), k, k, stride=stride, padding=pad)
    torch.cuda.synchronize()
    assert got.shape == nhwc(want).shape
    assert rel(got, nhwc(want)) < 2e-5


def test_epilogue_alpha_beta_act_accumulate():
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(5)
    N, Ci, H, W, Co = 3, 64, 16, 16, 80
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = torch.randn(Co, Ci, 3, 3, generator=g).cuda() / 24
    alpha = torch.rand(N, Co, generator=g).cuda() + 0.5
    beta = torch.randn(Co, generator=g).cuda()
    base = torch.randn(N, H, W, Co, generator=g).cuda()
    xs, ws = conv.pack_split32(nhwc(x)), conv.pack_weights(w)
    y = F.conv2d(x, w, padding=1)
    want = F.leaky_relu(y * alpha[:, :, None, None] + beta[None, :, None, None], 0.2)
    got = conv.conv2d(xs, ws, 3, 3, padding=1, alpha=alpha, beta=beta, act=2)
    assert rel(got, nhwc(want)) < 2e-5
    out = base.clone()
    conv.conv2d(xs, ws, 3, 3, padding=1, out=out, accumulate=True, act=1)
    assert rel(out, F.relu(nhwc(y) + base)) < 2e-5
    for bn in (16, 32, 64, 128):
        got = conv.conv2d(xs, ws, 3, 3, padding=1, force_bn=bn)
        assert rel(got, nhwc(y)) < 2e-5, bn


@pytest.mark.parametrize('N,Ci,H,Co', [(2, 64, 4, 64), (1, 32, 16, 48), (2, 128, 8, 64), (1, 64, 32, 32)])
def test_conv_transpose_stride2(N, Ci, H, Co):
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(N + Ci + H)
    x = torch.randn(N, Ci, H, H, generator=g).cuda()
    wt = torch.randn(Ci, Co, 3, 3, generator=g).cuda() / (Ci * 9) ** 0.5          # conv_transpose layout
    want = F.conv_transpose2d(x, wt, stride=2, padding=0)                        # (2H+1)^2
    ws = conv.pack_weights(wt.permute(1, 0, 2, 3).contiguous())                  # [9, Co, Ci]
    got = conv.conv_transpose2d_s2(conv.pack_split32(nhwc(x)), ws, 3)
    assert got.shape == nhwc(want).shape
    assert rel(got, nhwc(want)) < 2e-5
    # 6x6 composite kernel with crop 2 (transposed conv fused with the 4x4 blur)
    w6 = torch.randn(Ci, Co, 6, 6, generator=g).cuda() / (Ci * 36) ** 0.5
    want6 = F.conv_transpose2d(x, w6, stride=2, padding=2)
    got6 = conv.conv_transpose2d_s2(conv.pack_split32(nhwc(x)), conv.pack_weights(w6.permute(1, 0, 2, 3).contiguous()),
                                    6, crop=2)
    assert got6.shape == nhwc(want6).shape == (N, 2 * H, 2 * H, Co)
    assert rel(got6, nhwc(want6)) < 2e-5


def test_conv_transpose_phases_on_forked_streams_equal_the_serial_launches(monkeypatch):
    """conv.PHASE_STREAMS_MAX: the four output-phase launches of a small up-conv run on forked streams (parallel branches of a
    captured graph); same kernels, disjoint output pixels -> bit-identical to the serial order, eagerly and under capture."""
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 64, 8, 8, generator=g).cuda()
    wt = torch.randn(64, 96, 3, 3, generator=g).cuda() / 24.0
    xs, ws = conv.pack_split32(nhwc(x)), conv.pack_weights(wt.permute(1, 0, 2, 3).contiguous())
    monkeypatch.setattr(conv, 'PHASE_STREAMS_MAX', 0)
    serial = conv.conv_transpose2d_s2(xs, ws, 3)
    monkeypatch.setattr(conv, 'PHASE_STREAMS_MAX', 1 << 30)
    forked = conv.conv_transpose2d_s2(xs, ws, 3)
    torch.cuda.synchronize()
    assert torch.equal(serial, forked)
    out = torch.zeros_like(serial)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            conv.conv_transpose2d_s2(xs, ws, 3, out=out)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(serial, out)


def test_simt_twin_matches(monkeypatch):
    """The CUDA-core twin of the same contract agrees with the tensor-core kernel bit-for-bit-ish."""
    import subprocess, sys, os
    code = ("import torch, torch.nn.functional as F\n"
            "from warpedganspace_b200 import conv\n"
            "g = torch.Generator().manual_seed(1)\n"
            "x = torch.randn(2, 64, 8, 8, generator=g).cuda(); w = torch.randn(32, 64, 3, 3, generator=g).cuda() / 24\n"
            "torch.backends.cudnn.allow_tf32 = False\n"
            "y = conv.conv2d(conv.pack_split32(x.permute(0,2,3,1).contiguous()), conv.pack_weights(w), 3, 3, padding=1)\n"
            "r = F.conv2d(x, w, padding=1).permute(0,2,3,1)\n"
            "print(float((y - r).norm() / r.norm()))\n")
    env = dict(os.environ, WGS_CONV_IMPL='simt')
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=120,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, out.stderr[-2000:]
    assert float(out.stdout.strip().splitlines()[-1]) < 2e-5


WG_CASES = [
    # N, Ci, H, W, Co, k, stride, pad
    (2, 32, 16, 16, 32, 3, 1, 1),
    (4, 64, 32, 32, 64, 3, 1, 1),
    (2, 64, 16, 16, 128, 3, 2, 1),
    (2, 6, 32, 32, 64, 7, 2, 3),
    (3, 128, 8, 8, 256, 3, 1, 1),
    (2, 64, 16, 16, 128, 1, 2, 0),
    (4, 2, 32, 32, 6, 5, 1, 0),
    (4, 16, 5, 5, 120, 5, 1, 0),
    (1, 160, 12, 20, 96, 3, 1, 1),
]


@pytest.mark.parametrize('case', WG_CASES)
def test_wgrad_and_dgrad_match_torch(case):
    from warpedganspace_b200 import reconstructor as R
    N, Ci, H, W, Co, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    x = torch.randn(N, Ci, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g).cuda() / (Ci * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(Co, generator=g).cuda().requires_grad_(True)
    y_ref = F.conv2d(x, w, b, stride=stride, padding=pad)
    cot = torch.randn(y_ref.shape, generator=g).cuda()
    gx_ref, gw_ref, gb_ref = torch.autograd.grad((y_ref * cot).sum(), (x, w, b))
    y = R.conv2d(x.contiguous(memory_format=torch.channels_last), w, b, stride, pad)
    assert rel(y, y_ref) < 2e-5
    gx, gw, gb = torch.autograd.grad((y * cot).sum(), (x, w, b))
    assert rel(gx, gx_ref) < 2e-5
    assert rel(gw, gw_ref) < 3e-5
    assert rel(gb, gb_ref) < 1e-5


@pytest.mark.parametrize('case', [(2, 6, 32, 32, 64, 7, 2, 3), (3, 6, 64, 48, 64, 7, 2, 3), (4, 2, 32, 32, 6, 5, 1, 0)])
def test_stem_im2col_path(case):
    """Few-input-channel convs go through im2col -> 1-tap GEMM (forward and weight gradient)."""
    from warpedganspace_b200 import reconstructor as R
    N, Ci, H, W, Co, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case) + 7)
    x = torch.randn(N, Ci, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g).cuda() / (Ci * k * k) ** 0.5).requires_grad_(True)
    y_ref = F.conv2d(x, w, None, stride=stride, padding=pad)
    cot = torch.randn(y_ref.shape, generator=g).cuda()
    gx_ref, gw_ref = torch.autograd.grad((y_ref * cot).sum(), (x, w))
    y = R.conv2d(x.contiguous(memory_format=torch.channels_last), w, None, stride, pad)
    assert y.grad_fn.__class__.__name__.startswith('_StemConvFn')
    assert rel(y, y_ref) < 2e-5
    gx, gw = torch.autograd.grad((y * cot).sum(), (x, w))
    assert rel(gx, gx_ref) < 2e-5 and rel(gw, gw_ref) < 3e-5


@pytest.mark.parametrize('N,H,W,C,pad0,Ho,Wo', [(2, 9, 9, 32, 1, 8, 8), (1, 17, 33, 64, 1, 16, 32), (3, 8, 8, 16, 2, 9, 9),
                                                (1, 6, 10, 8, 2, 7, 11), (2, 5, 7, 4, 1, 4, 6),
                                                # wide maps (interior fast path, ragged strips, both paddings)
                                                (2, 33, 65, 32, 1, 32, 64), (1, 16, 40, 64, 2, 17, 41), (1, 19, 50, 32, 1, 18, 49),
                                                (2, 64, 64, 32, 2, 65, 65),
                                                # >= 128 pixels wide with C % 32 == 0: the TMA-fed shared-memory variant
                                                (2, 129, 129, 32, 1, 128, 128), (1, 65, 257, 64, 1, 64, 256),
                                                (1, 128, 144, 32, 2, 129, 145), (2, 70, 200, 96, 1, 69, 199)])
def test_fir4_act_matches_torch(N, H, W, C, pad0, Ho, Wo):
    """Separable 4-tap FIR + per-sample scale + noise + bias + sqrt2*lrelu against upfirdn2d-style torch code."""
    import ctypes
    from warpedganspace_b200 import _lib
    g = torch.Generator().manual_seed(N + H + W + C)
    y = torch.randn(N, H, W, C, generator=g).cuda()
    alpha = (torch.rand(N, C, generator=g) + 0.5).cuda()
    beta = torch.randn(C, generator=g).cuda()
    noise = torch.randn(Ho, Wo, generator=g).cuda()
    taps = (ctypes.c_float * 4)(0.25, 0.75, 0.75, 0.25)
    out = torch.empty(N, Ho, Wo, C, device='cuda')
    _lib.call('wgs_fir4_act', _lib.ptr(y), _lib.ptr(out), N, H, W, Ho, Wo, C, pad0, taps, _lib.ptr(alpha), _lib.ptr(beta),
              _lib.ptr(noise), 0.3, 3, None, None, 0, 0, _lib.stream())
    k1 = torch.tensor([0.25, 0.75, 0.75, 0.25], device='cuda')
    k2 = torch.outer(k1, k1)
    yp = F.pad(y.permute(0, 3, 1, 2), [pad0, Wo + 3 - W - pad0, pad0, Ho + 3 - H - pad0])
    f = F.conv2d(yp, torch.flip(k2, [0, 1]).view(1, 1, 4, 4).expand(C, 1, 4, 4), groups=C)
    want = f * alpha[:, :, None, None] + 0.3 * noise[None, None] + beta[None, :, None, None]
    want = 2 ** 0.5 * F.leaky_relu(want, 0.2)
    assert rel(out, want.permute(0, 2, 3, 1)) < 1e-6


@pytest.mark.parametrize('N,H,W,C,pad0', [(3, 129, 161, 32, 1), (2, 64, 130, 64, 2), (2, 33, 40, 32, 1)])
def test_fir4_act_split_output_and_selective_fp32(N, H, W, C, pad0):
    """The fused outputs of fir4_act (both kernel variants): split32(out * split_scale[n, c]) for every image, fp32 only for
    images n >= out_from_n (the back-propagated half of a pair batch), and the transposed-blur mode of the backward pass
    (pad0 = 2, no bias / noise / activation, split32 output only)."""
    import ctypes
    from warpedganspace_b200 import _lib
    g = torch.Generator().manual_seed(N + H + W + C)
    Ho, Wo = (H - 1, W - 1) if pad0 == 1 else (H + 1, W + 1)
    y = torch.randn(N, H, W, C, generator=g).cuda()
    alpha = (torch.rand(N, C, generator=g) + 0.5).cuda()
    beta = torch.randn(C, generator=g).cuda()
    noise = torch.randn(Ho, Wo, generator=g).cuda()
    sc = (torch.rand(N, C + 8, generator=g) + 0.5).cuda()[:, 4: 4 + C]            # a strided view, as the style slices are
    taps = (ctypes.c_float * 4)(0.25, 0.75, 0.75, 0.25)
    out = torch.full((N, Ho, Wo, C), 7.0, device='cuda')
    xs = torch.empty(N, Ho, Wo, C // 32, 64, dtype=torch.bfloat16, device='cuda')
    k1 = torch.tensor([0.25, 0.75, 0.75, 0.25], device='cuda')
    yp = F.pad(y.permute(0, 3, 1, 2), [pad0, Wo + 3 - W - pad0, pad0, Ho + 3 - H - pad0])
    f = F.conv2d(yp, torch.flip(torch.outer(k1, k1), [0, 1]).view(1, 1, 4, 4).expand(C, 1, 4, 4), groups=C).permute(0, 2, 3, 1)
    if pad0 == 1:
        _lib.call('wgs_fir4_act', _lib.ptr(y), _lib.ptr(out), N, H, W, Ho, Wo, C, pad0, taps, _lib.ptr(alpha), _lib.ptr(beta),
                  _lib.ptr(noise), 0.3, 3, _lib.ptr(xs), ctypes.c_void_p(sc.data_ptr()), sc.stride(0), 1, _lib.stream())
        want = 2 ** 0.5 * F.leaky_relu(f * alpha[:, None, None, :] + 0.3 * noise[None, :, :, None] + beta, 0.2)
        assert float((out[0] - 7.0).abs().max()) == 0.0                            # image 0 is not back-propagated
        assert rel(out[1:], want[1:]) < 1e-6
        assert rel(_unsplit(xs), want * sc[:, None, None, :]) < 2e-5
    else:
        _lib.call('wgs_fir4_act', _lib.ptr(y), None, N, H, W, Ho, Wo, C, pad0, taps, _lib.ptr(alpha), None, None, 0.0, 0,
                  _lib.ptr(xs), None, 0, 0, _lib.stream())
        assert rel(_unsplit(xs), f * alpha[:, None, None, :]) < 2e-5


@pytest.mark.parametrize('N,Ci,H,W,Co,k,pad', [(2, 6, 32, 32, 64, 7, 3), (1, 6, 33, 35, 64, 7, 3), (2, 64, 32, 32, 128, 3, 1),
                                               (3, 64, 17, 16, 96, 1, 0), (2, 32, 64, 48, 32, 3, 1), (1, 128, 16, 16, 64, 3, 1)])
def test_strided_dgrad_in_one_launch_matches_torch(N, Ci, H, W, Co, k, pad):
    """Phase-packed output: all four output phases of a stride-2 data-gradient in one launch (odd sizes, 6-channel
    groups narrower than a 16-channel block, groups of 32 / 64 / 128 channels, 1x1 with empty phases)."""
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(N + Ci + H + k)
    x = torch.randn(N, Ci, H, W, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).cuda()
    y = F.conv2d(x, w, stride=2, padding=pad)
    dy = torch.randn(y.shape, generator=g).cuda()
    y.backward(dy)
    before = conv._lib.launch_count()
    got = conv.conv_dgrad_merged(conv.pack_split32(nhwc(dy)), w, (H, W), 2, pad)
    assert got.shape == (N, H, W, Ci)
    assert rel(got, nhwc(x.grad)) < 3e-5
    base = torch.randn(N, H, W, Ci, generator=g).cuda()
    out = base.clone()
    conv.conv_dgrad_merged(conv.pack_split32(nhwc(dy)), w, (H, W), 2, pad, out=out, accumulate=True)
    assert rel(out, nhwc(x.grad) + base) < 3e-5
    assert conv._lib.launch_count() - before <= 6            # 2 x (pack dy + pack weights + ONE conv)


@pytest.mark.parametrize('N,Ci,H,W,Co', [(2, 64, 8, 8, 32), (1, 128, 16, 12, 64), (3, 32, 5, 7, 16)])
def test_conv_transpose_stride2_in_one_launch(N, Ci, H, W, Co):
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(N + Ci + H)
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5).cuda()                  # StyleGAN2 layout [Co, Ci, 3, 3]
    want = F.conv_transpose2d(x, w.transpose(0, 1), stride=2, padding=0)                  # model.py:206-212
    shifts, idx, G = conv._phase_plan('convT', 3, 3, 2, 0, x.device)
    wm = conv.merged_phase_weights(w.reshape(Co, Ci, 9).contiguous(), idx, len(shifts), G)
    got = conv.conv_transpose2d_s2_merged(conv.pack_split32(nhwc(x)), wm, 3, Co)
    assert got.shape == nhwc(want).shape == (N, 2 * H + 1, 2 * W + 1, Co)
    assert rel(got, nhwc(want)) < 2e-5


@pytest.mark.parametrize('Co', [16, 24, 32, 48, 64])
def test_stacked_weight_layout_matches_row_layout(Co, monkeypatch):
    """Weights with <= 64 rows run the two-MMA-per-K-slice kernels on the stacked layout; the row layout (three MMAs)
    must give the same numbers (both are exact products of the same bf16 pieces, fp32 accumulation order aside)."""
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(Co)
    x = torch.randn(2, 96, 24, 40, generator=g).cuda()
    w = (torch.randn(Co, 96, 3, 3, generator=g) / (96 * 9) ** 0.5).cuda()
    xs = conv.pack_split32(nhwc(x))
    want = nhwc(F.conv2d(x, w, padding=1))
    assert conv.weight_layout(conv.pack_weights(w)) == 1
    got = conv.conv2d(xs, conv.pack_weights(w), 3, 3, padding=1)
    monkeypatch.setattr(conv, 'STACK_MAX_COUT', 0)
    assert conv.weight_layout(conv.pack_weights(w)) == 0
    rows = conv.conv2d(xs, conv.pack_weights(w), 3, 3, padding=1)
    assert rel(got, want) < 2e-5 and rel(rows, want) < 2e-5 and rel(got, rows) < 2e-6


def _unsplit(xs):
    """split32 [..., chunks, 64] bf16 -> fp32 [..., chunks*32] (hi + lo)."""
    f = xs.float()
    return (f[..., :32] + f[..., 32:]).flatten(-2)


@pytest.mark.parametrize('N,Ci,H,W,Co', [(2, 32, 48, 136, 32), (1, 24, 20, 200, 24), (3, 32, 16, 128, 16),
                                         (2, 64, 48, 136, 64), (1, 40, 20, 200, 24), (1, 64, 16, 128, 32),      # two chunks
                                         (1, 32, 32, 1040, 32)])                                                # 16 tiles per CTA
def test_multi_tile_halo_kernel_plain_and_fused(N, Ci, H, W, Co):
    """Layers with <= 64 input / output channels wide enough for the multi-tile halo kernel (resident tap weights of one or
    two 32-channel chunks, patch ring, two TMEM accumulators; ragged last tile group, ragged rows; 8 and - at >= 512 pixels
    of width, the variant the 1024^2 layers run - 16 tiles per CTA): plain epilogue with demod / bias / noise / activation, and the
    fused epilogue (next layer's split32 operand, ToRGB accumulation, fp32 only for the back-propagated images)."""
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(N + H + W)
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5).cuda()
    alpha = (torch.rand(N, Co, generator=g) + 0.5).cuda()
    beta = torch.randn(Co, generator=g).cuda()
    noise = torch.randn(H, W, generator=g).cuda()
    xs, ws = conv.pack_split32(nhwc(x)), conv.pack_weights(w)
    y = nhwc(F.conv2d(x, w, padding=1))
    want = F.leaky_relu(y * alpha[:, None, None, :] + 0.3 * noise[None, :, :, None] + beta, 0.2) * 2 ** 0.5
    got = conv.conv2d(xs, ws, 3, 3, padding=1, alpha=alpha, beta=beta, noise=noise, noise_w=0.3, act=3)
    assert rel(got, want) < 2e-5
    if Co % 32:
        return
    sc = (torch.rand(N, Co, generator=g) + 0.5).cuda()
    rgb_w = torch.randn(N, 3, Co, generator=g).cuda()
    rgb0 = torch.randn(N, H, W, 3, generator=g).cuda()
    rgb = rgb0.clone()
    out = torch.full((N, H, W, Co), 7.0).cuda()
    nxt = torch.empty(N, H, W, Co // 32, 64, dtype=torch.bfloat16).cuda()
    conv.conv2d(xs, ws, 3, 3, padding=1, out=out, alpha=alpha, beta=beta, noise=noise, noise_w=0.3, act=3,
                out_split=nxt, split_scale=sc, out_from_n=1, rgb_w=rgb_w, rgb_out=rgb)
    assert float((out[0] - 7.0).abs().max()) == 0.0                       # image 0 is not back-propagated: no fp32 store
    assert rel(out[1:], want[1:]) < 2e-5
    assert rel(_unsplit(nxt), want * sc[:, None, None, :]) < 2e-5
    assert rel(rgb - rgb0, torch.einsum('nhwc,noc->nhwo', want, rgb_w)) < 2e-5


def test_cluster_split_k_tiny_m_layers():
    """Opt-in (WGS_CONV_SPLITK=1): tiny-M, deep-K launches (the 512-channel layers at 4^2 .. 32^2) run as thread-block
    clusters - the K range is split over up to 8 CTAs per output tile, every CTA parks its partial accumulator in shared
    memory and finishes BN / ksplit of the columns by summing its peers' partials through distributed shared memory
    inside the normal epilogue (demod / bias / activation, fused outputs, several images per tile, ragged channel tiles)."""
    import subprocess, sys, os
    code = """
import torch, torch.nn.functional as F
from warpedganspace_b200 import conv
torch.backends.cudnn.allow_tf32 = False
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
unsplit = lambda xs: (xs.float()[..., :32] + xs.float()[..., 32:]).flatten(-2)
worst = 0.0
for N, Ci, H, W, Co in [(2, 512, 8, 8, 512), (8, 512, 4, 4, 512), (4, 256, 16, 16, 384), (1, 512, 32, 32, 512), (3, 160, 8, 12, 200)]:
    g = torch.Generator().manual_seed(N + Ci + H + Co)
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5).cuda()
    alpha = (torch.rand(N, Co, generator=g) + 0.5).cuda()
    beta = torch.randn(Co, generator=g).cuda()
    xs, ws = conv.pack_split32(nhwc(x)), conv.pack_weights(w)
    y = nhwc(F.conv2d(x, w, padding=1))
    worst = max(worst, rel(conv.conv2d(xs, ws, 3, 3, padding=1), y))
    want = F.leaky_relu(y * alpha[:, None, None, :] + beta, 0.2) * 2 ** 0.5
    worst = max(worst, rel(conv.conv2d(xs, ws, 3, 3, padding=1, alpha=alpha, beta=beta, act=3), want))
    if Co % 32 == 0:
        sc = (torch.rand(N, Co, generator=g) + 0.5).cuda()
        rgb_w = torch.randn(N, 3, Co, generator=g).cuda()
        rgb = torch.zeros(N, H, W, 3).cuda()
        nxt = torch.empty(N, H, W, Co // 32, 64, dtype=torch.bfloat16).cuda()
        out = torch.empty(N, H, W, Co).cuda()
        conv.conv2d(xs, ws, 3, 3, padding=1, out=out, alpha=alpha, beta=beta, act=3, out_split=nxt, split_scale=sc,
                    rgb_w=rgb_w, rgb_out=rgb)
        worst = max(worst, rel(out, want), rel(unsplit(nxt), want * sc[:, None, None, :]),
                    rel(rgb, torch.einsum('nhwc,noc->nhwo', want, rgb_w)))
print(worst)
"""
    env = dict(os.environ, WGS_CONV_SPLITK='1')
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=180,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, out.stderr[-2000:]
    assert float(out.stdout.strip().splitlines()[-1]) < 2e-5


def test_fused_epilogue_multi_image_tiles_under_memcheck():
    """compute-sanitizer over the fused epilogue on tiles that span more images than the batch holds (4x4 and 8x8 maps,
    batch 2 and 3): lanes beyond the batch take part in the warp-collective stores and must not read per-image constants
    out of bounds (a stale read that only faults when the allocator has nothing mapped behind the tensor)."""
    import shutil, subprocess, sys, os
    tool = shutil.which('compute-sanitizer') or '/usr/local/cuda/bin/compute-sanitizer'
    if not os.path.exists(tool):
        pytest.skip('compute-sanitizer not installed')
    code = """
import torch
from warpedganspace_b200 import conv
for N, H, C in ((2, 4, 64), (3, 8, 32), (2, 16, 32)):
    g = torch.Generator().manual_seed(N + H)
    x = torch.randn(N, H, H, C, generator=g).cuda()
    w = (torch.randn(C, C, 3, 3, generator=g) / (C * 9) ** 0.5).cuda()
    xs, ws = conv.pack_split32(x), conv.pack_weights(w)
    alpha = torch.rand(N, C, generator=g).cuda() + 0.5
    sc = torch.rand(N, C, generator=g).cuda()
    rgb_w = torch.randn(N, 3, C, generator=g).cuda()
    rgb = torch.zeros(N, H, H, 3).cuda()
    nxt = torch.empty(N, H, H, C // 32, 64, dtype=torch.bfloat16).cuda()
    out = torch.empty(N, H, H, C).cuda()
    conv.conv2d(xs, ws, 3, 3, padding=1, out=out, alpha=alpha, act=3, out_split=nxt, split_scale=sc, out_from_n=1,
                rgb_w=rgb_w, rgb_out=rgb)
    conv.conv2d(xs, ws, 3, 3, padding=1, alpha=alpha, act=1)
torch.cuda.synchronize()
print('ok')
"""
    # --report-api-errors no: the CUDA runtime's own lazy-loading probe (cuKernelGetFunction -> CUDA_ERROR_INVALID_HANDLE on
    # the first launch of every process) is an API return code, not a memory error (profiles/r02_sanitizer.md); the caching
    # allocator is switched off so that every tensor is its own cudaMalloc and an out-of-bounds access cannot land in a pool
    out = subprocess.run([tool, '--tool', 'memcheck', '--report-api-errors', 'no', '--error-exitcode', '23', '--print-limit', '5',
                          sys.executable, '-c', code], env=dict(os.environ, PYTORCH_NO_CUDA_MEMORY_CACHING='1'),
                         capture_output=True, text=True, timeout=600,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and 'ok' in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


STAT_CASES = [
    # N, Ci, H, W, Co, k, stride, pad, split_k       kernel this lands on
    (2, 128, 32, 32, 128, 3, 1, 1, 0),             # conv_tc, one tile per CTA
    (2, 64, 64, 256, 64, 3, 1, 1, 0),              # multi-tile halo kernel, two chunks (statistics flushed after the last tile)
    (1, 32, 128, 128, 32, 3, 1, 1, 0),             # multi-tile halo kernel, one chunk
    (2, 64, 36, 40, 128, 3, 2, 1, 0),              # strided, ragged tiles (out-of-grid lanes must not count)
    (2, 512, 16, 16, 512, 3, 1, 1, 1),             # cluster split-K: every rank reduces its own columns
    (3, 64, 32, 32, 128, 1, 2, 0, 0),              # 1x1 / 2 shortcut conv
]


@pytest.mark.parametrize('case', STAT_CASES)
def test_batchnorm_statistics_in_the_conv_epilogue(case):
    """wgs_conv_desc.stat_sum / stat_sumsq / stat_shift: shifted first and second moments of the conv output per channel,
    accumulated in the epilogue (train-mode BatchNorm of the Reconstructor without a separate pass over the output)."""
    from warpedganspace_b200 import conv
    N, Ci, H, W, Co, k, stride, pad, split_k = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = torch.randn(Co, Ci, k, k, generator=g).cuda() / (Ci * k * k) ** 0.5
    shift = (0.3 * torch.randn(Co, generator=g)).cuda()
    s0, s1 = torch.zeros(Co).cuda(), torch.zeros(Co).cuda()
    out = conv.conv2d(conv.pack_split32(nhwc(x)), conv.pack_weights(w), k, k, stride=stride, padding=pad, cin=Ci,
                      split_k=split_k, stats=(s0, s1, shift))
    ref = F.conv2d(x, w, None, stride, pad)
    assert rel(out, nhwc(ref)) < 2e-5
    d = out.double() - shift.double()
    assert rel(s0, d.sum(dim=(0, 1, 2))) < 1e-5
    assert rel(s1, (d * d).sum(dim=(0, 1, 2))) < 1e-5
    # no shift given = plain sums
    s0.zero_(); s1.zero_()
    out2 = conv.conv2d(conv.pack_split32(nhwc(x)), conv.pack_weights(w), k, k, stride=stride, padding=pad, cin=Ci,
                       split_k=split_k, stats=(s0, s1, None))
    assert rel(s0, out2.double().sum(dim=(0, 1, 2))) < 1e-5 and rel(s1, (out2.double() ** 2).sum(dim=(0, 1, 2))) < 1e-5


@pytest.mark.parametrize('c', [3, 5])
def test_space_to_depth_pack_of_two_images_equals_the_pack_of_their_concatenation(c):
    """wgs_s2d_pack_split32_pair (the Reconstructor's cat([x1, x2], 1) folded into its stem's operand pack): bit-identical to
    packing the materialised concatenation; c = 3 takes the specialised one-thread-per-output-pixel kernel."""
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(5 + c)
    x1 = torch.randn(2, 12, 20, c, generator=g).cuda()
    x2 = torch.randn(2, 12, 20, c, generator=g).cuda()
    pair = conv.s2d_pack_split32(x1, x2)
    whole = conv.s2d_pack_split32(torch.cat([x1, x2], dim=3).contiguous())
    assert pair.shape == whole.shape and torch.equal(pair, whole)


@pytest.mark.parametrize('N,Ci,H,W,Co', [(2, 16, 64, 264, 3), (1, 3, 72, 256, 16), (2, 32, 64, 320, 32)])
def test_single_chunk_1x1_conv_on_a_wide_map(N, Ci, H, W, Co):
    """ProgGAN's to-RGB layer (16 -> 3, 1 x 1, models/ProgGAN/model.py:86-90) and its data gradient at 1024^2 are bandwidth
    work: they run on the multi-tile kernel with the single tap resident instead of one 128-pixel tile per CTA."""
    from warpedganspace_b200 import conv
    g = torch.Generator().manual_seed(N * 7 + Ci + Co)
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = (torch.randn(Co, Ci, 1, 1, generator=g) / Ci ** 0.5).cuda()
    b = torch.randn(Co, generator=g).cuda()
    got = conv.conv2d(conv.pack_split32(nhwc(x)), conv.pack_weights(w), 1, 1, beta=b, cin=Ci)
    assert rel(got, nhwc(F.conv2d(x, w, b))) < 2e-5


@pytest.mark.parametrize('N,Ci,H,W,Co,up', [(2, 32, 16, 40, 16, False), (2, 64, 24, 24, 32, False), (3, 128, 16, 16, 128, False),
                                            (2, 32, 16, 24, 16, True), (2, 256, 16, 16, 128, True), (1, 32, 64, 136, 16, False)])
def test_pixel_norm_fused_into_the_conv_epilogue(N, Ci, H, W, Co, up):
    """wgs_conv_desc.pixnorm_eps: the split32 output holds pixel_norm(leaky_relu(conv + b)) (ProgGAN's block boundary,
    models/ProgGAN/model.py:17-18,42-62), the fp32 output the un-normalised activation of the back-propagated rows only;
    16-channel outputs fill half a chunk (upper half written as zeros); with `up` the four output-phase launches of
    conv3x3(nearest_x2(x)) address the split32 tensor by output pixel."""
    from warpedganspace_b200 import conv
    from warpedganspace_b200.generators import up_conv_weights, up_conv_forward
    g = torch.Generator().manual_seed(N + Ci + Co + H)
    x = torch.randn(N, Ci, H, W, generator=g).cuda()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5).cuda()
    b = torch.randn(Co, generator=g).cuda()
    xs = conv.pack_split32(nhwc(x))
    src = F.interpolate(x, scale_factor=2, mode='nearest') if up else x
    act = F.leaky_relu(F.conv2d(src, w, b, padding=1), 0.2)
    want = nhwc(act * torch.rsqrt((act * act).mean(dim=1, keepdim=True) + 1e-8))
    oh, ow = act.shape[2], act.shape[3]
    out = torch.full((N, oh, ow, Co), 7.0).cuda()
    nxt = torch.full((N, oh, ow, (Co + 31) // 32, 64), 3.0, dtype=torch.bfloat16).cuda()
    if up:
        w_fwd, taps, _ = up_conv_weights(w)
        up_conv_forward(xs, w_fwd, taps, Co, Ci, out=out, beta=b, act=2, out_split=nxt, pixnorm_eps=1e-8, out_from_n=1)
    else:
        conv.conv2d(xs, conv.pack_weights(w), 3, 3, padding=1, out=out, beta=b, act=2, cin=Ci, out_split=nxt, pixnorm_eps=1e-8,
                    out_from_n=1)
    assert float((out[0] - 7.0).abs().max()) == 0.0 and rel(out[1:], nhwc(act)[1:]) < 2e-5
    got = _unsplit(nxt)
    assert rel(got[..., :Co], want) < 3e-5
    if Co % 32:
        assert float(got[..., Co:].abs().max()) == 0.0                  # the padding half-chunk is zero, not stale
