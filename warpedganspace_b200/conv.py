"""Host-side wrappers of the tensor-core convolution entry points (wgs_pack_split32 /
wgs_conv_split32).  Activations are NHWC; conv operands are "split32" bf16 (see include/wgs_b200.h)."""
import ctypes
import os

import torch

from . import _lib

MAX_TAPS = 64

# Optional per-launch profiling (bench.py): when PROFILE is a list, every tensor-core launch appends
# (kind, algorithmic_flops, start_event, end_event) recorded on the launching stream.
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_end(e0, kind, flops, nbytes=0, issued=None, tag=''):
    """flops = ALGORITHMIC FLOPs of the launch (real taps, true channel counts); issued = logical FLOPs the kernel
    actually contracts (zero weight blocks of phase-packed launches and channel padding to 32 included)."""
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.append((kind, flops, e0, e1, nbytes, flops if issued is None else issued, tag))


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ('inp', ctypes.c_void_p),
        ('in_n', ctypes.c_int), ('in_h', ctypes.c_int), ('in_w', ctypes.c_int), ('c_chunks', ctypes.c_int),
        ('w', ctypes.c_void_p),
        ('w_taps', ctypes.c_int), ('w_cout', ctypes.c_int),
        ('out_n', ctypes.c_int), ('grid_h', ctypes.c_int), ('grid_w', ctypes.c_int), ('in_stride', ctypes.c_int),
        ('num_taps', ctypes.c_int),
        ('tap_dy', ctypes.c_int * MAX_TAPS), ('tap_dx', ctypes.c_int * MAX_TAPS), ('tap_w', ctypes.c_int * MAX_TAPS),
        ('out', ctypes.c_void_p),
        ('out_sn', ctypes.c_longlong), ('out_sy', ctypes.c_longlong), ('out_sx', ctypes.c_longlong),
        ('out_y0', ctypes.c_int), ('out_x0', ctypes.c_int), ('out_ystep', ctypes.c_int), ('out_xstep', ctypes.c_int),
        ('cout', ctypes.c_int),
        ('alpha', ctypes.c_void_p), ('beta', ctypes.c_void_p),
        ('act', ctypes.c_int), ('accumulate', ctypes.c_int), ('force_bn', ctypes.c_int),
        ('noise', ctypes.c_void_p), ('noise_w', ctypes.c_float), ('noise_ld', ctypes.c_int),
        ('out_split', ctypes.c_void_p), ('split_scale', ctypes.c_void_p), ('split_scale_ld', ctypes.c_longlong),
        ('out_from_n', ctypes.c_int), ('rgb_w', ctypes.c_void_p), ('rgb_out', ctypes.c_void_p),
        ('group_size', ctypes.c_int), ('group_w', ctypes.c_int), ('out_h', ctypes.c_int), ('out_w', ctypes.c_int),
        ('w_layout', ctypes.c_int), ('split_k', ctypes.c_int),
        ('stat_sum', ctypes.c_void_p), ('stat_sumsq', ctypes.c_void_p), ('stat_shift', ctypes.c_void_p),
        ('pixnorm_eps', ctypes.c_float),
    ]


def _check_layout():
    lib = _lib.load()
    if lib.wgs_conv_desc_size() != ctypes.sizeof(ConvDesc):
        raise RuntimeError('wgs_conv_desc layout mismatch: C %d vs ctypes %d'
                           % (lib.wgs_conv_desc_size(), ctypes.sizeof(ConvDesc)))


def chunks_of(c):
    return (c + 31) // 32


def pack_split32(x, scale=None, rows_per_group=1, out=None):
    """x: fp32 [..., C] with contiguous rows (last-dim stride 1; a uniform row stride is allowed).
    Returns bf16 [..., ceil(C/32), 64].  scale: fp32 [groups, C]; row r uses group r // rows_per_group."""
    C = x.shape[-1]
    rows = x.numel() // C
    if not x.is_contiguous():
        x = x.contiguous()
    ld = C
    if out is None:
        out = torch.empty(*x.shape[:-1], chunks_of(C), 64, dtype=torch.bfloat16, device=x.device)
    scale_ld = scale.stride(0) if scale is not None else 0
    if scale is not None:
        assert scale.stride(1) == 1 and scale.shape[1] == C
    _lib.call('wgs_pack_split32', _lib.ptr(x), rows, C, ld,
              ctypes.c_void_p(scale.data_ptr()) if scale is not None else None, scale_ld, int(rows_per_group),
              _lib.ptr(out), _lib.stream())
    return out


# Weight tensors with at most STACK_MAX_COUT rows per tap are stored in the STACKED layout
# [T][chunks][hi | lo][Co][32] (include/wgs_b200.h, wgs_conv_desc.w_layout = 1) and run the two-MMA-per-K-slice
# kernels; wider ones keep the row layout [T][Co][chunks][hi32 | lo32].  Both are returned with the nominal shape
# [T, Co, chunks, 64]; every weight pack goes through pack_weight_rows so that conv_taps can infer the layout.
STACK_MAX_COUT = 64 if os.environ.get('WGS_STACK', '1') != '0' else 0


def weight_layout(w_split):
    return 1 if w_split.shape[1] <= STACK_MAX_COUT else 0


def pack_weight_rows(wt):
    """wt: fp32 [T, Co, Ci] contiguous -> bf16 [T, Co, ceil(Ci/32), 64] in the layout weight_layout() reports."""
    T, co, ci = wt.shape
    if co > STACK_MAX_COUT:
        return pack_split32(wt)
    if not wt.is_contiguous():
        wt = wt.contiguous()
    out = torch.empty(T, co, chunks_of(ci), 64, dtype=torch.bfloat16, device=wt.device)
    _lib.call('wgs_pack_weights_stacked', _lib.ptr(wt), T, co, ci, ci, _lib.ptr(out), _lib.stream())
    return out


class PackProblem(ctypes.Structure):
    """Mirror of wgs_pack_problem (include/wgs_b200.h)."""
    _fields_ = [('src', ctypes.c_void_p), ('dst', ctypes.c_void_p), ('co', ctypes.c_int), ('ci', ctypes.c_int),
                ('kh', ctypes.c_int), ('kw', ctypes.c_int), ('mode', ctypes.c_int), ('layout', ctypes.c_int),
                ('S', ctypes.c_int), ('G', ctypes.c_int), ('idx', ctypes.c_byte * 64),
                ('ci_src', ctypes.c_int), ('ci_off', ctypes.c_int)]


PACK_FWD, PACK_TRANSPOSED, PACK_IM2COL, PACK_MERGED_DGRAD, PACK_S2D = 0, 1, 2, 3, 4
_pack_layout_checked = False


def pack_weights_group(specs):
    """Every weight pack of a network in one launch (wgs_pack_weights_group).  specs: list of (w [Co, Ci, kh, kw] fp32
    contiguous CUDA tensor, mode, extra) with extra = (stride, padding) for PACK_MERGED_DGRAD, the padding for PACK_S2D
    (see s2d_geometry), None otherwise.
    Returns the packed tensors with nominal shape [T, rows, chunks, 64] (same as pack_weights / merged_phase_weights)."""
    global _pack_layout_checked
    lib = _lib.load()
    if not _pack_layout_checked:
        if lib.wgs_pack_problem_size() != ctypes.sizeof(PackProblem):
            raise RuntimeError('wgs_pack_problem layout mismatch between header and ctypes mirror')
        _pack_layout_checked = True
    arr = (PackProblem * len(specs))()
    outs = []
    for q, (w, mode, extra) in zip(arr, specs):
        if not (w.is_cuda and w.is_contiguous() and w.dtype == torch.float32):
            raise RuntimeError('pack_weights_group needs contiguous fp32 CUDA weights (no CPU fallback)')
        co, ci, kh, kw = w.shape
        q.src, q.co, q.ci, q.kh, q.kw, q.mode = w.data_ptr(), co, ci, kh, kw, mode
        if mode == PACK_FWD:
            T, rows, K = kh * kw, co, ci
        elif mode == PACK_TRANSPOSED:
            T, rows, K = kh * kw, ci, co
        elif mode == PACK_IM2COL:
            T, rows, K = 1, co, kh * kw * ci
        elif mode == PACK_S2D:
            S, a_min, G = s2d_geometry(kh, extra)
            assert kh == kw
            T, rows, K = S * S, co, 4 * ci
            q.S, q.G = S, G
        else:
            stride, padding = extra[:2]
            shifts, idx, G = _phase_plan('dgrad', kh, kw, stride, padding, w.device)
            if len(extra) == 4:                  # data gradient w.r.t. input channels [ci_off, ci_off + ci_n) only
                q.ci_src, q.ci_off, q.ci = ci, extra[2], extra[3]
                ci = extra[3]
            T, rows, K = len(shifts), G * ci, co
            q.S, q.G = T, G
            for i, t in enumerate(_phase_idx_host('dgrad', kh, kw, stride, padding)):
                q.idx[i] = t
        q.layout = 1 if rows <= STACK_MAX_COUT else 0
        out = torch.empty(T, rows, chunks_of(K), 64, dtype=torch.bfloat16, device=w.device)
        q.dst = out.data_ptr()
        outs.append(out)
    _lib.check(lib.wgs_pack_weights_group(arr, len(specs), _lib.stream()))
    return outs


def s2d_geometry(k, padding):
    """A stride-2 k-tap (per axis) conv with `padding` as a stride-1 conv over the 2x2 space-to-depth input: input index
    2*o + ky - padding = 2*(o + a) + p with a = floor((ky - padding) / 2) in [a_min, a_max].
    -> (taps per axis S, a_min, kernel offset G) with ky = 2*(a - a_min) + p - G."""
    a_min, a_max = (-padding) // 2, (k - 1 - padding) // 2
    return a_max - a_min + 1, a_min, -(padding + 2 * a_min)


def s2d_pack_split32(x_nhwc, x2_nhwc=None):
    """fp32 NHWC [N, H, W, C] (H, W even) -> split32 [N, H/2, W/2, ceil(4C/32), 64], channel (py*2+px)*C + c.
    With x2_nhwc: the same over the channel concatenation [x_nhwc ; x2_nhwc] without materialising it."""
    n, h, w, c = x_nhwc.shape
    if x2_nhwc is None:
        out = torch.empty(n, h // 2, w // 2, chunks_of(4 * c), 64, dtype=torch.bfloat16, device=x_nhwc.device)
        _lib.call('wgs_s2d_pack_split32', _lib.ptr(x_nhwc), n, h, w, c, _lib.ptr(out), _lib.stream())
        return out
    assert x2_nhwc.shape[:3] == x_nhwc.shape[:3] and x_nhwc.is_contiguous() and x2_nhwc.is_contiguous()
    c2 = x2_nhwc.shape[3]
    out = torch.empty(n, h // 2, w // 2, chunks_of(4 * (c + c2)), 64, dtype=torch.bfloat16, device=x_nhwc.device)
    _lib.call('wgs_s2d_pack_split32_pair', _lib.ptr(x_nhwc), _lib.ptr(x2_nhwc), n, h, w, c, c2, _lib.ptr(out), _lib.stream())
    return out


def s2d_taps(k, padding):
    S, a_min, _ = s2d_geometry(k, padding)
    return [(ty + a_min, tx + a_min, ty * S + tx) for ty in range(S) for tx in range(S)], S


def pack_weights(w):
    """w: fp32 [Co, Ci, kh, kw] (torch conv layout) -> bf16 [kh*kw, Co, ceil(Ci/32), 64]; tap = ky*kw+kx."""
    co, ci, kh, kw = w.shape
    return pack_weight_rows(w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous())


def conv_taps(x_split, w_split, taps, out, *, grid, in_stride=1, out_origin=(0, 0), out_step=(1, 1), cout=None,
              alpha=None, beta=None, act=0, accumulate=False, force_bn=0, noise=None, noise_w=0.0, cin=None,
              out_split=None, split_scale=None, out_from_n=0, rgb_w=None, rgb_out=None, out_n=None, groups=None,
              algo_macs_per_pixel=None, split_k=False, stats=None, pixnorm_eps=0.0):
    """Generic tap-list conv.  x_split [N, H, W, chunks, 64] bf16; w_split [T, Co, chunks, 64] bf16;
    taps: list of (dy, dx, weight_tap); out: fp32 NHWC [N, OH, OW, Cstride] (any strides, channel stride 1);
    grid: (grid_h, grid_w) virtual output grid; output pixel = grid*out_step + out_origin."""
    _check_layout()
    n, h, w_, chunks, _ = x_split.shape
    assert w_split.shape[2] == chunks, (w_split.shape, x_split.shape)
    d = ConvDesc()
    d.inp = x_split.data_ptr()
    d.in_n, d.in_h, d.in_w, d.c_chunks = n, h, w_, chunks
    d.w = w_split.data_ptr()
    d.w_taps, d.w_cout = w_split.shape[0], w_split.shape[1]
    d.w_layout = weight_layout(w_split)
    d.out_n, d.grid_h, d.grid_w, d.in_stride = (out.shape[0] if out is not None else out_n), grid[0], grid[1], in_stride
    d.num_taps = len(taps)
    for i, (dy, dx, tw) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i], d.tap_w[i] = dy, dx, tw
    if out is not None:
        assert out.dtype == torch.float32 and out.stride(3) == 1
        d.out = out.data_ptr()
        d.out_sn, d.out_sy, d.out_sx = out.stride(0), out.stride(1), out.stride(2)
    if out_split is not None:
        assert out_split.is_contiguous() and out_split.dtype == torch.bfloat16
        d.out_split = out_split.data_ptr()
        if split_scale is not None:
            assert split_scale.stride(1) == 1
            d.split_scale, d.split_scale_ld = split_scale.data_ptr(), split_scale.stride(0)
    d.out_from_n = out_from_n
    if stats is not None:                        # (sum, sumsq, shift or None): BatchNorm statistics of the output, in the epilogue
        d.stat_sum, d.stat_sumsq = stats[0].data_ptr(), stats[1].data_ptr()
        d.stat_shift = stats[2].data_ptr() if stats[2] is not None else None
    d.split_k = int(split_k)                     # 0 off, 1 batch-dependent (Reconstructor), 2 geometry-only (frozen generators)
    if rgb_out is not None:
        assert rgb_w.is_contiguous() and rgb_out.is_contiguous()
        d.rgb_w, d.rgb_out = rgb_w.data_ptr(), rgb_out.data_ptr()
    d.out_y0, d.out_x0 = out_origin
    d.out_ystep, d.out_xstep = out_step
    d.cout = cout if cout is not None else w_split.shape[1]
    d.alpha = alpha.data_ptr() if alpha is not None else None
    d.beta = beta.data_ptr() if beta is not None else None
    d.act, d.accumulate, d.force_bn = act, int(accumulate), force_bn
    if groups is not None:                       # phase-packed output: (group_size, group_w), see include/wgs_b200.h
        d.group_size, d.group_w = groups
        d.out_h, d.out_w = out.shape[1], out.shape[2]
    d.pixnorm_eps = float(pixnorm_eps)
    if out_split is not None and (tuple(out_origin) != (0, 0) or tuple(out_step) != (1, 1)):
        d.out_h, d.out_w = out_split.shape[1], out_split.shape[2]      # split32 output addressed by output pixel
    if noise is not None:
        assert noise.is_contiguous() and noise.dim() == 2
        d.noise, d.noise_w, d.noise_ld = noise.data_ptr(), float(noise_w), noise.shape[1]
    for t in (x_split, w_split):
        if not t.is_cuda:
            raise RuntimeError('conv needs CUDA tensors; there is no CPU fallback')
    e0 = _prof_begin()
    _lib.check(_lib.load().wgs_conv_split32(ctypes.byref(d), _lib.stream()))
    if e0 is not None:
        # algorithmic HBM bytes of this launch: operands once, every output once (accumulate / ToRGB: read + write)
        npix = d.out_n * grid[0] * grid[1]
        nbytes = x_split.numel() * 2 + w_split.numel() * 2
        if out is not None:
            nbytes += (d.out_n - out_from_n) * grid[0] * grid[1] * d.cout * 4 * (2 if accumulate else 1)
        if out_split is not None:
            nbytes += out_split.numel() * 2
        if rgb_out is not None:
            nbytes += npix * 3 * 4 * 2
        # algorithmic work: real taps x true channels (phase-packed launches pass the MACs per grid pixel of their real,
        # non-zero weight blocks); issued work: what the MMAs contract, padding and zero blocks included
        macs = algo_macs_per_pixel if algo_macs_per_pixel is not None else \
            d.cout * (cin or chunks * 32) * len(taps)
        issued = 2.0 * npix * ((d.cout + 15) // 16 * 16) * chunks * 32 * len(taps)
        tag = '%s%d->%d k%d s%d @%dx%d n%d%s' % ('packed ' if groups else '', cin or chunks * 32,
                                                 groups[0] if groups else d.cout, len(taps), in_stride, grid[0], grid[1], d.out_n,
                                                 ' fused' if (out_split is not None or rgb_out is not None) else '')
        _prof_end(e0, 'conv', 2.0 * npix * macs, nbytes, issued, tag)
    return out


def conv2d(x_split, w_split, kh, kw, *, stride=1, padding=0, out=None, **kw_args):
    """Plain cross-correlation (F.conv2d semantics) on split32 operands -> fp32 NHWC."""
    n, h, w_, _, _ = x_split.shape
    oh = (h + 2 * padding - kh) // stride + 1
    ow = (w_ + 2 * padding - kw) // stride + 1
    co = kw_args.get('cout') or w_split.shape[1]
    no_f32 = kw_args.pop('no_f32', False)
    if out is None and not no_f32:
        out = torch.empty(n, oh, ow, co, dtype=torch.float32, device=x_split.device)
    taps = [(ky - padding, kx - padding, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
    return conv_taps(x_split, w_split, taps, out, grid=(oh, ow), in_stride=stride, out_n=n, **kw_args)


def conv_transpose2d_s2(x_split, w_split, k, *, out=None, crop=0, **kw_args):
    """F.conv_transpose2d(stride=2, padding=crop) semantics for a k x k kernel, as 4 output-phase convs.
    w_split holds taps in (ky*k + kx) order of the *transposed-conv* weight [Ci, Co, k, k] packed as
    [k*k, Co, Ci]:  out[2*iy + ky - crop, 2*ix + kx - crop] += x[iy, ix] * w[ky, kx]."""
    n, h, w_, _, _ = x_split.shape
    full = 2 * (h - 1) + k
    oh = full - 2 * crop
    fullw = 2 * (w_ - 1) + k
    ow = fullw - 2 * crop
    co = kw_args.get('cout') or w_split.shape[1]
    if out is None:
        out = torch.empty(n, oh, ow, co, dtype=torch.float32, device=x_split.device)
    phases = []
    for py in range(2):
        for px in range(2):
            # output Y = 2*q + py (in cropped coords); uncropped Yf = Y + crop = 2*iy + ky
            kys = [ky for ky in range(k) if (ky - py - crop) % 2 == 0]
            kxs = [kx for kx in range(k) if (kx - px - crop) % 2 == 0]
            gh = (oh - py + 1) // 2
            gw = (ow - px + 1) // 2
            if gh <= 0 or gw <= 0 or not kys or not kxs:
                continue
            taps = [((py + crop - ky) // 2, (px + crop - kx) // 2, ky * k + kx) for ky in kys for kx in kxs]
            phases.append((taps, (gh, gw), (py, px)))
    run_phases([lambda taps=taps, grid=grid, origin=origin: conv_taps(x_split, w_split, taps, out, grid=grid, out_origin=origin,
                                                                      out_step=(2, 2), **kw_args)
                for taps, grid, origin in phases], n * phases[0][1][0] * phases[0][1][1] * co, x_split.device)
    return out


def run_phases(launches, elements, device):
    """Run the output-phase launches of one up-sampling conv (callables; they write disjoint pixels of one tensor).  On the
    low-resolution layers each of them fills a fraction of the SMs (8 .. 256 CTAs of ~30 us), so they run side by side on forked
    streams (event fork / join; inside a CUDA graph capture this becomes parallel branches).  PHASE_STREAMS_MAX bounds the layer
    size (output elements of one phase); per-launch profiling keeps them serial."""
    fork = len(launches) > 1 and PROFILE is None and elements <= PHASE_STREAMS_MAX
    if not fork:
        for launch in launches:
            launch()
        return
    cur = torch.cuda.current_stream()
    side = _phase_streams(device, len(launches) - 1)
    start = torch.cuda.Event()
    start.record(cur)
    for i, launch in enumerate(launches):
        if i == 0:
            launch()
            continue
        side[i - 1].wait_event(start)
        with torch.cuda.stream(side[i - 1]):
            launch()
            done = torch.cuda.Event()
            done.record(side[i - 1])
        cur.wait_event(done)


PHASE_STREAMS_MAX = int(os.environ.get('WGS_PHASE_STREAMS_MAX', '9000000'))      # up to 256 ch @ 65 x 65 x 8 images per phase
_PHASE_STREAMS = {}


def _phase_streams(device, count):
    key = (device.index if device.index is not None else torch.cuda.current_device())
    pool = _PHASE_STREAMS.setdefault(key, [])
    while len(pool) < count:
        pool.append(torch.cuda.Stream(device=device, priority=-1))
    return pool[:count]


# ---- phase-packed (merged) strided data-gradients / transposed convs -----------------------------------------------
# Every output phase (py, px) of a stride-s data-gradient or transposed conv is a short tap list over the SAME small
# set of input shifts.  Launching them separately runs N = C-wide MMAs whose cost is dominated by fetching the
# 128-row A operand (profiles/r01_conv_ncu_step.md); stacking the phases along N gives one launch with
# N = s*s*C, one patch load per shift, and zero weight blocks where a phase has no tap at a shift.
_PHASE_PLANS = {}


def _phase_plan(kind, kh, kw, stride, padding, device):
    """-> (shifts [(sy, sx)], idx int64 [S*G] of tap ids (kh*kw = zero tap), G) for
    kind 'dgrad': dx[s*q+py] = sum_k dy[q + (py+p-k)/s] w[k];  kind 'convT': out[s*q+py] = sum_k x[q + (py-k)/s] w[k]."""
    key = (kind, kh, kw, stride, padding, str(device))
    hit = _PHASE_PLANS.get(key)
    if hit is not None:
        return hit
    s, T = stride, kh * kw
    off = padding if kind == 'dgrad' else 0
    ent = {}
    for py in range(s):
        for px in range(s):
            for ky in range(kh):
                for kx in range(kw):
                    if (py + off - ky) % s == 0 and (px + off - kx) % s == 0:
                        ent[((py + off - ky) // s, (px + off - kx) // s, py * s + px)] = ky * kw + kx
    shifts = sorted({(a, b) for a, b, _ in ent})
    G = s * s
    idx = [ent.get((sy, sx, g), T) for sy, sx in shifts for g in range(G)]
    hit = (shifts, torch.tensor(idx, dtype=torch.int64, device=device), G)
    _PHASE_PLANS[key] = hit
    return hit


_REAL_BLOCKS = {}


def _real_blocks(idx, T):
    """Number of (shift, phase) blocks of a phase plan that hold a real tap (host-side, cached: no device sync per launch)."""
    key = (idx.data_ptr(), T)
    hit = _REAL_BLOCKS.get(key)
    if hit is None:
        hit = _REAL_BLOCKS[key] = int((idx != T).sum())
    return hit


_PHASE_IDX_HOST = {}


def _phase_idx_host(kind, kh, kw, stride, padding):
    """The tap table of _phase_plan as a host list (-1 = zero block), for wgs_pack_problem.idx."""
    key = (kind, kh, kw, stride, padding)
    hit = _PHASE_IDX_HOST.get(key)
    if hit is None:
        s, T = stride, kh * kw
        ent = {}
        for py in range(s):
            for px in range(s):
                for ky in range(kh):
                    for kx in range(kw):
                        if (py + padding - ky) % s == 0 and (px + padding - kx) % s == 0:
                            ent[((py + padding - ky) // s, (px + padding - kx) // s, py * s + px)] = ky * kw + kx
        shifts = sorted({(a, b) for a, b, _ in ent})
        hit = _PHASE_IDX_HOST[key] = [ent.get((sy, sx, g), -1) for sy, sx in shifts for g in range(s * s)]
    return hit


def merged_phase_weights(w_src, idx, S, G):
    """w_src fp32 [rows, K, T] -> split32 [S, G*rows, ceil(K/32), 64] with block (s, g) = w_src[:, :, idx[s*G+g]]
    (zero where idx == T)."""
    rows, K, T = w_src.shape
    w_ext = torch.cat([w_src, w_src.new_zeros(rows, K, 1)], dim=2)
    sel = w_ext.index_select(2, idx)                                    # [rows, K, S*G]
    return pack_weight_rows(sel.permute(2, 0, 1).reshape(S, G * rows, K).contiguous())


def conv_dgrad_merged(dys, w, in_hw, stride, padding, out=None, accumulate=False, w_merged=None, ci_sub=None):
    """Data gradient of F.conv2d(x, w, stride, padding) for stride > 1 in ONE launch -> dx fp32 [N, H, W, Ci].
    ci_sub = (offset, count): the gradient w.r.t. that slice of the input channels only (w_merged packed to match)."""
    co, ci, kh, kw = w.shape
    h, wd = in_hw
    n = dys.shape[0]
    shifts, idx, G = _phase_plan('dgrad', kh, kw, stride, padding, dys.device)
    if ci_sub is not None:
        w = w[:, ci_sub[0]: ci_sub[0] + ci_sub[1]]
        ci = ci_sub[1]
    if w_merged is None:
        w_merged = merged_phase_weights(w.detach().permute(1, 0, 2, 3).reshape(ci, co, kh * kw), idx, len(shifts), G)
    dx = out if out is not None else torch.empty(n, h, wd, ci, device=dys.device, dtype=torch.float32)
    taps = [(sy, sx, i) for i, (sy, sx) in enumerate(shifts)]
    gh, gw = (h + stride - 1) // stride, (wd + stride - 1) // stride
    conv_taps(dys, w_merged, taps, dx, grid=(gh, gw), out_step=(stride, stride), cout=G * ci, cin=co,
              accumulate=accumulate, groups=(ci, stride), algo_macs_per_pixel=_real_blocks(idx, kh * kw) * ci * co)
    return dx


def conv_transpose2d_s2_merged(x_split, w_merged, k, co, out=None, cin=None, split_k=0):
    """F.conv_transpose2d(stride=2, padding=0) for a k x k kernel in ONE launch; w_merged from
    merged_phase_weights(w[Co, Ci, k*k], *_phase_plan('convT', k, k, 2, 0))."""
    n, h, w_, _, _ = x_split.shape
    oh, ow = 2 * (h - 1) + k, 2 * (w_ - 1) + k
    shifts, idx, G = _phase_plan('convT', k, k, 2, 0, x_split.device)
    if out is None:
        out = torch.empty(n, oh, ow, co, dtype=torch.float32, device=x_split.device)
    taps = [(sy, sx, i) for i, (sy, sx) in enumerate(shifts)]
    conv_taps(x_split, w_merged, taps, out, grid=((oh + 1) // 2, (ow + 1) // 2), out_step=(2, 2), cout=G * co,
              groups=(co, 2), algo_macs_per_pixel=_real_blocks(idx, k * k) * co * (cin or x_split.shape[3] * 32), cin=cin,
              split_k=split_k)
    return out
