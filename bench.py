#!/usr/bin/env python
"""bench.py — throughput of the WarpedGANSpace hot path on B200s, one JSON line per run.

  python bench.py --gpus N --steps K --warmup W [--config c3]      product arm (libwgs_b200)
  python bench.py --impl reference --gpus N --steps K --warmup W   reference arm: the reference algorithm on the HOST cores
                                                                   (oracle port: /root/reference is not on the GPU box)
  python bench.py --impl reference-gpu [--config ..]               the reference ALGORITHM on the GPU through torch / cuDNN
                                                                   (oracle functions on cuda tensors; TF32 on and off) - the
                                                                   library path the hand-written kernels have to beat

Configs (BASELINE.json `configs`):
  c3 (default, the headline metric)  StyleGAN2-1024, K=128 D=32, ResNet-18 R @1024^2, 4 latents per GPU  -> image-pairs/s
  c2                                 ProgGAN-1024,  K=128 D=32, ResNet-18 R @1024^2, 8 latents per GPU   -> image-pairs/s
  c4                                 BigGAN-128 (the reference's BigGAN; "-deep 256" does not exist in it, SURVEY mismatch 1),
                                     K=120 D=256, ResNet-18 R @128^2, 8 latents per GPU                  -> image-pairs/s
  c5                                 latent traversal, StyleGAN2-1024, generator only: per step `chains` (z, path) chains of
                                     33 frames: RBF chains in one launch + batched inference             -> images/s

A training "step" = one paired step on one batch of synthetic latents: RBF warp, G(z) and G(z+dz), Reconstructor forward,
CE + L1 loss, full backward (generator data-gradient only), both Adam updates.  Weak scaling: every GPU gets
`--batch-per-gpu` latents; the only collectives are the all-reduces of the flat S / R gradients (R's overlapped with the
generator's data-gradient pass).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

# forward FLOPs per image (F_G) / per pair (F_R), SURVEY.md Appendix A
CONFIGS = {
    'c3': dict(metric='image-pairs/sec StyleGAN2-FFHQ-1024 K=128 train step', unit='pairs/s', gan='StyleGAN2', K=128, D=32, d=512,
               batch=4, size=1024, F_G=148.5e9, F_R=80.7e9,
               workload='StyleGAN2-1024 paired step, K=128 D=32 d=512, ResNet-18 R @1024^2, Z-space shift'),
    'c2': dict(metric='image-pairs/sec ProgGAN-CelebAHQ-1024 K=128 train step', unit='pairs/s', gan='ProgGAN', K=128, D=32, d=512,
               batch=8, size=1024, F_G=85.5e9, F_R=80.7e9,
               workload='ProgGAN-1024 paired step, K=128 D=32 d=512, ResNet-18 R @1024^2'),
    'c4': dict(metric='image-pairs/sec BigGAN-ImageNet-128 K=120 train step', unit='pairs/s', gan='BigGAN', K=120, D=256, d=120,
               batch=8, size=128, F_G=42.7e9, F_R=1.26e9,
               workload='BigGAN-128 (class 239) paired step, K=120 D=256 d=120, ResNet-18 R @128^2'),
    'c5': dict(metric='images/sec StyleGAN2-FFHQ-1024 latent traversal (generator only)', unit='images/s', gan='StyleGAN2', K=128,
               D=32, d=512, batch=8, size=1024, F_G=148.5e9, F_R=0.0,
               workload='StyleGAN2-1024 traversal: per step `batch_per_gpu` (z, path) chains x 33 frames (16 RBF steps each way, '
                        'eps 0.15), frames rendered in batches of 33'),
}
SHIFT_STEPS, EPS = 16, 0.15


def algo_flops_per_unit(cfg):
    """Training: 2 G forwards + 1 G data-gradient + R fwd / dgrad / wgrad per pair; traversal: one G forward per image."""
    return cfg['F_G'] if cfg is CONFIGS['c5'] else 3 * cfg['F_G'] + 3 * cfg['F_R']


CONV_SOURCES = ('warpedganspace_b200/csrc/conv.cu', 'warpedganspace_b200/csrc/ptx.cuh', 'warpedganspace_b200/csrc/common.cuh',
                'include/wgs_b200.h')


def source_stamp():
    """sha1 over the sources of the tensor-core conv family (the kernels whose DRAM traffic the capture measures): ties a
    committed ncu traffic capture to the kernels it measured."""
    import hashlib
    h = hashlib.sha1()
    for rel in CONV_SOURCES:
        with open(os.path.join(ROOT, rel), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def recorded_conv_traffic():
    """Average DRAM bytes (read + write) per tensor-core conv launch of one training step from the newest committed ncu
    capture (profiles/r*_conv_traffic.json, tools/ncu_traffic.py).  The capture is stamped with the hash of the CUDA
    sources it measured; a stale capture is reported as such instead of being passed off as current."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_conv_traffic.json')))
    if not paths:
        return None
    try:
        with open(paths[-1]) as f:
            t = json.load(f)
    except Exception:
        return None
    t['file'] = os.path.relpath(paths[-1], ROOT)
    t['stale'] = t.get('source_stamp') != source_stamp()
    return t


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p.get('hbm_gbs', 6650.0), bf16_tflops=p.get('bf16_tflops_sustained', p.get('bf16_tflops', 1400.0)),
                    source='measured (MEASURED_PEAKS.json, sustained bf16)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons (NVML, every 10 ms) while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.max_mhz = None
        self.active = False          # samples are kept only while the timed region runs (NVML init happens before it)
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.ready.set()
            while not self.stop_flag:
                if not self.active:
                    time.sleep(0.002)
                    continue
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((mhz, reasons))
                time.sleep(0.01)
        except Exception:
            # fall back to nvidia-smi (slow: a few samples only)
            q = 'clocks.sm,clocks.max.sm'
            self.ready.set()
            while not self.stop_flag:
                if not self.active:
                    time.sleep(0.002)
                    continue
                try:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                    a, b = [float(x) for x in out.strip().split(',')[:2]]
                    self.max_mhz = b
                    self.samples.append((a, 0))
                except Exception:
                    pass
                time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unsampled']}
        mhz = sorted(s[0] for s in self.samples)
        bits = 0
        for _, r in self.samples:
            bits |= r
        names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
        return {'sm_mhz': mhz[len(mhz) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': [n for b, n in names.items() if bits & b],
                'samples': len(self.samples)}


# ---- product arm -----------------------------------------------------------------------------------------------------------
def build_generator(cfg, device):
    """Random-init generator of the config's architecture behind its reference wrapper (identical on every rank)."""
    import math
    from warpedganspace_b200 import gan_load
    torch.manual_seed(0)
    if cfg['gan'] == 'StyleGAN2':
        from warpedganspace_b200.stylegan2 import Generator
        G = Generator(cfg['size'], 512, 8)
        with torch.no_grad():
            for name, p in G.named_parameters():
                if name.endswith('noise.weight'):
                    p.fill_(0.1)                                # exercise the noise path (the reference's init is 0)
        return gan_load.StyleGAN2Wrapper(G, shift_in_w_space=False).to(device)
    if cfg['gan'] == 'ProgGAN':
        from warpedganspace_b200.generators import ProgGANGenerator
        G = ProgGANGenerator()
        with torch.no_grad():                                   # statistics of the released model (the constructor's
            for name, p in G.named_parameters():                # own init is bias-dominated and hides conv work / errors)
                if name.endswith('conv.weight'):
                    fan_in = p.shape[1] * p.shape[2] * p.shape[3]
                    p.normal_()
                    scale = dict(G.named_parameters())[name.replace('conv.weight', 'wscale.scale')]
                    scale.fill_((1.0 if name.startswith('output') else math.sqrt(2.0)) / math.sqrt(fan_in))
                elif name.endswith('wscale.b'):
                    p.zero_()
        return gan_load.ProgGANWrapper(G).to(device)
    if cfg['gan'] == 'BigGAN':
        from warpedganspace_b200.generators import BigGANGenerator
        G = BigGANGenerator(resolution=cfg['size'])
        with torch.no_grad():
            for name, b in G.named_buffers():                   # non-trivial eval-mode BatchNorm statistics
                if name.endswith('stored_var'):
                    b.uniform_(0.5, 1.5)
                elif name.endswith('stored_mean'):
                    b.normal_(0.0, 0.1)
        return gan_load.BigGANWrapper(G, target_classes=(239,)).to(device)
    raise ValueError(cfg['gan'])


def build_product(device, batch, cfg=None):
    from warpedganspace_b200 import SupportSets
    from warpedganspace_b200.reconstructor import Reconstructor
    from warpedganspace_b200.trainer import PairedTrainer
    cfg = cfg or CONFIGS['c3']
    W = build_generator(cfg, device)
    S = SupportSets(cfg['K'], cfg['D'], cfg['d'], learn_alphas=False, learn_gammas=True, gamma=1.0 / cfg['d'])
    R = Reconstructor('ResNet', cfg['K'], 3)
    return PairedTrainer(W, S.to(device), R.to(device))


def make_batches(n, batch, device, seed, pinned=False, cfg=None):
    from warpedganspace_b200.trainer import sample_shift_magnitudes
    cfg = cfg or CONFIGS['c3']
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        z = torch.randn(batch, cfg['d'], generator=g)
        idx = torch.randint(0, cfg['K'], (batch,), generator=g)
        mag = sample_shift_magnitudes(batch, 0.1, 0.2, 'cpu', generator=g)
        if pinned:
            out.append((z.pin_memory(), idx.pin_memory(), mag.pin_memory()))
        else:
            out.append((z.to(device), idx.to(device), mag.to(device)))
    return out


# ---- the reference algorithm (oracle port): CPU legs and the GPU library arm ---------------------------------------------
def oracle_problem(cfg, device='cpu', seed=0):
    """State dicts + step closure of the oracle (CPU restatement of the reference) for a config, on `device`."""
    import oracle.support_sets as o_ss
    import oracle.stylegan2 as o_sg2
    import oracle.proggan as o_pg
    import oracle.biggan as o_bg
    import oracle.reconstructor as o_rec
    import oracle.step as o_step
    g = torch.Generator().manual_seed(seed)
    if cfg['gan'] == 'StyleGAN2':
        g_sd = o_sg2.init_state(size=cfg['size'], generator=g)
        kw = dict(size=cfg['size'])
    elif cfg['gan'] == 'ProgGAN':
        g_sd = o_pg.init_state(generator=g, pretrained_like=True)
        kw = {}
    else:
        g_sd = o_bg.init_state(cfg['size'], generator=g)
        kw = dict(resolution=cfg['size'])
    s_sd = o_ss.init_state(cfg['K'], cfg['D'], cfg['d'], generator=g)
    r_sd = o_rec.init_state('ResNet', cfg['K'], 3, generator=g)
    mv = lambda sd: {k: v.to(device) for k, v in sd.items()}
    g_sd, s_sd, r_sd = mv(g_sd), mv(s_sd), mv(r_sd)

    def step(batch, g_requires_grad=False):
        z = torch.randn(batch, cfg['d'], generator=g).to(device)
        idx = torch.randint(0, cfg['K'], (batch,), generator=g).to(device)
        mag = o_step.sample_shift_magnitudes(batch, 0.1, 0.2, generator=g).to(device)
        if cfg['gan'] == 'BigGAN':
            kw['classes'] = torch.full((batch,), 239, device=device)
        gs = g_sd
        if g_requires_grad:        # the reference leaves the frozen generator's weights trainable and discards the result
            gs = {k: (v.detach().requires_grad_(True) if v.is_floating_point() else v) for k, v in g_sd.items()}
        gen_fn, _ = o_step.make_generator(cfg['gan'], gs, **kw)
        res = o_step.paired_step(gen_fn, s_sd, r_sd, z, idx, mag, reconstructor_type='ResNet')
        for k, gr in res['grads']['S'].items():
            o_step.adam_update(s_sd[k], gr, torch.zeros_like(gr), torch.zeros_like(gr), 1)
        for k, gr in res['grads']['R'].items():
            o_step.adam_update(r_sd[k], gr, torch.zeros_like(gr), torch.zeros_like(gr), 1)
        return res['loss']

    def frames(n):
        """Traversal (config 5): one chain of RBF steps + n generator forwards."""
        z0 = torch.randn(1, cfg['d'], generator=g).to(device)
        codes, shifts = o_step.traverse_chain(s_sd, z0, 3, EPS, SHIFT_STEPS)
        with torch.no_grad():
            return o_sg2.generate(g_sd, codes[:n], shifts[:n], size=cfg['size'])

    return step, frames


def oracle_step_timer(cfg, batch, max_seconds, warmup, steps):
    """Times the oracle on the host cores; returns (units_per_s, ms_per_step, steps_run, threads, sample text)."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step, frames = oracle_problem(cfg)
    t_all = time.time()

    def one():
        t0 = time.time()
        if cfg is CONFIGS['c5']:
            frames(batch)
        else:
            step(batch)
        return time.time() - t0

    times = []
    for _ in range(warmup):
        one()
        if time.time() - t_all > max_seconds * 0.4:
            break
    for _ in range(steps):
        times.append(one())
        if time.time() - t_all > max_seconds:
            break
    ms = 1e3 * sum(times) / len(times)
    what = ('%d frame(s) of one traversal chain' % batch) if cfg is CONFIGS['c5'] else ('%d pair(s)' % batch)
    return batch / (ms / 1e3), ms, len(times), threads, what


def run_reference(args, cfg, world, rank):
    if rank != 0:
        return
    batch = 1
    ups, ms, ran, threads, what = oracle_step_timer(cfg, batch, args.reference_seconds, min(args.warmup, 1), args.steps)
    line = {
        'impl': 'reference', 'metric': cfg['metric'], 'value': ups, 'unit': cfg['unit'], 'n_gpus': args.gpus, 'steps': ran,
        'warmup': min(args.warmup, 1), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'units_per_step': batch, 'steps_requested': args.steps,
                   'note': 'bounded sample: %s per step on the host CPU, time-capped; the oracle does NOT form the generator '
                           'weight-gradients the real reference computes and discards (lib/trainer.py:190,250), so this arm is '
                           'faster than the reference itself would be' % what},
        'cpu_baseline': {'value': ups, 'unit': cfg['unit'], 'cores': threads, 'kind': 'port',
                         'sample': '%d step(s) of %s, oracle port of the reference (reference tree not on the box; '
                                   'StyleGAN2 has no CPU path in the reference itself)' % (ran, what)},
        'e2e': {'value': ups, 'unit': cfg['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu(args, cfg, rank):
    """The reference ALGORITHM on one GPU through torch / cuDNN: the oracle's functional modules on cuda tensors (F.conv2d
    with groups = batch for StyleGAN2's modulated convs, exactly the reference's formulation, models/StyleGAN2/model.py:
    202-226), cuDNN's default TF32 convolutions and with TF32 off, with and without the generator weight-gradients the
    reference computes and throws away.  Precision note: TF32 misses the 1e-3 bar at 1024^2 (profiles/r01_precision_study.md)."""
    if rank != 0:
        return
    if cfg is CONFIGS['c5']:
        raise SystemExit('--impl reference-gpu covers the training configs (c2, c3, c4)')
    device = torch.device('cuda', 0)
    B = args.batch_per_gpu or cfg['batch']
    out = {}
    for tf32 in (True, False):
        for wasted in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True                    # lib/trainer.py:163
            step, _ = oracle_problem(cfg, device=device)
            try:
                for _ in range(max(2, args.warmup)):
                    step(B, wasted)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    step(B, wasted)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                out['tf32_%s%s' % ('on' if tf32 else 'off', '_with_discarded_generator_wgrads' if wasted else '')] = \
                    {'value': B / (ms / 1e3), 'ms_per_step': ms}
            except RuntimeError as e:                                # e.g. out of memory with the wasted weight-gradients
                out['tf32_%s%s' % ('on' if tf32 else 'off', '_with_discarded_generator_wgrads' if wasted else '')] = \
                    {'error': str(e).splitlines()[0][:200]}
                torch.cuda.empty_cache()
    best = out.get('tf32_on', {})
    line = {'impl': 'reference-gpu', 'metric': cfg['metric'], 'value': best.get('value'), 'unit': cfg['unit'], 'n_gpus': 1,
            'steps': args.steps, 'warmup': max(2, args.warmup), 'ms_per_step': best.get('ms_per_step'), 'higher_is_better': True,
            'dtype': 'tf32 (cuDNN default) / f32', 'data': 'synthetic',
            'config': {'workload': cfg['workload'], 'batch_per_gpu': B,
                       'note': 'oracle port of the reference step on cuda through torch/cuDNN (library path), eager'},
            'variants': out}
    print(json.dumps(line), flush=True)


# ---- config 5: traversal ---------------------------------------------------------------------------------------------------
def run_traversal(args, cfg, world, rank, local, device):
    from warpedganspace_b200 import SupportSets, _lib, conv as C, dist as wdist
    from warpedganspace_b200.traversal import traverse_paths
    from warpedganspace_b200.image_out import images_to_uint8
    G = build_generator(cfg, device).eval()
    S = SupportSets(cfg['K'], cfg['D'], cfg['d'], learn_alphas=False, learn_gammas=True, gamma=1.0 / cfg['d']).to(device)
    chains = args.batch_per_gpu or cfg['batch']
    frames_per = 2 * SHIFT_STEPS + 1
    total = args.warmup + args.steps
    g = torch.Generator().manual_seed(3000 + rank)
    zs = [torch.randn(1, cfg['d'], generator=g) for _ in range(total)]
    paths = list(range(chains))

    copy_stream = torch.cuda.Stream(device=device)

    def step(z, host_out=None):
        """host_out (pinned uint8 [images, H, W, 3]): every batch of frames is converted on the device and copied out on a second
        stream as soon as it exists, underneath the rendering of the next batch; the step ends when the last copy has landed."""
        keep, done = [], 0

        def to_host(lo, hi, img):
            nonlocal done
            u8 = images_to_uint8(img, adaptive=True)
            ready = torch.cuda.Event()
            ready.record()
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                host_out[done: done + u8.shape[0]].copy_(u8, non_blocking=True)
            keep.append(u8)
            done += u8.shape[0]

        traverse_paths(G, S, z, paths=paths, eps=EPS, shift_steps=SHIFT_STEPS, batch_size=frames_per, return_images=False,
                       on_frames=to_host if host_out is not None else (lambda lo, hi, img: None))
        if host_out is not None:
            assert done == host_out.shape[0], (done, host_out.shape)
            torch.cuda.current_stream().wait_stream(copy_stream)
            return keep

    sampler = ClockSampler(local)
    sampler.start()
    dev_z = [z.to(device) for z in zs]
    for i in range(args.warmup):
        step(dev_z[i])
    torch.cuda.synchronize()
    sampler.ready.wait(timeout=10.0)
    wdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    e0.record()
    for i in range(args.warmup, total):
        step(dev_z[i])
    e1.record()
    torch.cuda.synchronize()
    sampler.active = False
    wdist.barrier()
    ms_step = wdist.max_over_ranks(e0.elapsed_time(e1), device) / args.steps
    imgs = chains * frames_per
    value = world * imgs / (ms_step / 1e3)
    # e2e: pinned-host z in, uint8 frames (the reference's tensor2image pixels) back in pinned host memory every step
    pinned = [z.pin_memory() for z in zs]
    host_out = torch.empty(imgs, cfg['size'], cfg['size'], 3, dtype=torch.uint8).pin_memory()
    for i in range(args.warmup):
        step(pinned[i].to(device, non_blocking=True), host_out)
    torch.cuda.synchronize()
    wdist.barrier()
    e0.record()
    for i in range(args.warmup, total):
        step(pinned[i].to(device, non_blocking=True), host_out)
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    wdist.barrier()
    e2e_ms = wdist.max_over_ranks(e0.elapsed_time(e1), device) / args.steps
    # per-launch profile of one more step
    C.PROFILE = []
    _lib.reset_launch_count()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    step(dev_z[0])
    p1.record()
    torch.cuda.synchronize()
    prof, C.PROFILE = C.PROFILE, None
    launches = _lib.launch_count()
    sampler.stop_flag = True
    if rank != 0:
        return
    roofline = conv_roofline(prof, p0.elapsed_time(p1), 1, cfg, imgs, ms_step)
    line = base_line(args, cfg, world, value, ms_step, chains, sampler, roofline, launches, graphed=False)
    line['e2e'] = {'value': world * imgs / (e2e_ms / 1e3), 'unit': cfg['unit'], 'ms_per_step': e2e_ms,
                   'h2d_bytes_per_step': cfg['d'] * 4, 'd2h_bytes_per_step': imgs * cfg['size'] * cfg['size'] * 3}
    line['config'].update(chains_per_step=chains, frames_per_chain=frames_per, images_per_step=imgs,
                          output='uint8 RGB frames (device-side tensor2image), JPEG encode on the host is outside the timed region')
    if world == 1 and not args.no_cpu_baseline:
        ups, ms, ran, threads, what = oracle_step_timer(cfg, 2, args.cpu_baseline_seconds, 0, 2)
        line['cpu_baseline'] = {'value': ups, 'unit': cfg['unit'], 'cores': threads, 'kind': 'port',
                                'sample': '%d run(s) of %s, %.1f s each' % (ran, what, ms / 1e3)}
    print(json.dumps(line), flush=True)


# ---- roofline bookkeeping --------------------------------------------------------------------------------------------------
def conv_roofline(prof, prof_ms_total, n_prof, cfg, units_per_step, ms_step):
    pk = peaks()
    traffic = recorded_conv_traffic()
    conv = [r for r in prof if r[0] == 'conv']
    wg = [r for r in prof if r[0] == 'wgrad']
    t_of = lambda r: r[2].elapsed_time(r[3])
    conv_ms = sum(t_of(r) for r in conv)
    conv_fl = sum(r[1] for r in conv)
    issued_fl = sum(r[5] for r in conv)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    layers = {}
    for r in conv:
        e = layers.setdefault(r[6], [0, 0.0, 0.0, 0.0])
        e[0] += 1; e[1] += t_of(r); e[2] += r[1]; e[3] += r[5]
    table = [{'layer': k, 'launches_per_step': v[0] / n_prof, 'us_per_launch': round(1e3 * v[1] / v[0], 1),
              'algorithmic_tflops': round(v[2] / (v[1] * 1e-3) / 1e12, 1), 'contracted_tflops': round(v[3] / (v[1] * 1e-3) / 1e12, 1),
              'ms_per_step': round(v[1] / n_prof, 3)}
             for k, v in sorted(layers.items(), key=lambda kv: -kv[1][1])]
    stale = bool(traffic and traffic.get('stale'))
    return {
        'bound': 'tensor', 'kernel': 'wgs::conv_tc_kernel family (conv_tc / conv_halo / conv_halo_mt: one entry point, wgs_conv_split32)',
        'achieved': achieved, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['bf16_tflops'],
        'traffic': None if (not traffic or stale) else traffic.get('dram_bytes_per_launch'),
        'traffic_source': (traffic or {}).get('source'), 'traffic_file': (traffic or {}).get('file'), 'traffic_stale': stale,
        'algorithmic_bytes_per_launch': sum(r[4] for r in conv) / max(1, len(conv)),
        'peak_source': pk['source'],
        'accounting': 'achieved = ALGORITHMIC FLOPs (2 x output pixels x real taps x true Cin x Cout; no zero blocks of '
                      'phase-packed launches, no channel padding) / CUDA-event time of every launch; contracted = the same '
                      'with zero blocks and padding (contracted/algorithmic = %.3f)' % (issued_fl / conv_fl if conv_fl else 0.0),
        'precision': 'fp32-accurate 3xbf16 split: every contracted MAC issues 3 bf16 MMAs (issued bf16 / peak = %.3f)'
                     % (3 * issued_fl / (conv_ms * 1e-3) / 1e12 / pk['bf16_tflops'] if conv_ms else 0.0),
        'launches': len(conv), 'avg_launch_ms': conv_ms / max(1, len(conv)),
        'share_of_step': conv_ms / prof_ms_total if prof_ms_total else None,
        'measured_over': '%d instrumented eager step(s) on ONE stream (CUDA events around every launch, side streams off so that no launch shares the SMs; %.2f ms/step eager)' % (n_prof, prof_ms_total / n_prof),
        'wgrad_kernel': {'achieved': (sum(r[1] for r in wg) / (sum(t_of(r) for r in wg) * 1e-3) / 1e12) if wg else None,
                         'launches': len(wg), 'share_of_step': sum(t_of(r) for r in wg) / prof_ms_total if prof_ms_total and wg else None},
        'whole_step_algorithmic_tflops': algo_flops_per_unit(cfg) * units_per_step / (ms_step * 1e-3) / 1e12,
        'per_layer': table[:40],
    }


def base_line(args, cfg, world, value, ms_step, per_gpu, sampler, roofline, launches, graphed):
    return {
        'metric': cfg['metric'], 'value': value, 'unit': cfg['unit'], 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16x3 (fp32 operands split hi+lo, fp32 accumulate in TMEM)', 'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'name': args.config, 'batch_per_gpu': per_gpu, 'global_batch': world * per_gpu,
                   'parallelism': 'latents sharded dp%d%s' % (world, '' if cfg is CONFIGS['c5'] else ', S and R grad all-reduce (R overlapped)'),
                   'weights': 'random init (reference constructors; StyleGAN2 noise strength 0.1, ProgGAN released-model statistics)',
                   'l2': 'working set per step (GBs of activations) exceeds the 126 MB L2; no flush needed'},
        'clocks': sampler.summary(), 'gpu_launches': int(launches), 'roofline': roofline, 'cuda_graph': bool(graphed),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='product', choices=['product', 'reference', 'reference-gpu'])
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS))
    ap.add_argument('--batch-per-gpu', type=int, default=0, help='latents (c5: chains) per GPU per step; 0 = the config default')
    ap.add_argument('--reference-seconds', type=float, default=150.0)
    ap.add_argument('--cpu-baseline-seconds', type=float, default=40.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run the step eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'product' else args.warmup
    cfg = CONFIGS[args.config]

    from warpedganspace_b200 import dist as wdist
    world, rank, local = wdist.env_world()
    if args.impl == 'reference':
        run_reference(args, cfg, world, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py (%s arm) needs a CUDA device: libwgs_b200 has no CPU fallback' % args.impl)
    if args.impl == 'reference-gpu':
        run_reference_gpu(args, cfg, rank)
        return
    world, rank, local = wdist.init_from_env('nccl' if world > 1 else None)
    device = torch.device('cuda', local)
    from warpedganspace_b200 import _lib, conv as C
    _lib.call('wgs_device_info', None, None)
    if cfg is CONFIGS['c5']:
        run_traversal(args, cfg, world, rank, local, device)
        return
    B = args.batch_per_gpu or cfg['batch']
    trainer = build_product(device, B, cfg)
    total = args.warmup + args.steps
    batches = make_batches(total, B, device, seed=1000 + rank, cfg=cfg)

    # ---- device-resident run: `value` -----------------------------------------------------------------
    # (capture first: a CUDA graph must be captured before any eager backward pass, see PairedTrainer.capture)
    graphed = False
    if not args.no_graph:
        graphed = trainer.capture(*batches[0])
        if not graphed and rank == 0:
            print('CUDA-graph capture failed, running eagerly: %s' % getattr(trainer, 'capture_error', '?'), file=sys.stderr)
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(args.warmup):
        trainer.step(*batches[i])
    torch.cuda.synchronize()
    sampler.ready.wait(timeout=10.0)
    wdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.active = True
    e0.record()
    for i in range(args.warmup, total):
        trainer.step(*batches[i])
    e1.record()
    torch.cuda.synchronize()
    sampler.active = False
    wdist.barrier()
    ms_total = wdist.max_over_ranks(e0.elapsed_time(e1), device)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- end-to-end run through the public API with host buffers: `e2e` ---------------------------------------
    host = make_batches(total, B, device, seed=2000 + rank, pinned=True, cfg=cfg)

    def e2e_step(hb):
        z, idx, mag = (t.to(device, non_blocking=True) for t in hb)
        out = trainer.step(z, idx, mag)
        return float(out['loss'].item())                       # device -> host read of the step's result
    for i in range(args.warmup):
        e2e_step(host[i])
    torch.cuda.synchronize()
    wdist.barrier()
    sampler.active = True
    e0.record()
    for i in range(args.warmup, total):
        e2e_step(host[i])
    e1.record()
    torch.cuda.synchronize()
    sampler.active = False
    wdist.barrier()
    sampler.stop_flag = True
    e2e_ms = wdist.max_over_ranks(e0.elapsed_time(e1), device) / args.steps
    e2e = {'value': world * B / (e2e_ms / 1e3), 'unit': cfg['unit'], 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': B * (cfg['d'] * 4 + 8 + 4), 'd2h_bytes_per_step': 4}

    # ---- instrumented eager steps: per-launch CUDA-event timing of the tensor-core kernels (roofline) -----------------
    # (events cannot be recorded inside a CUDA graph, so the per-launch numbers come from eager steps of the same
    #  workload run right after the timed regions)
    # The side streams are switched off for these steps: a per-launch event pair must bracket ONE kernel running alone, not a
    # conv sharing the SMs with a weight-gradient kernel of the side stream (that is what the timed regions above measure).
    side_env = os.environ.get('WGS_SIDE_STREAMS')
    os.environ['WGS_SIDE_STREAMS'] = '0'
    for i in range(2):
        trainer.step(*batches[i], eager=True)
    torch.cuda.synchronize()
    C.PROFILE = []
    _lib.reset_launch_count()
    n_prof = min(args.steps, 5)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(args.warmup, args.warmup + n_prof):
        trainer.step(*batches[i], eager=True)
    p1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() / n_prof
    prof, C.PROFILE = C.PROFILE, None
    if side_env is None:
        os.environ.pop('WGS_SIDE_STREAMS', None)
    else:
        os.environ['WGS_SIDE_STREAMS'] = side_env
    if rank != 0:
        return
    roofline = conv_roofline(prof, p0.elapsed_time(p1), n_prof, cfg, B, ms_step)
    line = base_line(args, cfg, world, value, ms_step, B, sampler, roofline, launches, graphed)
    line['e2e'] = e2e
    if world == 1 and not args.no_cpu_baseline:
        ups, ms, ran, threads, what = oracle_step_timer(cfg, 1, args.cpu_baseline_seconds, 0, 2)
        line['cpu_baseline'] = {'value': ups, 'unit': cfg['unit'], 'cores': threads, 'kind': 'port',
                                'sample': '%d step(s) of %s (same workload, batch 1), %.1f s/step' % (ran, what, ms / 1e3)}
    print(json.dumps(line), flush=True)


def _shutdown():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    try:
        main()
    finally:
        _shutdown()
