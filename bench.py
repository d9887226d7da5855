#!/usr/bin/env python
"""bench.py — image-pairs/sec of the WarpedGANSpace paired training step (StyleGAN2-1024, K=128).

  python bench.py --gpus N --steps K --warmup W            product arm (libwgs_b200 on B200s)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference algorithm on the host
                                                           cores (oracle port: /root/reference is not on the box)

A "step" = one paired training step on one batch of synthetic latents: RBF warp, G(z) and G(z+dz), Reconstructor
forward, CE + L1 loss, full backward (generator data-gradient only), both Adam updates.  Weak scaling: every GPU
gets `--batch-per-gpu` latents (BASELINE config 3: batch 32 over 8 GPUs = 4 per GPU); the only collective is one
all-reduce of the flat S / R gradients.  One JSON line is printed by rank 0.
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

METRIC = 'image-pairs/sec StyleGAN2-FFHQ-1024 K=128 train step'
K_SETS, DIPOLES, DIM = 128, 32, 512
F_G, F_R = 148.5e9, 80.7e9                       # forward FLOPs per image / per pair
ALGO_FLOPS_PER_PAIR = 3 * F_G + 3 * F_R          # 2 G forwards + 1 G data-gradient + R fwd/dgrad/wgrad


def recorded_conv_traffic():
    """Average DRAM bytes (read + write) per tensor-core conv launch of one training step, from the committed ncu
    capture of this workload (profiles/r01_conv_traffic.json, written by tools/ncu_traffic.py); None if absent."""
    path = os.path.join(ROOT, 'profiles', 'r01_conv_traffic.json')
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p.get('hbm_gbs', 6650.0), bf16_tflops=p.get('bf16_tflops_sustained', p.get('bf16_tflops', 1400.0)),
                    source='measured (MEASURED_PEAKS.json, sustained bf16)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons (NVML, every 20 ms) while the timed region runs."""

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


def build_product(device, batch):
    from warpedganspace_b200 import SupportSets
    from warpedganspace_b200.stylegan2 import Generator
    from warpedganspace_b200.gan_load import StyleGAN2Wrapper
    from warpedganspace_b200.reconstructor import Reconstructor
    from warpedganspace_b200.trainer import PairedTrainer
    torch.manual_seed(0)                                        # identical random-init weights on every rank
    G = Generator(1024, 512, 8)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith('noise.weight'):
                p.fill_(0.1)                                    # exercise the noise path (reference init is 0)
    S = SupportSets(K_SETS, DIPOLES, DIM, learn_alphas=False, learn_gammas=True, gamma=1.0 / DIM)
    R = Reconstructor('ResNet', K_SETS, 3)
    W = StyleGAN2Wrapper(G, shift_in_w_space=False).to(device)
    return PairedTrainer(W, S.to(device), R.to(device))


def make_batches(n, batch, device, seed, pinned=False):
    from warpedganspace_b200.trainer import sample_shift_magnitudes
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        z = torch.randn(batch, DIM, generator=g)
        idx = torch.randint(0, K_SETS, (batch,), generator=g)
        mag = sample_shift_magnitudes(batch, 0.1, 0.2, 'cpu', generator=g)
        if pinned:
            out.append((z.pin_memory(), idx.pin_memory(), mag.pin_memory()))
        else:
            out.append((z.to(device), idx.to(device), mag.to(device)))
    return out


def oracle_step_timer(batch, max_seconds, warmup, steps):
    """Times the oracle (CPU restatement of the reference) on the same workload; returns
    (pairs_per_s, ms_per_step, steps_run, threads)."""
    import oracle.support_sets as o_ss
    import oracle.stylegan2 as o_sg2
    import oracle.reconstructor as o_rec
    import oracle.step as o_step
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    g_sd = o_sg2.init_state(size=1024, generator=g)
    s_sd = o_ss.init_state(K_SETS, DIPOLES, DIM, generator=g)
    r_sd = o_rec.init_state('ResNet', K_SETS, 3, generator=g)
    gen_fn, _ = o_step.make_generator('StyleGAN2', g_sd, size=1024)
    t_all = time.time()

    def one():
        z = torch.randn(batch, DIM, generator=g)
        idx = torch.randint(0, K_SETS, (batch,), generator=g)
        mag = o_step.sample_shift_magnitudes(batch, 0.1, 0.2, generator=g)
        t0 = time.time()
        res = o_step.paired_step(gen_fn, s_sd, r_sd, z, idx, mag, reconstructor_type='ResNet')
        m = {k: torch.zeros_like(v) for k, v in res['grads']['S'].items()}
        for k, gr in res['grads']['S'].items():
            o_step.adam_update(s_sd[k], gr, m[k], torch.zeros_like(gr), 1)
        for k, gr in res['grads']['R'].items():
            o_step.adam_update(r_sd[k], gr, torch.zeros_like(gr), torch.zeros_like(gr), 1)
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
    return batch / (ms / 1e3), ms, len(times), threads


def run_reference(args, world, rank):
    if rank != 0:
        return
    batch = 1
    pps, ms, ran, threads = oracle_step_timer(batch, args.reference_seconds, min(args.warmup, 1), args.steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': pps, 'unit': 'pairs/s', 'n_gpus': args.gpus, 'steps': ran,
        'warmup': min(args.warmup, 1), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'StyleGAN2-1024 paired step, K=128 D=32 d=512, ResNet-18 R @1024^2', 'pairs_per_step': batch,
                   'steps_requested': args.steps, 'note': 'bounded sample: 1 pair per step on the host CPU, time-capped'},
        'cpu_baseline': {'value': pps, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d step(s) of 1 pair, oracle port of the reference (reference tree not on the box; '
                                   'StyleGAN2 has no CPU path in the reference itself)' % ran},
        'e2e': {'value': pps, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='product', choices=['product', 'reference'])
    ap.add_argument('--batch-per-gpu', type=int, default=4)
    ap.add_argument('--reference-seconds', type=float, default=150.0)
    ap.add_argument('--cpu-baseline-seconds', type=float, default=40.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run the step eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'product' else args.warmup

    from warpedganspace_b200 import dist as wdist
    world, rank, local = wdist.env_world()
    if args.impl == 'reference':
        run_reference(args, world, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py (product arm) needs a CUDA device: libwgs_b200 has no CPU fallback')
    world, rank, local = wdist.init_from_env('nccl' if world > 1 else None)
    device = torch.device('cuda', local)
    from warpedganspace_b200 import _lib, conv as C
    _lib.call('wgs_device_info', None, None)
    B = args.batch_per_gpu
    trainer = build_product(device, B)
    total = args.warmup + args.steps
    batches = make_batches(total, B, device, seed=1000 + rank)

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
    host = make_batches(total, B, device, seed=2000 + rank, pinned=True)
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
    e2e = {'value': world * B / (e2e_ms / 1e3), 'unit': 'pairs/s', 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': B * (DIM * 4 + 8 + 4), 'd2h_bytes_per_step': 4}

    # ---- instrumented eager steps: per-launch CUDA-event timing of the tensor-core kernels (roofline) -----------------
    # (events cannot be recorded inside a CUDA graph, so the per-launch numbers come from eager steps of the same
    #  workload run right after the timed regions)
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
    prof_ms_total = p0.elapsed_time(p1)

    # ---- roofline of the dominant kernel (tensor-core conv), per launch, from the same timed region -------
    pk = peaks()
    traffic = recorded_conv_traffic()
    conv = [(r[1], r[2].elapsed_time(r[3])) for r in prof if r[0] == 'conv']
    conv_bytes = sum(r[4] for r in prof if r[0] == 'conv' and len(r) > 4)
    wg = [(r[1], r[2].elapsed_time(r[3])) for r in prof if r[0] == 'wgrad']
    conv_ms = sum(t for _, t in conv)
    conv_fl = sum(f for f, _ in conv)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    roofline = {
        'bound': 'tensor', 'kernel': 'wgs::conv_tc_kernel family (conv_tc / conv_halo / conv_halo_mt: one entry point, wgs_conv_split32)',
        'achieved': achieved, 'peak': pk['bf16_tflops'],
        'unit': 'TFLOP/s', 'frac': achieved / pk['bf16_tflops'],
        'traffic': (traffic or {}).get('dram_bytes_per_launch'), 'traffic_source': (traffic or {}).get('source'),
        'algorithmic_bytes_per_launch': conv_bytes / max(1, len(conv)),
        'peak_source': pk['source'],
        'precision': 'fp32-accurate 3xbf16 split: every algorithmic MAC issues 3 bf16 MMAs, so issued tensor work is 3x '
                     'achieved (issued/peak = %.3f)' % (3 * achieved / pk['bf16_tflops']),
        'launches': len(conv), 'avg_launch_ms': conv_ms / max(1, len(conv)),
        'share_of_step': conv_ms / prof_ms_total if prof_ms_total else None,
        'measured_over': '%d instrumented eager steps (CUDA events around every launch; %.2f ms/step eager)' % (n_prof, prof_ms_total / n_prof),
        'wgrad_kernel': {'achieved': (sum(f for f, _ in wg) / (sum(t for _, t in wg) * 1e-3) / 1e12) if wg else None,
                         'launches': len(wg), 'share_of_step': sum(t for _, t in wg) / prof_ms_total if prof_ms_total else None},
        'whole_step_algorithmic_tflops': ALGO_FLOPS_PER_PAIR * B / (ms_step * 1e-3) / 1e12,
    }

    if rank != 0:
        return
    line = {
        'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16x3 (fp32 operands split hi+lo, fp32 accumulate in TMEM)', 'data': 'synthetic',
        'config': {'workload': 'StyleGAN2-1024 paired step, K=128 D=32 d=512, ResNet-18 R @1024^2, Z-space shift',
                   'batch_per_gpu': B, 'global_batch': world * B, 'parallelism': 'latents sharded dp%d, 1 grad all-reduce' % world,
                   'weights': 'random init (reference constructors), noise strength 0.1',
                   'l2': 'working set per step (>10 GB of activations) exceeds the 126 MB L2; no flush needed'},
        'clocks': sampler.summary(), 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
        'cuda_graph': bool(graphed),
    }
    if world == 1 and not args.no_cpu_baseline:
        pps, ms, ran, threads = oracle_step_timer(1, args.cpu_baseline_seconds, 0, 2)
        line['cpu_baseline'] = {'value': pps, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                                'sample': '%d step(s) of 1 pair (same workload, batch 1), %.1f s/step' % (ran, ms / 1e3)}
    print(json.dumps(line), flush=True)


def _shutdown():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    try:
        main()
    finally:
        _shutdown()
