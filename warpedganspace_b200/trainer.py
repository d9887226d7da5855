"""Paired-image training step (the loop body of the reference ``Trainer.train``, lib/trainer.py:184-261)
restated for one process per GPU.

Differences from the reference loop, none of which change the numbers it produces for given draws:
  * G(z) and G(z + shift) run as one batched pass; the frozen generator gets no weight gradients
    (the reference back-propagates into it and throws the result away, lib/trainer.py:190,250);
  * path indices go to the RBF kernel directly (no Python loop building the one-hot mask, :227-231, which
    costs B host<->device syncs per step) and the shift magnitude is fused into that kernel (:235);
  * parameters, gradients and Adam moments of S and R live in flat buffers: zero_grad is one memset,
    each optimiser is one kernel, and under torch.distributed the step does ONE all-reduce of the
    flat gradients (sum, then 1/world inside the Adam kernel) — nothing else crosses GPUs;
  * statistics stay on the device until asked for (the reference does 3 .item() syncs per step, :257-261).
Train-mode BatchNorm in R uses per-rank batch statistics, as the reference's DataParallel replicas do.
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from . import _lib
from . import dist as wdist


class FlatParams:
    """Re-homes a list of parameters into one flat fp32 buffer (params become views), with matching flat
    gradient and Adam-moment buffers."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FlatParams needs CUDA parameters (no CPU fallback)')
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]            # keep every view 16-byte aligned
        self.n = sum(sizes)
        self.flat = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        off = 0
        with torch.no_grad():
            for p, sz in zip(self.params, sizes):
                view = self.flat[off: off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[off: off + p.numel()].view_as(p)
                p._wgs_flat_grad = True          # resnet_fused may accumulate weight gradients straight into this view
                off += sz
        self.step_count = 0
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)      # device-side step counter (graph replay)

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:                      # autograd may have replaced .grad; re-point it at the flat view
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or \
                    p.grad.data_ptr() >= self.grad.data_ptr() + self.n * 4:
                raise RuntimeError('parameter gradient left the flat buffer')

    def adam_step(self, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        # the step counter lives on the device so that the same launches can be replayed from a CUDA graph
        self.step_count += 1
        _lib.call('wgs_step_increment', _lib.ptr(self.step_dev), _lib.stream())
        _lib.call('wgs_adam_step', _lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.exp_avg),
                  _lib.ptr(self.exp_avg_sq), self.n, float(lr), float(betas[0]), float(betas[1]), float(eps),
                  0, float(grad_scale), _lib.ptr(self.step_dev), _lib.stream())


def sample_shift_magnitudes(batch, min_mag, max_mag, device, generator=None):
    """lib/trainer.py:212-221, including the index-weighted multinomial draw."""
    pos = (min_mag - max_mag) * torch.rand(batch, device=device, generator=generator) + max_mag
    neg = (min_mag - max_mag) * torch.rand(batch, device=device, generator=generator) - min_mag
    pool = torch.cat((neg, pos))
    ids = torch.arange(len(pool), dtype=torch.float, device=device)
    return pool[torch.multinomial(ids, batch, replacement=False, generator=generator)]


class PairedTrainer:
    """Owns the S / R optimiser state and runs training steps for a frozen generator wrapper."""

    def __init__(self, generator, support_sets, reconstructor, *, support_set_lr=1e-4, reconstructor_lr=1e-4,
                 lambda_cls=1.0, lambda_reg=0.25, shift_in_w_space=False, process_group=None):
        self.G, self.S, self.R = generator, support_sets, reconstructor
        self.G.eval()
        for p in self.G.parameters():
            p.requires_grad_(False)
        self.S.train()
        self.R.train()
        self.lr_s, self.lr_r = support_set_lr, reconstructor_lr
        self.lambda_cls, self.lambda_reg = lambda_cls, lambda_reg
        self.shift_in_w_space = shift_in_w_space
        self.flat_s = FlatParams(self.S.parameters())
        self.flat_r = FlatParams(self.R.parameters())
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.sync_replicas()

    def sync_replicas(self):
        """Make every rank a replica of rank 0: S / R parameters and R's BatchNorm buffers are broadcast (the reference's
        DataParallel re-broadcasts rank 0's modules on every forward, lib/trainer.py:164-165; only gradients are reduced
        afterwards, so replicas that start apart never converge).  No-op on one rank."""
        if self.world > 1:
            with torch.no_grad():
                wdist.broadcast_([self.flat_s.flat, self.flat_r.flat] + [b for b in self.R.buffers()], src=0, group=self.pg)

    # Side streams (forked from / joined to the current stream with events, so the same code is captured into the CUDA graph):
    #   _side : R's weight-gradient kernels, then R's gradient all-reduce, then (inside step()) R's Adam update - all of it
    #           underneath the generator's data-gradient pass; joined before forward_backward returns.  Default (= lowest)
    #           priority; the graph is captured on a high-priority stream so that this work only fills what the main chain
    #           leaves idle;
    #   _pack : the grouped weight pack of R at the start of the step, underneath the RBF / mapping / low-resolution layers.
    # WGS_SIDE_STREAMS=0 keeps everything on one stream (A/B switch).
    _side = None
    _pack = None
    _side_busy = False          # work is in flight on _side
    _r_reduced = False          # R's gradient all-reduce of this step has been issued
    _r_stepped = False          # R's Adam update of this step has been issued
    _early_adam = False         # step(): R's Adam update may run as soon as R's gradients are final

    def _fused_resnet(self):
        return (os.environ.get('WGS_SIDE_STREAMS', '1') != '0' and getattr(self.R, 'reconstructor_type', None) == 'ResNet'
                and getattr(self.R, 'fused', False))

    def _streams(self):
        if self._side is None:
            self._side = torch.cuda.Stream()
            self._pack = torch.cuda.Stream()
        return self._side, self._pack

    def forward_backward(self, z, indices, magnitudes):
        """One forward + backward; gradients land in the flat buffers.  Returns a dict of device tensors."""
        self.flat_s.zero_grad()
        self.flat_r.zero_grad()
        if self._fused_resnet():
            from . import resnet_fused
            side, pack = self._streams()
            net = self.R.features_extractor
            net._wgs_wgrad_stream = side
            resnet_fused.prepack(net, (self.R.channels, self.R.channels), pack)
        where = self.G.get_w(z).detach() if self.shift_in_w_space else z            # lib/trainer.py:236
        shift = self.S.warp(indices, where, magnitudes)                              # :235
        if hasattr(self.G, 'forward_pair'):
            img, img_shifted = self.G.forward_pair(z, shift)                         # :200, :239
        else:
            with torch.no_grad():
                img = self.G(z)
            img_shifted = self.G(z, shift)
        # loss.backward() (:250) in two legs cut at the Reconstructor's input: after the first leg every gradient of R is
        # final, so its all-reduce (45 MB) runs on a side stream underneath the generator's data-gradient pass (~3 ms);
        # only the S gradients, final after the very last kernel, are reduced in the open (all_reduce_gradients).
        x2 = img_shifted.detach().requires_grad_(True)
        logits, pred = self.R(img.detach(), x2)                                      # :242
        cls = F.cross_entropy(logits, indices)                                       # :245
        reg = torch.mean(torch.abs(pred - magnitudes))                               # :246
        loss = self.lambda_cls * cls + self.lambda_reg * reg                         # :249
        loss.backward()                                                              # R's leg of :250
        self._start_r_reduce()
        img_shifted.backward(x2.grad)                                                # G data-gradient + RBF leg
        self._join_side()
        acc = (logits.argmax(dim=1) == indices).float().mean()
        return dict(loss=loss.detach(), cls=cls.detach(), reg=reg.detach(), accuracy=acc, logits=logits.detach(),
                    pred=pred.detach(), shift=shift.detach(), img=img.detach(), img_shifted=img_shifted.detach())

    def _start_r_reduce(self):
        """After R's leg of the backward pass: on the side stream (behind R's weight-gradient kernels, which already run
        there) sum R's flat gradient over ranks and, inside step(), apply R's Adam update, while the main stream goes on
        with the generator's data-gradient pass."""
        fused = self._fused_resnet()
        if self.world <= 1 and not fused:
            return
        side, _ = self._streams()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            if self.world > 1:
                wdist.all_reduce_sum_([self.flat_r.grad], group=self.pg)
                self._r_reduced = True
            if self._early_adam:
                self.flat_r.adam_step(self.lr_r, grad_scale=1.0 / self.world)          # lib/trainer.py:254
                self._r_stepped = True
        self._side_busy = True

    def _join_side(self):
        if self._side_busy:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_busy = False
        net = getattr(self.R, 'features_extractor', None)
        if net is not None:
            net._wgs_keep = None                               # the side stream's readers are ordered before us now
            net._wgs_wgrad_stream = None                       # a backward pass outside this engine stays on one stream

    def all_reduce_gradients(self):
        """Two NCCL all-reduces per step: R's (started inside forward_backward, overlapped) and S's (here)."""
        if self.world > 1:
            if not self._r_reduced:                            # gradients produced outside forward_backward
                wdist.all_reduce_sum_([self.flat_r.grad], group=self.pg)
            wdist.all_reduce_sum_([self.flat_s.grad], group=self.pg)
        self._r_reduced = False

    def optimizer_step(self):
        scale = 1.0 / self.world
        self.flat_s.adam_step(self.lr_s, grad_scale=scale)                            # :253
        if not self._r_stepped:
            self.flat_r.adam_step(self.lr_r, grad_scale=scale)                        # :254
        self._r_stepped = False

    def step(self, z, indices, magnitudes, eager=False):
        if self._graph is not None and not eager:
            return self._replay(z, indices, magnitudes)
        return self._full_step(z, indices, magnitudes)

    def _full_step(self, z, indices, magnitudes):
        self._early_adam = True
        try:
            out = self.forward_backward(z, indices, magnitudes)
        finally:
            self._early_adam = False
        self.all_reduce_gradients()
        self.optimizer_step()
        return out

    # ---- CUDA-graph mode: the whole step (≈330 of our launches + ≈500 small ATen ones) becomes one graph launch ----
    _graph = None

    def _training_state(self):
        """Everything a step mutates: flat parameters, Adam moments and step counters, R's BatchNorm buffers."""
        ts = []
        for f in (self.flat_s, self.flat_r):
            ts += [f.flat, f.exp_avg, f.exp_avg_sq, f.step_dev]
        ts += [b for b in self.R.buffers()]
        return ts

    def capture(self, z, indices, magnitudes, warmup=3, preserve_state=False):
        """Capture forward + backward + all-reduce + both Adam updates for this batch shape.  Afterwards step()
        copies its arguments into the static inputs and replays.  Returns True on success; on failure the trainer
        stays in eager mode (still entirely on the CUDA kernels).  Call it BEFORE any eager backward pass: autograd
        caches each parameter's AccumulateGrad node together with the stream it first ran on, and a node bound to
        the default stream invalidates the capture.  The warm-up passes are real training steps on the given batch;
        with ``preserve_state`` the parameters, optimiser state and BatchNorm buffers are put back afterwards so that
        capturing does not advance training (the Trainer driver relies on this to follow the reference step for step)."""
        saved = [(t, t.clone()) for t in self._training_state()] if preserve_state else []
        counts = (self.flat_s.step_count, self.flat_r.step_count)
        ok = self._capture(z, indices, magnitudes, warmup)
        if preserve_state:
            with torch.no_grad():
                for t, c in saved:
                    t.copy_(c)
            self.flat_s.step_count, self.flat_r.step_count = counts
            torch.cuda.synchronize()
        return ok

    def _capture(self, z, indices, magnitudes, warmup):
        self._static_in = (z.clone(), indices.clone(), magnitudes.clone())
        main = torch.cuda.Stream(priority=-1)         # above the side streams' (default = lowest) priority
        main.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(main):
                for _ in range(warmup):
                    self._full_step(*self._static_in)
            torch.cuda.current_stream().wait_stream(main)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(graph, stream=main):
                out = self._full_step(*self._static_in)
            self.launches_per_step = _lib.launch_count() - before
            self._static_out = out
            self._graph = graph
            return True
        except Exception as e:                                   # pragma: no cover - depends on driver / NCCL support
            self._graph = None
            self.capture_error = repr(e)
            torch.cuda.synchronize()
            return False

    def _replay(self, z, indices, magnitudes):
        zi, ii, mi = self._static_in
        zi.copy_(z, non_blocking=True)
        ii.copy_(indices, non_blocking=True)
        mi.copy_(magnitudes, non_blocking=True)
        self._graph.replay()
        self.flat_s.step_count += 1
        self.flat_r.step_count += 1
        return self._static_out


class Trainer(object):
    """Drop-in for the reference training driver ``lib.trainer.Trainer`` (lib/trainer.py:24-319): same constructor
    (``params`` = the argparse namespace of train.py:54-90, ``exp_dir``, ``use_cuda``, ``multi_gpu``), the same
    ``train(generator, support_sets, reconstructor)`` entry point and the same files under
    ``experiments/{wip,complete}/<exp_dir>/``: ``stats.json`` ({iteration: means of the last log window}),
    ``models/support_sets_init.pt``, ``models/checkpoint.pt`` ({'iter', 'support_sets', 'reconstructor'} every
    ``ckp_freq``), ``models/support_sets.pt`` and ``models/reconstructor.pt`` at the end, then wip -> complete
    without the checkpoint.  Resume restarts at the checkpointed iteration with fresh optimiser state, as the
    reference does (:83-89).

    What differs, by design:
      * every iteration is one ``PairedTrainer.step`` (optionally a CUDA-graph replay) instead of ~40 framework
        calls, and the four statistics stay on the device until a log line needs them (the reference reads three
        scalars back per iteration, :257-261);
      * latents, path indices and magnitudes are still drawn on the HOST with the reference's calls in the
        reference's order (:187, :193, :204-214), so a seeded run consumes the same random stream;
      * ``multi_gpu`` means one process per GPU under torchrun (torch.distributed) rather than nn.DataParallel:
        every rank draws the full batch from the same seed and takes its shard, gradients are summed in one
        all-reduce, rank 0 writes the files;
      * there is no CPU path: ``use_cuda=False`` raises.
    """

    def __init__(self, params=None, exp_dir=None, use_cuda=False, multi_gpu=False, root='experiments'):
        import json
        import os
        import os.path as osp
        if params is None:
            raise ValueError('Cannot build a Trainer instance with empty params: params={}'.format(params))
        self.params = params
        self.use_cuda = use_cuda
        self.multi_gpu = multi_gpu
        self.world, self.rank, self.local_rank = wdist.env_world()
        self.wip_dir = osp.join(root, 'wip', exp_dir)
        self.complete_dir = osp.join(root, 'complete', exp_dir)
        self.stats_json = osp.join(self.wip_dir, 'stats.json')
        self.models_dir = osp.join(self.wip_dir, 'models')
        self.checkpoint = osp.join(self.models_dir, 'checkpoint.pt')
        if self.rank == 0:
            os.makedirs(self.models_dir, exist_ok=True)
            if not osp.isfile(self.stats_json):
                with open(self.stats_json, 'w') as out:
                    json.dump({}, out)
        from .aux import TrainingStatTracker
        self.stat_tracker = TrainingStatTracker()
        self.iter_times = []
        self.tb_writer = None
        if getattr(params, 'tensorboard', False) and self.rank == 0:
            try:                                                          # optional, as in the reference (:56-64)
                from torch.utils.tensorboard import SummaryWriter
                tb_dir = osp.join(self.wip_dir, 'tensorboard')
                os.makedirs(tb_dir, exist_ok=True)
                self.tb_writer = SummaryWriter(log_dir=tb_dir)
            except Exception as e:                                        # pragma: no cover
                print('#. TensorBoard unavailable: %r' % (e,))

    # ---- checkpoint I/O (state dicts are saved as CPU clones: parameters are views into flat buffers) ----
    @staticmethod
    def _cpu_state(module):
        return {k: v.detach().to('cpu', copy=True) for k, v in module.state_dict().items()}

    def get_starting_iteration(self, support_sets, reconstructor):
        """lib/trainer.py:74-89."""
        import os.path as osp
        starting_iter = 1
        if osp.isfile(self.checkpoint):
            ckpt = torch.load(self.checkpoint, map_location='cpu')
            starting_iter = ckpt['iter']
            support_sets.load_state_dict(ckpt['support_sets'])
            reconstructor.load_state_dict(ckpt['reconstructor'])
        return starting_iter

    def _finish(self, quiet=False):
        import shutil
        if self.rank != 0:
            return
        try:
            shutil.copytree(src=self.wip_dir, dst=self.complete_dir, ignore=shutil.ignore_patterns('checkpoint.pt'))
        except IOError as e:
            if not quiet:
                print('  \\__Already exists -- {}'.format(e))

    def log_progress(self, iteration, mean_iter_time, elapsed_time, eta):
        """lib/trainer.py:91-127: fold the window means into stats.json (keyed by iteration) and print them."""
        import json
        from .aux import sec2dhms
        stats = self.stat_tracker.get_means()
        self.stat_tracker.flush()
        if self.rank != 0:
            return stats
        with open(self.stats_json) as f:
            stats_dict = json.load(f)
        stats_dict.update({iteration: {k: float(v) for k, v in stats.items()}})
        with open(self.stats_json, 'w') as out:
            json.dump(stats_dict, out)
        if not getattr(self.params, 'quiet', False):
            print('  \\__.Training [bs: {}] [iter: {:06d}/{:06d}]'.format(self.params.batch_size, iteration, self.params.max_iter))
            print('      \\__Batch accuracy      : {:.03f}'.format(stats['accuracy']))
            print('      \\__Classification loss : {:.08f}'.format(stats['classification_loss']))
            print('      \\__Regression loss     : {:.08f}'.format(stats['regression_loss']))
            print('      \\__Total loss          : {:.08f}'.format(stats['total_loss']))
            print('      \\__Mean iter time      : {:.3f} sec'.format(mean_iter_time))
            print('      \\__Elapsed time        : {}'.format(sec2dhms(elapsed_time)))
            print('      \\__ETA                 : {}'.format(sec2dhms(eta)))
        return stats

    def draw_batch(self, dim_z):
        """The reference's host draws in the reference's order (lib/trainer.py:187-221): z, path indices, then the
        magnitude pool and its index-weighted multinomial pick."""
        from .aux import sample_z
        p = self.params
        z = sample_z(batch_size=p.batch_size, dim_z=dim_z, truncation=getattr(p, 'z_truncation', None))
        indices = torch.randint(0, p.num_support_sets, [p.batch_size])
        magnitudes = _reference_magnitudes(p.batch_size, p.min_shift_magnitude, p.max_shift_magnitude)
        return z, indices, magnitudes

    def _device(self):
        return torch.device('cuda', self.local_rank if self.multi_gpu else torch.cuda.current_device())

    def _make_engine(self, generator, support_sets, reconstructor):
        p = self.params
        return PairedTrainer(generator, support_sets, reconstructor, support_set_lr=p.support_set_lr,
                             reconstructor_lr=p.reconstructor_lr, lambda_cls=p.lambda_cls, lambda_reg=p.lambda_reg,
                             shift_in_w_space=bool(getattr(p, 'shift_in_w_space', False)))

    def train(self, generator, support_sets, reconstructor):
        import os.path as osp
        import sys
        import time
        p = self.params
        if not self.use_cuda:
            raise RuntimeError('warpedganspace_b200.Trainer needs use_cuda=True: libwgs_b200 has no CPU fallback')
        if self.rank == 0:
            torch.save(self._cpu_state(support_sets), osp.join(self.models_dir, 'support_sets_init.pt'))
        device = self._device()
        generator.to(device).eval()
        support_sets.to(device).train()
        reconstructor.to(device).train()
        starting_iter = self.get_starting_iteration(support_sets, reconstructor)
        if starting_iter == p.max_iter:                                            # :169-177
            print('#. This experiment has already been completed and can be found @ {}'.format(self.wip_dir))
            self._finish()
            sys.exit()
        parallel = self.multi_gpu and self.world > 1
        if parallel:
            # one process per GPU under torchrun: the process group must exist BEFORE the engine is built (the engine reads
            # its world size from it); a WORLD_SIZE > 1 environment without one would silently train N independent models
            wdist.init_from_env()
            if not wdist.is_parallel():
                raise RuntimeError('multi_gpu=True with WORLD_SIZE=%d but torch.distributed could not be initialised' % self.world)
        engine = self._make_engine(generator, support_sets, reconstructor)
        if parallel and getattr(engine, 'world', self.world) != self.world:
            raise RuntimeError('engine spans %d rank(s) but the launcher started %d' % (engine.world, self.world))
        lo, hi = wdist.shard_range(p.batch_size, self.rank, self.world, require_equal=True) if parallel else (0, p.batch_size)
        use_graph = bool(getattr(p, 'cuda_graph', False))
        window = []                                                                # device-side [acc, cls, reg, loss] rows
        t0 = time.time()
        for iteration in range(starting_iter, p.max_iter + 1):
            iter_t0 = time.time()
            z, indices, magnitudes = self.draw_batch(generator.dim_z)
            full = [t.to(device, non_blocking=True) for t in (z, indices, magnitudes)]
            if parallel:                         # rank 0's host draws are THE draws (ranks need not share a seed)
                wdist.broadcast_(full, src=0)
            batch = tuple(t[lo:hi] for t in full)
            if use_graph and engine._graph is None and iteration == starting_iter:
                engine.capture(*batch, preserve_state=True)
            out = engine.step(*batch)
            window.append(torch.stack([out['accuracy'], out['cls'], out['reg'], out['loss']]).clone())
            if iteration % p.log_freq == 0 or iteration % p.ckp_freq == 0 or iteration == p.max_iter:
                rows = torch.stack(window)
                if parallel:                                                       # equal shards: mean of rank means
                    wdist.all_reduce_sum_([rows])
                    rows /= self.world
                for acc, cls, reg, tot in rows.cpu().tolist():                     # the only device->host read
                    self.stat_tracker.update(acc, cls, reg, tot)
                window = []
                if self.tb_writer is not None:
                    for key, value in self.stat_tracker.get_means().items():
                        self.tb_writer.add_scalar(key, value, iteration)
            self.iter_times.append(time.time() - iter_t0)
            elapsed = time.time() - t0
            if iteration % p.log_freq == 0:
                eta = elapsed * ((p.max_iter - iteration) / (iteration - starting_iter + 1))
                self.log_progress(iteration, sum(self.iter_times) / len(self.iter_times), elapsed, eta)
            if iteration % p.ckp_freq == 0 and self.rank == 0:                     # :288-296
                torch.save({'iter': iteration, 'support_sets': self._cpu_state(support_sets),
                            'reconstructor': self._cpu_state(reconstructor)}, self.checkpoint)
        if self.rank == 0:                                                         # :301-319
            torch.save(self._cpu_state(support_sets), osp.join(self.models_dir, 'support_sets.pt'))
            torch.save(self._cpu_state(reconstructor), osp.join(self.models_dir, 'reconstructor.pt'))
        self._finish(quiet=True)
        return engine


def _reference_magnitudes(batch, min_mag, max_mag):
    """Host draw of lib/trainer.py:204-214 from the global torch RNG (positive pool first, as there)."""
    pos = (min_mag - max_mag) * torch.rand(batch) + max_mag
    neg = (min_mag - max_mag) * torch.rand(batch) - min_mag
    pool = torch.cat((neg, pos))
    ids = torch.arange(len(pool), dtype=torch.float)
    return pool[torch.multinomial(input=ids, num_samples=batch, replacement=False)]
