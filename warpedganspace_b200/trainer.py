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

    def forward_backward(self, z, indices, magnitudes):
        """One forward + backward; gradients land in the flat buffers.  Returns a dict of device tensors."""
        self.flat_s.zero_grad()
        self.flat_r.zero_grad()
        where = self.G.get_w(z).detach() if self.shift_in_w_space else z            # lib/trainer.py:236
        shift = self.S.warp(indices, where, magnitudes)                              # :235
        if hasattr(self.G, 'forward_pair'):
            img, img_shifted = self.G.forward_pair(z, shift)                         # :200, :239
        else:
            with torch.no_grad():
                img = self.G(z)
            img_shifted = self.G(z, shift)
        logits, pred = self.R(img.detach(), img_shifted)                             # :242
        cls = F.cross_entropy(logits, indices)                                       # :245
        reg = torch.mean(torch.abs(pred - magnitudes))                               # :246
        loss = self.lambda_cls * cls + self.lambda_reg * reg                         # :249
        loss.backward()                                                              # :250
        acc = (logits.argmax(dim=1) == indices).float().mean()
        return dict(loss=loss.detach(), cls=cls.detach(), reg=reg.detach(), accuracy=acc, logits=logits.detach(),
                    pred=pred.detach(), shift=shift.detach(), img=img.detach(), img_shifted=img_shifted.detach())

    def all_reduce_gradients(self):
        if self.world > 1:
            wdist.all_reduce_sum_([self.flat_s.grad, self.flat_r.grad], group=self.pg)

    def optimizer_step(self):
        scale = 1.0 / self.world
        self.flat_s.adam_step(self.lr_s, grad_scale=scale)                            # :253
        self.flat_r.adam_step(self.lr_r, grad_scale=scale)                            # :254

    def step(self, z, indices, magnitudes, eager=False):
        if self._graph is not None and not eager:
            return self._replay(z, indices, magnitudes)
        out = self.forward_backward(z, indices, magnitudes)
        self.all_reduce_gradients()
        self.optimizer_step()
        return out

    # ---- CUDA-graph mode: the whole step (≈330 of our launches + ≈500 small ATen ones) becomes one graph launch ----
    _graph = None

    def capture(self, z, indices, magnitudes, warmup=3):
        """Capture forward + backward + all-reduce + both Adam updates for this batch shape.  Afterwards step()
        copies its arguments into the static inputs and replays.  Returns True on success; on failure the trainer
        stays in eager mode (still entirely on the CUDA kernels).  Call it BEFORE any eager backward pass: autograd
        caches each parameter's AccumulateGrad node together with the stream it first ran on, and a node bound to
        the default stream invalidates the capture."""
        self._static_in = (z.clone(), indices.clone(), magnitudes.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self.forward_backward(*self._static_in)
                    self.all_reduce_gradients()
                    self.optimizer_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(graph):
                out = self.forward_backward(*self._static_in)
                self.all_reduce_gradients()
                self.optimizer_step()
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
