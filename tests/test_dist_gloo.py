"""CPU, world_size 2 over gloo: the N>1 host logic — latent sharding and the single gradient all-reduce.
The oracle stands in for the CUDA compute (this is the only way to exercise the collective path without
GPUs); the contract checked is the one PairedTrainer relies on: mean-of-shard-gradients after
sum-all-reduce and 1/world scaling, identical on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from warpedganspace_b200 import dist as wdist
import oracle.support_sets as o_ss
import oracle.sngan as o_sn
import oracle.reconstructor as o_rec
import oracle.step as o_step


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [wdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_training_shards_must_be_equal_and_non_empty():
    assert wdist.shard_range(32, 3, 8, require_equal=True) == (12, 16)
    for n, world in ((6, 4), (3, 8), (0, 2)):
        with pytest.raises(ValueError):
            wdist.shard_range(n, 0, world, require_equal=True)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _problem():
    K, D, d, B = 8, 4, 128, 4
    g_sd = o_sn.init_state('sn_resnet32', 1, generator=_gen(1))
    s_sd = o_ss.init_state(K, D, d, generator=_gen(2))
    r_sd = o_rec.init_state('LeNet', K, 1, generator=_gen(3))
    g = _gen(4)
    z = torch.randn(B, d, generator=g)
    idx = torch.randint(0, K, (B,), generator=g)
    mag = o_step.sample_shift_magnitudes(B, 0.15, 0.25, generator=g)
    return g_sd, s_sd, r_sd, z, idx, mag


def _shard_grads(g_sd, s_sd, r_sd, z, idx, mag, lo, hi):
    gen_fn, _ = o_step.make_generator('SNGAN', g_sd, model='sn_resnet32')
    res = o_step.paired_step(gen_fn, s_sd, r_sd, z[lo:hi], idx[lo:hi], mag[lo:hi], reconstructor_type='LeNet')
    return res['grads']['S']['SUPPORT_SETS'], res['grads']['R']['path_indices.3.weight']


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    w, r, _ = wdist.init_from_env(backend='gloo')
    assert (w, r) == (world, rank)
    prob = _problem()
    lo, hi = wdist.shard_range(prob[3].shape[0], rank, world)
    gs, gr = _shard_grads(*prob, lo, hi)
    gs, gr = gs.clone(), gr.clone()
    wdist.all_reduce_sum_([gs, gr])
    gs /= world
    gr /= world
    slowest = wdist.max_over_ranks(float(rank + 1), 'cpu')
    wdist.barrier()
    if rank == 0:
        torch.save(dict(gs=gs, gr=gr, slowest=slowest), out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce(tmp_path):
    world = 2
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    prob = _problem()
    exp_s, exp_r = 0, 0
    for r in range(world):
        lo, hi = wdist.shard_range(prob[3].shape[0], r, world)
        a, b = _shard_grads(*prob, lo, hi)
        exp_s, exp_r = exp_s + a / world, exp_r + b / world
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    # thread counts differ between the workers and this process, so conv summation order does too
    assert rel(got['gs'], exp_s) < 1e-4, rel(got['gs'], exp_s)
    assert rel(got['gr'], exp_r) < 1e-4, rel(got['gr'], exp_r)
    assert got['slowest'] == 2.0


# ---- the training driver under two ranks (host logic only: the CUDA engine is replaced by a stub) ----------------------
class _StubEngine:
    """Reports statistics that depend on the rank's shard, so the cross-rank mean in stats.json can be checked."""
    _graph = None

    def __init__(self):
        self.seen = []

    def step(self, z, indices, magnitudes):
        self.seen.append((z.clone(), indices.clone(), magnitudes.clone()))
        m = z.mean()
        return dict(accuracy=torch.tensor(0.5), cls=m, reg=2 * m, loss=3 * m)


def _trainer_worker(rank, world, port, root, out_dir):
    import argparse
    from torch import nn
    from warpedganspace_b200.trainer import Trainer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    wdist.init_from_env(backend='gloo')
    params = argparse.Namespace(batch_size=6, num_support_sets=8, min_shift_magnitude=0.1, max_shift_magnitude=0.2,
                                z_truncation=None, support_set_lr=1e-4, reconstructor_lr=1e-4, lambda_cls=1.0, lambda_reg=0.25,
                                max_iter=4, log_freq=2, ckp_freq=4, tensorboard=False, quiet=True)

    class _M(nn.Module):
        dim_z = 16

        def __init__(self):
            super().__init__()
            self.p = nn.Parameter(torch.zeros(3))

    T = Trainer(params, 'exp', use_cuda=True, multi_gpu=True, root=root)
    eng = _StubEngine()
    T._device = lambda: torch.device('cpu')
    T._make_engine = lambda g, s, r: eng
    torch.manual_seed(11 + 100 * rank)                            # ranks need NOT share a seed: rank 0's draws are broadcast
    T.train(_M(), _M(), _M())
    torch.save(eng.seen, os.path.join(out_dir, 'seen_%d.pt' % rank))
    wdist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_trainer_shards_draws_and_averages_statistics(tmp_path):
    import json
    from warpedganspace_b200.trainer import _reference_magnitudes
    world = 2
    root, out_dir = str(tmp_path / 'experiments'), str(tmp_path)
    mp.spawn(_trainer_worker, args=(world, _free_port(), root, out_dir), nprocs=world, join=True)
    seen = [torch.load(os.path.join(out_dir, 'seen_%d.pt' % r)) for r in range(world)]
    # the single-process draw sequence (reference order) ...
    torch.manual_seed(11)
    full = []
    for _ in range(4):
        z = torch.randn(6, 16)
        idx = torch.randint(0, 8, [6])
        full.append((z, idx, _reference_magnitudes(6, 0.1, 0.2)))
    # ... is what the two ranks saw, split 3 + 3
    for it in range(4):
        for r in range(world):
            lo, hi = wdist.shard_range(6, r, world)
            for got, want in zip(seen[r][it], full[it]):
                assert torch.equal(got, want[lo:hi])
    # rank 0 alone wrote the files; every window holds the mean over ranks and iterations
    wip = os.path.join(root, 'wip', 'exp')
    stats = json.load(open(os.path.join(wip, 'stats.json')))
    assert sorted(stats) == ['2', '4']
    for key, its in (('2', (0, 1)), ('4', (2, 3))):
        want = sum(float(full[i][0][lo:hi].mean()) for i in its for lo, hi in (wdist.shard_range(6, r, world) for r in range(world))) / 4
        assert stats[key]['classification_loss'] == pytest.approx(want, rel=1e-5, abs=1e-7)
        assert stats[key]['total_loss'] == pytest.approx(3 * want, rel=1e-5, abs=1e-7)
        assert stats[key]['accuracy'] == pytest.approx(0.5)
    assert sorted(os.listdir(os.path.join(wip, 'models'))) == ['checkpoint.pt', 'reconstructor.pt', 'support_sets.pt',
                                                                'support_sets_init.pt']
    assert os.path.isdir(os.path.join(root, 'complete', 'exp'))
