"""Host-side helpers of the training driver, mirroring the reference ``lib/aux.py`` names:
``sample_z`` (lib/aux.py:39-53), ``TrainingStatTracker`` (:13-36), ``create_exp_dir`` (:56-104) and
``sec2dhms``.  No compute happens here; the draws are made on the host in the reference's order so that a
seeded run sees the same latents as the reference loop.
"""
import json
import os
import os.path as osp
import sys

import numpy as np
import torch


def sample_z(batch_size, dim_z, truncation=None):
    """N(0, I) latents [batch_size, dim_z]; with ``truncation`` t != 1 every coordinate is drawn from the
    standard normal truncated to [-t, t] (scipy ``truncnorm.rvs`` on the host, numpy global RNG, fp64 -> fp32),
    exactly the reference's two branches (lib/aux.py:50-53)."""
    if truncation is None or truncation == 1.0:
        return torch.randn(batch_size, dim_z)
    from scipy.stats import truncnorm
    draws = truncnorm.rvs(-truncation, truncation, size=(batch_size, dim_z))
    return torch.from_numpy(draws).to(torch.float)


class TrainingStatTracker(object):
    """Running lists of the four per-iteration statistics between two log lines (lib/aux.py:13-36)."""

    KEYS = ('accuracy', 'classification_loss', 'regression_loss', 'total_loss')

    def __init__(self):
        self.stat_tracker = {k: [] for k in self.KEYS}

    def update(self, accuracy, classification_loss, regression_loss, total_loss):
        for k, v in zip(self.KEYS, (accuracy, classification_loss, regression_loss, total_loss)):
            self.stat_tracker[k].append(float(v))

    def get_means(self):
        return {k: np.mean(v) for k, v in self.stat_tracker.items()}

    def flush(self):
        for k in self.stat_tracker:
            self.stat_tracker[k] = []


def exp_dir_name(args):
    """``<gan>(-<res>-{Z,W})(-<biggan classes>)-<R>-K<K>-D<D>(-LearnAlphas)(-LearnGammas)-eps<min>_<max>``
    (lib/aux.py:71-88)."""
    parts = [str(args.gan_type)]
    if args.gan_type == 'StyleGAN2':
        parts += [str(args.stylegan2_resolution), 'W' if args.shift_in_w_space else 'Z']
    if args.gan_type == 'BigGAN':
        parts.append(''.join(str(c) for c in args.biggan_target_classes))
    parts += [str(args.reconstructor_type), 'K%s' % args.num_support_sets, 'D%s' % args.num_support_dipoles]
    if args.learn_alphas:
        parts.append('LearnAlphas')
    if args.learn_gammas:
        parts.append('LearnGammas')
    parts.append('eps%s_%s' % (args.min_shift_magnitude, args.max_shift_magnitude))
    return '-'.join(parts)


def create_exp_dir(args, root='experiments'):
    """Creates ``<root>/wip/<name>/`` with ``args.json`` and ``command.sh`` (lib/aux.py:90-104); returns the name."""
    name = exp_dir_name(args)
    wip = osp.join(root, 'wip', name)
    os.makedirs(wip, exist_ok=True)
    with open(osp.join(wip, 'args.json'), 'w') as f:
        json.dump(args.__dict__, f)
    with open(osp.join(wip, 'command.sh'), 'w') as f:
        f.write('#!/usr/bin/bash\n' + ' '.join(sys.argv) + '\n')
    return name


def sec2dhms(t):
    d, rem = divmod(int(t), 86400)
    h, rem = divmod(rem, 3600)
    m, s = divmod(rem, 60)
    return '%02d days, %02d hours, %02d minutes, %02d seconds' % (d, h, m, s)
