"""Experiment-directory helpers around the files the training driver writes (lib/trainer.py:74-89,288-308,
checkpoint2model.py:37-49, traverse_latent_space.py:252-297).

    <exp>/args.json                         argparse namespace of the run (lib/aux.py:93-95)
    <exp>/models/checkpoint.pt              {'iter', 'support_sets', 'reconstructor'}
    <exp>/models/support_sets(-<iter>).pt   SupportSets state dict (SUPPORT_SETS, ALPHAS, LOGGAMMA)
    <exp>/models/reconstructor(-<iter>).pt  Reconstructor state dict
"""
import argparse
import json
import os.path as osp

import torch


def checkpoint_to_models(exp_dir):
    """checkpoint2model.py: split ``models/checkpoint.pt`` into ``support_sets-<iter>.pt`` and
    ``reconstructor-<iter>.pt``; same errors for a malformed experiment directory.  Returns the iteration."""
    if not osp.isdir(exp_dir):
        raise NotADirectoryError('Invalid given directory: {}'.format(exp_dir))
    models_dir = osp.join(exp_dir, 'models')
    if not osp.isdir(models_dir):
        raise NotADirectoryError('Invalid models directory: {}'.format(models_dir))
    ckpt_file = osp.join(models_dir, 'checkpoint.pt')
    if not osp.isfile(ckpt_file):
        raise FileNotFoundError('Checkpoint file not found: {}'.format(ckpt_file))
    ckpt = torch.load(ckpt_file, map_location='cpu')
    it = ckpt['iter']
    torch.save(ckpt['support_sets'], osp.join(models_dir, 'support_sets-{}.pt'.format(it)))
    torch.save(ckpt['reconstructor'], osp.join(models_dir, 'reconstructor-{}.pt'.format(it)))
    return it


def load_experiment(exp_dir, iteration=None):
    """The set-up half of traverse_latent_space.py:252-297: read ``args.json``, build SupportSets with the recorded
    hyper-parameters and load ``models/support_sets.pt`` (or ``support_sets-<iteration>.pt``).  Returns (args, S);
    the generator is built separately with gan_load.build_* from args.gan_type."""
    from .support_sets import SupportSets
    with open(osp.join(exp_dir, 'args.json')) as f:
        args = argparse.Namespace(**json.load(f))
    name = 'support_sets.pt' if iteration is None else 'support_sets-{}.pt'.format(iteration)
    path = osp.join(exp_dir, 'models', name)
    if not osp.isfile(path):
        raise FileNotFoundError('Support sets weights not found: {}'.format(path))
    sd = torch.load(path, map_location='cpu')
    K, two_d_dim = sd['SUPPORT_SETS'].shape
    n_vec = sd['ALPHAS'].shape[1]
    dim = two_d_dim // n_vec
    if K != args.num_support_sets or n_vec != 2 * args.num_support_dipoles:
        raise ValueError('support_sets.pt (K=%d, 2D=%d) does not match args.json (K=%d, D=%d)'
                         % (K, n_vec, args.num_support_sets, args.num_support_dipoles))
    S = SupportSets(num_support_sets=K, num_support_dipoles=args.num_support_dipoles, support_vectors_dim=dim,
                    learn_alphas=args.learn_alphas, learn_gammas=args.learn_gammas,
                    gamma=1.0 / dim if getattr(args, 'gamma', None) is None else args.gamma)
    S.load_state_dict(sd)
    return args, S
