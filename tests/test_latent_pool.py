"""CPU: the latent-pool format (sample_gan.py:156-179) against the 58 pools the reference ships
(tests/golden/latent_pool.pt, written by oracle/gen_golden.py::pin_latent_pool): the directory name of every pool entry
is the sha1 of its [1, dim_z] fp32 tensor, and save -> load round-trips bit-exactly."""
import os

import pytest
import torch

from warpedganspace_b200 import latent_pool as lp


def test_reference_pool_hashes(golden):
    fx = golden('latent_pool.pt')
    assert len(fx) == 58
    for key, z in fx.items():
        assert z.dim() == 2 and z.shape[0] == 1 and z.dtype == torch.float32
        assert lp.latent_code_hash(z) == key.split('/')[-1].split('_')[-1], key       # some pools are ordered: NNN_<hash>


def test_pool_round_trip(golden, tmp_path):
    fx = golden('latent_pool.pt')
    group = sorted(k for k in fx if k.startswith('SNGAN_AnimeFaces/SNGAN_AnimeFaces_6/'))
    zs = torch.cat([fx[k] for k in group])
    hashes = lp.save_latent_pool(zs, str(tmp_path / 'pool'))
    assert hashes == [k.split('/')[-1] for k in group]
    got_hashes, got = lp.load_latent_pool(str(tmp_path / 'pool'))
    assert got_hashes == sorted(hashes)
    by_hash = dict(zip(got_hashes, got))
    for h, z in zip(hashes, zs):
        assert torch.equal(by_hash[h], z)
        assert sorted(os.listdir(tmp_path / 'pool' / h)) == ['latent_code.pt']


def test_pool_rejects_tampering_and_bad_shapes(tmp_path):
    z = torch.randn(1, 16)
    (h,) = lp.save_latent_pool(z, str(tmp_path / 'pool'))
    torch.save(z + 1e-7, str(tmp_path / 'pool' / h / 'latent_code.pt'))
    with pytest.raises(ValueError, match='does not match'):
        lp.load_latent_pool(str(tmp_path / 'pool'))
    assert lp.load_latent_pool(str(tmp_path / 'pool'), verify=False)[0] == [h]
    with pytest.raises(ValueError):
        lp.latent_code_hash(torch.randn(2, 16))
    with pytest.raises(FileNotFoundError):
        os.makedirs(tmp_path / 'empty')
        lp.load_latent_pool(str(tmp_path / 'empty'))
