"""CPU: the C-ABI library builds, loads, and exports every symbol include/wgs_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os

import pytest
import torch

from warpedganspace_b200 import _lib
from warpedganspace_b200.csrc import build as wgs_build


@pytest.fixture(scope='module')
def lib():
    wgs_build.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_entry_points():
    protos = _lib.header_prototypes()
    assert len(protos) >= 8
    for name in ('wgs_last_error', 'wgs_rbf_warp_forward', 'wgs_rbf_warp_backward', 'wgs_rbf_traverse'):
        assert name in protos


def test_library_exports_every_declared_symbol(lib):
    for name in _lib.header_prototypes():
        assert hasattr(lib, name), 'libwgs_b200.so lacks %s declared in include/wgs_b200.h' % name


def test_loader_binds_signatures():
    l = _lib.load()
    assert l.wgs_version() >= 100
    assert l.wgs_last_error() is not None


def test_no_cpu_fallback():
    from warpedganspace_b200 import SupportSets
    S = SupportSets(4, 2, 8, learn_gammas=True, gamma=1.0 / 8)
    mask = torch.zeros(2, 4)
    mask[:, 1] = 1
    with pytest.raises(RuntimeError):
        S(mask, torch.randn(2, 8))


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_struct_mirrors_match_the_header_layout():
    """The ctypes mirrors of wgs_conv_desc / wgs_linear_problem have the size the compiled library reports (no GPU needed:
    the size functions are plain host code)."""
    from warpedganspace_b200 import conv, stylegan2
    l = _lib.load()
    assert l.wgs_conv_desc_size() == ctypes.sizeof(conv.ConvDesc)
    assert l.wgs_linear_problem_size() == ctypes.sizeof(stylegan2.LinearProblem)


def test_every_product_module_refuses_cpu_tensors():
    """No CPU fallback anywhere on the product path: each public entry raises on CPU input instead of computing."""
    from warpedganspace_b200 import conv, image_out
    from warpedganspace_b200.reconstructor import Reconstructor
    with pytest.raises(RuntimeError):
        conv.pack_split32(torch.randn(4, 32))
    with pytest.raises(RuntimeError):
        image_out.images_to_uint8(torch.randn(1, 3, 8, 8))
    R = Reconstructor('LeNet', 4, 1)
    with pytest.raises(RuntimeError):
        R(torch.randn(2, 1, 32, 32), torch.randn(2, 1, 32, 32))
