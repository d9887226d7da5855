"""Builds nn.Module trees from flat (dotted-name -> tensor) specs so that state-dict keys equal the
reference's, without re-declaring the reference's module classes."""
import torch
from torch import nn


class Node(nn.Module):
    """Bare container."""


def add(root, name, tensor, buffer=False, requires_grad=True):
    parts = name.split('.')
    m = root
    for p in parts[:-1]:
        child = m._modules.get(p)
        if child is None:
            child = Node()
            m.add_module(p, child)
        m = child
    if buffer:
        m.register_buffer(parts[-1], tensor)
    else:
        m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=requires_grad))


def tensors(module):
    """Live name -> tensor map (parameters and buffers)."""
    d = dict(module.named_parameters())
    d.update(dict(module.named_buffers()))
    return d
