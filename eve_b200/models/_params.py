"""Parameter trees with the reference's state_dict key names.

The kernels take flat pointer tables, so the modules here own nothing but ``nn.Parameter``s
arranged under the same dotted names the reference's nn.Module hierarchy produces
(e.g. ``cnn_layers.layer1.0.conv1.weight``, ``network.between_module.encoder_blocks.0.
layers.2.weight``) -- which is what CheckpointManager, Adam and the released weights see
(reference: src/core/checkpoint_manager.py:50-92, src/utils/load_model.py:35-57).
"""
import math

import torch
from torch import nn


class ParamNode(nn.Module):
    """A nameless container; children / parameters are attached by dotted path."""

    def forward(self, *args, **kwargs):
        raise RuntimeError('ParamNode only holds parameters; call the owning network instead.')


def attach(root, shapes, init_fn):
    """Create ``nn.Parameter``s under ``root`` for every (dotted name -> shape)."""
    for name, shape in shapes.items():
        parts = name.split('.')
        node = root
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, ParamNode())
            node = node._modules[part]
        node.register_parameter(parts[-1], nn.Parameter(init_fn(name, tuple(shape))))


def lookup(root, name):
    node = root
    parts = name.split('.')
    for part in parts[:-1]:
        node = node._modules[part]
    return node._parameters[parts[-1]]


def kaiming_normal_fan_out(shape):
    """nn.init.kaiming_normal_(mode='fan_out', nonlinearity='relu') for OIHW conv weights."""
    fan_out = shape[0] * int(math.prod(shape[2:]))
    return torch.randn(shape) * math.sqrt(2.0 / fan_out)


def linear_default_weight(shape):
    """nn.Linear / nn.*Cell default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (kaiming_uniform a=sqrt5)."""
    bound = 1.0 / math.sqrt(shape[1])
    return (torch.rand(shape) * 2.0 - 1.0) * bound


def uniform(shape, bound):
    return (torch.rand(shape) * 2.0 - 1.0) * bound
