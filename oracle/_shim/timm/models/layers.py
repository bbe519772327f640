"""Stand-ins for timm.models.layers.{DropPath,to_2tuple,trunc_normal_}; the reference
uses them at networks/swin_transformer_sr.py:10,107,196,199 and rdst_variations.py:1311.
None of them contributes forward arithmetic for the RDST-E1 configuration."""
import torch
from torch import nn


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep
