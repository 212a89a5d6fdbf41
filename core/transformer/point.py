"""Parameter containers of the point-cloud encoder (``PointEncoderEmbed``), reference checkpoint schema.

Mirrors ``/root/reference/core/transformer/point.py`` (``DummyLatent`` :23, ``PointEmbed`` :37, ``FeedForward``
:74, ``ResCrossAttBlock`` :108, ``PointEncoderEmbed`` :172).  Weights only; the arithmetic (Fourier embedding,
2048 x 8192 cross-attention, GEGLU FFN) runs in ``edgerunner_b200/csrc``.  The FPS ``PointEncoder`` variant needs the
un-vendored torch_cluster and is used by no preset: out of scope.
"""

import numpy as np
import torch
from torch import nn


class DummyLatent:
    """Deterministic "posterior": sample() == mode() == mean; kl() is the L2 penalty 0.5 * sum(mean^2)."""

    def __init__(self, mean):
        self.mean = mean

    def sample(self):
        return self.mean

    def mode(self):
        return self.mean

    def kl(self):
        return 0.5 * torch.sum(torch.pow(self.mean, 2))


class PointEmbed(nn.Module):
    def __init__(self, dim=512, freq_embed_dim=48):
        super().__init__()
        assert freq_embed_dim % 6 == 0
        n = freq_embed_dim // 6
        octaves = torch.pow(2, torch.arange(n)).float() * np.pi
        basis = torch.zeros(3, 3 * n)
        for axis in range(3):
            basis[axis, axis * n:(axis + 1) * n] = octaves
        self.register_buffer('basis', basis)
        self.mlp = nn.Linear(freq_embed_dim + 3, dim)


class GEGLU(nn.Module):
    pass


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), GEGLU(), nn.Linear(dim * mult, dim))


class _CrossAttentionParams(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.q_proj = nn.Linear(dim, dim)
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)


class ResCrossAttBlock(nn.Module):
    def __init__(self, dim, num_heads, gradient_checkpointing=True):
        super().__init__()
        self.ln1 = nn.LayerNorm(dim)
        self.att = _CrossAttentionParams(dim, num_heads)
        self.ln2 = nn.LayerNorm(dim)
        self.mlp = FeedForward(dim)


class PointEncoderEmbed(nn.Module):
    def __init__(self, hidden_dim=1024, num_heads=16, latent_size=2048, latent_dim=64, gradient_checkpointing=True):
        super().__init__()
        self.latent_size = latent_size
        self.query_embed = nn.Parameter(torch.randn(1, latent_size, hidden_dim) / hidden_dim ** 0.5)
        self.point_embed = PointEmbed(dim=hidden_dim)
        self.ln = nn.LayerNorm(hidden_dim)
        self.cross_att = ResCrossAttBlock(hidden_dim, num_heads, gradient_checkpointing)
        self.linear = nn.Linear(hidden_dim, latent_dim)
