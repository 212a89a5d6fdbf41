"""``DiT`` — the latent denoiser of the image-conditioned path.  Mirrors ``/root/reference/core/transformer/dit.py``
(``GEGLU`` :26, ``FeedForward`` :32, ``Timesteps`` :45, ``TimestepEmbedding`` :79, ``DiTLayer`` :100, ``DiT`` :141): same
constructor arguments, parameter names and shapes (so the reference's checkpoints load), same ``forward(x, c, t)``.

The modules here only hold parameters.  ``DiT.forward`` hands device pointers to the sm_100a library
(``edgerunner_b200/csrc/dit.cu``: tcgen05 GEMMs and flash attention, fused LayerNorm + adaLN modulation, gated residuals) and
returns what the reference returns under ``torch.autocast('cuda', fp16)`` with ``.half()`` weights (infer_dit.py:70,106): fp16
``[B, N, latent_dim]``.  Inference only: no autograd graph, no CPU fallback.
"""

import torch
import torch.nn as nn


class GEGLU(nn.Module):
    pass


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), GEGLU(), nn.Linear(dim * mult, dim))


class Timesteps(nn.Module):
    """Sinusoidal timestep features; parameter-free (computed by ``dit_timestep_kernel``)."""

    def __init__(self, num_channels=256, flip_sin_to_cos=False, downscale_freq_shift=0, scale=1, max_period=10000):
        super().__init__()
        if (num_channels, flip_sin_to_cos, downscale_freq_shift, scale, max_period) != (256, False, 0, 1, 10000):
            raise NotImplementedError('only the configuration DiT uses (256 channels, sin | cos, period 10000) is implemented')
        self.num_channels = num_channels


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, sample_proj_bias=True):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim, sample_proj_bias)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim, sample_proj_bias)


class _SelfAttentionParams(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv_proj = nn.Linear(dim, 3 * dim)
        self.out_proj = nn.Linear(dim, dim)


class _CrossAttentionParams(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.q_proj = nn.Linear(dim, dim)
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)


class DiTLayer(nn.Module):
    """PixArt-alpha style block: adaLN-modulated self-attention, cross-attention to the condition, adaLN-modulated GEGLU FF."""

    def __init__(self, dim, num_heads, gradient_checkpointing=True):
        super().__init__()
        self.dim, self.num_heads, self.gradient_checkpointing = dim, num_heads, gradient_checkpointing
        self.norm1 = nn.LayerNorm(dim, eps=1e-6, elementwise_affine=False)
        self.attn1 = _SelfAttentionParams(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6, elementwise_affine=False)
        self.attn2 = _CrossAttentionParams(dim, num_heads)
        self.ff = FeedForward(dim)
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim ** 0.5)


class DiT(nn.Module):
    def __init__(self, hidden_dim=1024, num_heads=16, latent_size=2048, latent_dim=64, num_layers=24, gradient_checkpointing=True,
                 cond_tokens=257, cond_dim=1280):
        super().__init__()
        self.hidden_dim, self.num_heads, self.latent_size, self.latent_dim, self.num_layers = hidden_dim, num_heads, latent_size, latent_dim, num_layers
        self.cond_tokens, self.cond_dim = cond_tokens, cond_dim          # only sizes the engine's adaptor slots (MDiT.proj_cond)
        self.proj_in = nn.Linear(latent_dim, hidden_dim)
        self.pos_embed = nn.Parameter(torch.randn(1, latent_size, hidden_dim) / hidden_dim ** 0.5)
        self.timestep_embed = Timesteps(num_channels=256)
        self.timestep_proj = TimestepEmbedding(256, hidden_dim)
        self.adaln_linear = nn.Linear(hidden_dim, hidden_dim * 6, bias=True)
        self.layers = nn.ModuleList([DiTLayer(hidden_dim, num_heads, gradient_checkpointing) for _ in range(num_layers)])
        self.norm_out = nn.LayerNorm(hidden_dim, eps=1e-6, elementwise_affine=False)
        self.scale_shift_table = nn.Parameter(torch.randn(2, hidden_dim) / hidden_dim ** 0.5)
        self.proj_out = nn.Linear(hidden_dim, latent_dim)
        self._engine = None
        self._engine_key = None

    # ---- engine management (stand-alone use; MDiT shares one engine that also holds its adaptor weights) ------------------
    def _fingerprint(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def get_engine(self, extra_state=None):
        from edgerunner_b200.dit_engine import DiTEngine
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError('DiT: parameters are not on a CUDA device; the B200 path has no CPU fallback (call .to("cuda") first)')
        key = (p0.device, self._fingerprint(), None if extra_state is None else tuple((v.data_ptr(), v._version) for v in extra_state.values()))
        if self._engine is None or self._engine.device != p0.device:
            self._engine = DiTEngine(p0.device, self.hidden_dim, self.num_heads, self.num_layers, self.latent_size, self.latent_dim,
                                     self.cond_tokens, self.cond_dim)
            self._engine_key = None
        if self._engine_key != key:
            sd = {'dit.' + k: v for k, v in self.state_dict().items()}
            if extra_state is None:        # stand-alone DiT: the adaptor slots are unused, fill them with zeros
                C = self.hidden_dim
                z = lambda *s: torch.zeros(*s, dtype=torch.float16, device=p0.device)
                extra_state = {'proj_cond.weight': z(C, self.cond_dim), 'proj_cond.bias': z(C), 'norm_cond.weight': z(C), 'norm_cond.bias': z(C)}
            sd.update(extra_state)
            self._engine.load_state_dict(sd)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def forward(self, x, c, t):
        """x [B, N, latent_dim] latents, c [B, M, hidden_dim] condition (normed + projected), t [B] timesteps -> [B, N, latent_dim] fp16."""
        if self.training:
            raise NotImplementedError('training-mode DiT forward (gradient checkpointing, autograd graph) is not on the B200 path: call .eval()')
        if c.shape[1] != self.cond_tokens:
            raise NotImplementedError(f'the engine was sized for {self.cond_tokens} condition tokens, got {c.shape[1]}')
        return self.get_engine().forward(x, c, t)
