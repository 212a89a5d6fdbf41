"""Deterministic synthetic weights and inputs (no network: there are no checkpoints or datasets).

``state_dict_spec`` enumerates the reference checkpoint schema (verified against
``LMM(opt).state_dict()`` of /root/reference by ``oracle/gen_golden.py``), ``synth_state_dict`` fills
it from a seeded CPU generator.  The SAME tensors are loaded into the reference model (when golden
vectors are generated in the build container), into the CPU oracle and into the B200 engine, so every
parity comparison starts from bit-identical fp16 weights.

Distributions follow the reference initialisers (``core/transformer/modeling_opt.py:443-458``,
torch defaults for the encoder) except that biases and LayerNorm affines are randomised so that a
kernel that drops one of them fails parity.  ``eos_logit`` plants a constant EOS logit (see below)
so that run length is pinned by ``max_new_tokens`` on both sides (SURVEY.md §7 "hard parts").
"""

from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch


def vocab_size_of(opt) -> int:
    """``core/models.py:77-84``."""
    if opt.use_meto:
        if opt.meto_backend == 'LR':
            return 2 * opt.discrete_bins + 3 + 3
        return opt.discrete_bins + 3 + 3
    return opt.discrete_bins + 3


def state_dict_spec(opt) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered (name, shape, init-kind) for cond_mode in {'point', 'point_latent', 'none'}."""
    C = opt.hidden_dim
    F = opt.hidden_dim * 4 if opt.intermediate_dim is None else opt.intermediate_dim
    V = vocab_size_of(opt)
    P = opt.max_seq_length + opt.num_cond_tokens + 10
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def lin(prefix, out_f, in_f, kind, bias=True):
        spec.append((prefix + '.weight', (out_f, in_f), kind))
        if bias:
            spec.append((prefix + '.bias', (out_f,), 'bias:' + kind))

    def ln(prefix, dim):
        spec.append((prefix + '.weight', (dim,), 'ln_w'))
        spec.append((prefix + '.bias', (dim,), 'ln_b'))

    if opt.cond_mode == 'point':
        E, Ld = opt.point_hidden_dim, opt.point_latent_dim
        pe = 'point_encoder.'
        spec.append((pe + 'query_embed', (1, opt.point_latent_size, E), 'normal:%r' % (1.0 / math.sqrt(E))))
        spec.append((pe + 'point_embed.basis', (3, 24), 'basis'))
        lin(pe + 'point_embed.mlp', E, 51, 'uniform')
        ln(pe + 'ln', E)
        ln(pe + 'cross_att.ln1', E)
        for p in ('q_proj', 'k_proj', 'v_proj', 'out_proj'):
            lin(pe + 'cross_att.att.' + p, E, E, 'uniform')
        ln(pe + 'cross_att.ln2', E)
        lin(pe + 'cross_att.mlp.net.0', E * 8, E, 'uniform')
        lin(pe + 'cross_att.mlp.net.2', E, E * 4, 'uniform')
        lin(pe + 'linear', Ld, E, 'uniform')
    if opt.cond_mode in ('point', 'point_latent'):
        lin('proj_cond', C, opt.point_latent_dim, 'uniform')
        ln('norm_cond', C)
    if opt.use_num_face_cond:
        spec.append(('embed_num_face.weight', (10, C), 'normal:1.0'))

    md = 'mesh_decoder.model.'
    spec.append((md + 'embd.weight', (V, C), 'embd'))
    spec.append((md + 'embed_positions.weight', (P, C), 'normal:0.02'))
    out_std = 0.02 / math.sqrt(2 * opt.num_layers)
    for i in range(opt.num_layers):
        lp = md + 'layers.%d.' % i
        lin(lp + 'self_attn.k_proj', C, C, 'normal:0.02')
        lin(lp + 'self_attn.v_proj', C, C, 'normal:0.02')
        lin(lp + 'self_attn.q_proj', C, C, 'normal:0.02')
        lin(lp + 'self_attn.out_proj', C, C, 'normal:%r' % out_std)
        ln(lp + 'self_attn_layer_norm', C)
        lin(lp + 'fc1', F, C, 'normal:0.02')
        lin(lp + 'fc2', C, F, 'normal:0.02')
        ln(lp + 'final_layer_norm', C)
    spec.append(('mesh_decoder.lm_head.weight', (V, C), 'normal:0.02'))
    return spec


def fourier_basis() -> torch.Tensor:
    """``core/transformer/point.py:43-50``: [3, 24] block-diagonal 2^k * pi, k = 0..7."""
    e = torch.pow(2, torch.arange(8)).float() * np.pi
    z = torch.zeros(8)
    return torch.stack([torch.cat([e, z, z]), torch.cat([z, e, z]), torch.cat([z, z, e])])


def synth_state_dict(opt, seed: int = 0, eos_logit: float | None = -30.0,
                     dtype: torch.dtype = torch.float32, fp16_exact: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint in the reference key schema.

    ``eos_logit``: if not None, the last ``final_layer_norm`` gets weight=1, bias=1 and the EOS row
    of ``lm_head`` becomes ``eos_logit / C`` everywhere.  LayerNorm output sums to zero, so the EOS
    logit is the constant ``eos_logit`` (+ rounding) at every step: greedy / top-k sampling never emit
    EOS and the run length equals ``max_new_tokens`` on both the reference and this implementation.
    """
    g = torch.Generator(device='cpu')
    sd: Dict[str, torch.Tensor] = {}
    spec = state_dict_spec(opt)
    for idx, (name, shape, kind) in enumerate(spec):
        g.manual_seed((seed * 1000003 + idx * 7919 + 12345) & 0x7FFFFFFF)
        if kind == 'basis':
            t = fourier_basis()
        elif kind.startswith('normal:'):
            t = torch.randn(shape, generator=g) * float(kind.split(':')[1])
        elif kind == 'uniform':
            bound = 1.0 / math.sqrt(shape[-1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind.startswith('bias:'):
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == 'ln_w':
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == 'ln_b':
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == 'embd':
            t = torch.randn(shape, generator=g) * 0.02
            t[opt.pad_token_id].zero_()
        else:  # pragma: no cover
            raise ValueError(kind)
        # every value is fp16-representable: `model.half()` (infer.py:56) is then exact, so the reference's fp32 CPU
        # run (goldens) and the fp16 GPU path start from bit-identical weights and differ by activation rounding only
        if fp16_exact and kind != 'basis':
            t = t.to(torch.float16).to(torch.float32)
        sd[name] = t.to(dtype).contiguous()
    if eos_logit is not None:
        last = 'mesh_decoder.model.layers.%d.final_layer_norm.' % (opt.num_layers - 1)
        sd[last + 'weight'] = torch.ones_like(sd[last + 'weight'])
        sd[last + 'bias'] = torch.ones_like(sd[last + 'bias'])
        C = opt.hidden_dim
        sd['mesh_decoder.lm_head.weight'][opt.eos_token_id] = float(torch.tensor(eos_logit / C).half()) if fp16_exact else eos_logit / C
    return sd


def synth_point_cloud(seed: int = 0, n: int = 8192) -> torch.Tensor:
    """Uniform cloud in the normalised cube (``infer.py:88`` normalises to bound 0.95): [1, n, 3] fp32."""
    g = torch.Generator(device='cpu')
    g.manual_seed(1234567 + seed)
    return torch.rand((1, n, 3), generator=g) * 1.9 - 0.95


def tiny_options(**over):
    """A small ArAE-shaped configuration that the CPU oracle finishes in seconds.

    head_dim stays 96 (decoder) / 64 (encoder) because the kernels are specialised on those."""
    from core.options import config_defaults
    from dataclasses import replace
    base = dict(hidden_dim=192, num_heads=2, num_layers=2, point_hidden_dim=128, point_num_heads=2,
                point_latent_size=64, point_latent_dim=16, point_num=256, num_cond_tokens=65,
                max_seq_length=512, generate_mode='greedy')
    base.update(over)
    return replace(config_defaults['ArAE'], **base)


def synth_dit_state_dict(hidden_dim, num_heads, latent_size, latent_dim, num_layers, cond_dim=None, seed=0, gain=1.0):
    """Seeded synthetic weights with the reference DiT's state-dict schema (+ proj_cond / norm_cond when cond_dim is given), scaled so
    that activations stay O(1) through the stack (there is no pretrained checkpoint offline)."""
    g = torch.Generator().manual_seed(seed)
    C = hidden_dim
    sd = {}

    def lin(name, out_f, in_f, s=1.0):
        sd[name + '.weight'] = torch.randn(out_f, in_f, generator=g) * (s * gain / math.sqrt(in_f))
        sd[name + '.bias'] = torch.randn(out_f, generator=g) * 0.02

    lin('proj_in', C, latent_dim)
    sd['pos_embed'] = torch.randn(1, latent_size, C, generator=g) / math.sqrt(C)
    lin('timestep_proj.linear_1', C, 256)
    lin('timestep_proj.linear_2', C, C)
    lin('adaln_linear', 6 * C, C, 0.5)
    for l in range(num_layers):
        p = f'layers.{l}.'
        lin(p + 'attn1.qkv_proj', 3 * C, C)
        lin(p + 'attn1.out_proj', C, C)
        for n in ('q_proj', 'k_proj', 'v_proj', 'out_proj'):
            lin(p + 'attn2.' + n, C, C)
        lin(p + 'ff.net.0', 8 * C, C)
        lin(p + 'ff.net.2', C, 4 * C)
        sd[p + 'scale_shift_table'] = torch.randn(6, C, generator=g) / math.sqrt(C)
    sd['scale_shift_table'] = torch.randn(2, C, generator=g) / math.sqrt(C)
    lin('proj_out', latent_dim, C)
    if cond_dim is not None:
        out = {'dit.' + k: v for k, v in sd.items()}
        w = torch.randn(C, cond_dim, generator=g) / math.sqrt(cond_dim)
        out.update({'proj_cond.weight': w, 'proj_cond.bias': torch.randn(C, generator=g) * 0.02,
                    'norm_cond.weight': 1 + 0.1 * torch.randn(C, generator=g), 'norm_cond.bias': 0.05 * torch.randn(C, generator=g)})
        return out
    return sd
