"""CPU ORACLE of the DiT denoiser path — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import this module; the product
(``core.models_dit.MDiT`` -> ``edgerunner_b200.dit_engine`` -> CUDA) never does.

A plain functional restatement (torch tensor ops, no ``nn.Module``; runs on whatever device the tensors are on) of

    dit_forward     core/transformer/dit.py:165-196 (DiT.forward), :118-138 (DiTLayer._forward), :45-76 (Timesteps),
                    :79-97 (TimestepEmbedding), :26-42 (GEGLU feed-forward); core/transformer/attention.py:98-152
                    (SelfAttention / CrossAttention), :27-62 (attention, non-causal)
    get_cond_adaptor core/models_dit.py:116 (norm_cond(proj_cond(h)))
    ddim_tables / ddim_step / sample_loop
                    core/models_dit.py:184-229 (MDiT.run) + diffusers==0.30.2 DDIMScheduler (third-party, un-vendored,
                    requirements.lock.txt:18; NOT installed in this image): scaled-linear betas, "leading" spacing with
                    steps_offset 1, eta = 0 update for epsilon / v_prediction.

Pinning: ``oracle/gen_golden.py --only dit`` executes the REFERENCE ``DiT`` module itself (imported from /root/reference, CPU,
fp32, naive attention) on seeded weights and commits input / output to ``tests/golden/dit.npz``; ``tests/test_dit_cpu.py`` checks
``mode='fp32'`` of this file against it.  ``mode='ledger'`` adds the fp16 rounding points of ``model_dit.half()`` +
``torch.autocast('cuda', fp16)`` (infer_dit.py:70,106): Linear outputs fp16, LayerNorm outputs fp32, fp16 (op) fp16 -> fp16,
fp32 (op) fp16 -> fp32.  On the GPU box ``tests/test_gpu_dit.py`` compares the CUDA engine with the reference module ITSELF
(oracle/_ref/py, .half() + autocast + flash-attn), which pins the ledger.  The scheduler part has no reference to run against
(diffusers is absent here and on the GPU box): **parity unpinned for ddim_tables / ddim_step** — restated from the published algorithm.
"""

from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


def _r16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float16).to(torch.float32)


class DitOracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], num_heads: int, mode: str = 'ledger', device='cpu'):
        """state_dict: keys of the reference ``DiT`` module (no prefix) or of ``MDiT`` (``dit.`` prefix + proj_cond / norm_cond)."""
        assert mode in ('fp32', 'ledger')
        self.ledger = mode == 'ledger'
        self.H = num_heads
        self.w = {}
        for k, v in state_dict.items():
            v = v.detach().to(device=device, dtype=torch.float32)
            self.w[k[4:] if k.startswith('dit.') else k] = _r16(v) if self.ledger else v          # model.half()
        self.NL = 1 + max(int(k.split('.')[1]) for k in self.w if k.startswith('layers.'))

    def r(self, t):
        return _r16(t) if self.ledger else t

    def linear(self, x, name):
        """nn.Linear under autocast: input cast to fp16, fp32 accumulate + bias, one rounding to fp16."""
        return self.r(self.r(x) @ self.w[name + '.weight'].t() + self.w[name + '.bias'])

    def attention(self, q, k, v):
        """q [B,N,H,D], k/v [B,M,H,D] -> [B,N,H*D] (attention.py:47-62, non-causal; fp32 softmax, fp16 output)."""
        s = torch.einsum('bnhd,bmhd->bhnm', q, k) / math.sqrt(q.shape[-1])
        return self.r(torch.einsum('bhnm,bmhd->bnhd', torch.softmax(s, dim=-1), v)).flatten(2)

    def timestep_features(self, t):
        """Timesteps(256).forward (dit.py:54-76): fp32."""
        half = 128
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
        emb = t[:, None].float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)

    def silu(self, x):
        return self.r(F.silu(x))

    def t_embed(self, t):
        """-> (t_emb [B,C], t_adaln [B,6,C])  (dit.py:178-180)."""
        e = self.linear(self.timestep_features(t), 'timestep_proj.linear_1')
        e = self.linear(self.silu(e), 'timestep_proj.linear_2')
        return e, self.linear(self.silu(e), 'adaln_linear').view(t.shape[0], 6, -1)

    def modulate(self, x, scale, shift, eps=1e-6):
        """norm(x) * (1 + scale) + shift: LayerNorm in fp32 (autocast), `1 + scale` in fp16, product and sum in fp32."""
        xn = F.layer_norm(x, (x.shape[-1],), None, None, eps)
        return xn * self.r(1 + scale) + shift

    def layer(self, l, x, c, t_adaln):
        p = f'layers.{l}.'
        B, N, C = x.shape
        H, D = self.H, C // self.H
        mod = self.r(self.w[p + 'scale_shift_table'][None] + t_adaln)              # [B,6,C]
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = [m for m in mod.split(1, dim=1)]
        x = self.modulate(x, scale_msa, shift_msa)                                   # NOTE: x is re-bound (dit.py:127-128)
        qkv = self.linear(x, p + 'attn1.qkv_proj').view(B, N, 3, H, D)
        a = self.linear(self.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]), p + 'attn1.out_proj')
        x = x + self.r(gate_msa * a)
        q = self.linear(x, p + 'attn2.q_proj').view(B, N, H, D)
        k = self.linear(c, p + 'attn2.k_proj').view(B, -1, H, D)
        v = self.linear(c, p + 'attn2.v_proj').view(B, -1, H, D)
        x = x + self.linear(self.attention(q, k, v), p + 'attn2.out_proj')
        x = self.modulate(x, scale_mlp, shift_mlp)
        h = self.linear(x, p + 'ff.net.0')
        a, g = h.chunk(2, dim=-1)
        h = self.r(a * self.r(F.gelu(g)))
        x = x + self.r(gate_mlp * self.linear(h, p + 'ff.net.2'))
        return x

    def forward(self, x, c, t):
        """x [B,N,Dl] fp32, c [B,M,C] fp32, t [B] -> [B,N,Dl] (fp16-valued in ledger mode)."""
        x = self.r(self.linear(x, 'proj_in') + self.w['pos_embed'])
        t_emb, t_adaln = self.t_embed(t)
        for l in range(self.NL):
            x = self.layer(l, x, c, t_adaln)
        mod = self.r(self.w['scale_shift_table'][None] + t_emb[:, None])          # [B,2,C]
        shift, scale = mod[:, 0:1], mod[:, 1:2]
        return self.linear(self.modulate(x, scale, shift), 'proj_out')

    def cond_adaptor(self, h):
        """models_dit.py:116: norm_cond(proj_cond(h)) -> fp32."""
        y = self.linear(h, 'proj_cond')
        return F.layer_norm(y, (y.shape[-1],), self.w['norm_cond.weight'], self.w['norm_cond.bias'], 1e-5)

    # ---- sampling loop ------------------------------------------------------------------------------------------------------
    def sample_loop(self, cond, latents, timesteps, coef, guidance_scale=7.5, prediction_type='v_prediction'):
        """MDiT.run's loop (models_dit.py:209-227) with the dtypes of the fp16 path: noise_pred fp16, guidance in fp16, latents fp32."""
        lat = latents.clone().float()
        cc = torch.cat([torch.zeros_like(cond), cond], dim=0)
        for i, t in enumerate(timesteps):
            t_in = torch.full((2 * lat.shape[0],), float(t), dtype=torch.float32, device=lat.device)
            pred = self.forward(torch.cat([lat] * 2, dim=0), cc, t_in)
            u, c = pred.chunk(2)
            m = self.r(u + self.r(guidance_scale * self.r(c - u)))
            lat = ddim_step(m, lat, coef[i], prediction_type, self.ledger)
        return lat


def ddim_tables(num_inference_steps, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
    """-> (timesteps int64 [S], coef fp32 [S,4] = sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)); set_alpha_to_one=False."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    ratio = num_train_timesteps // num_inference_steps
    ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset
    coef = torch.empty(len(ts), 4, dtype=torch.float32)
    for i, t in enumerate(ts.tolist()):
        prev = t - ratio
        a_t, a_prev = ac[t], (ac[prev] if prev >= 0 else ac[0])
        coef[i] = torch.stack([a_t ** 0.5, (1 - a_t) ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5])
    return ts, coef


def ddim_step(model_output, sample, k, prediction_type='v_prediction', ledger=True):
    """DDIMScheduler.step, eta = 0, no clipping.  In the fp16 path `model_output` is fp16 and the scalars are 0-dim fp32 tensors: a 0-dim
    tensor does not promote, so (scalar * model_output) is rounded to fp16, while anything touching `sample` (fp32) is fp32."""
    r = _r16 if ledger else (lambda z: z)
    sa, sb, sap, sdir = [float(v) for v in k]
    sa32, sb32, sap32, sdir32 = [torch.tensor(v, dtype=torch.float32) for v in (sa, sb, sap, sdir)]
    if prediction_type == 'v_prediction':
        x0 = sa32 * sample - r(sb32 * model_output)
        eps = r(sa32 * model_output) + sb32 * sample
        return sap32 * x0 + sdir32 * eps
    if prediction_type == 'epsilon':
        x0 = (sample - r(sb32 * model_output)) / sa32
        return sap32 * x0 + r(sdir32 * model_output)
    raise ValueError(prediction_type)


def synth_dit_state(*a, **k):
    """Seeded synthetic DiT weights (edgerunner_b200/synth.py::synth_dit_state_dict — shared with bench.py, which may not import oracle/)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from edgerunner_b200.synth import synth_dit_state_dict
    return synth_dit_state_dict(*a, **k)
