"""CPU ORACLE of the training step — TEST INFRASTRUCTURE ONLY (see oracle/er_oracle.py for the rules).

What the reference does in a training step (main.py:168-172): ``out = model(data); accelerator.backward(out['loss'])`` — torch autograd over
``LMM.forward`` (core/models.py:147-202) in ``model.train()`` mode:

    encode_cond        core/models.py:101-144; with ``opt.freeze_encoder`` the point encoder runs under ``torch.no_grad()`` (:105, :115-117)
    decoder layer      core/transformer/modeling_opt.py:253-298 — post-LN; ``F.dropout(p=config.dropout)`` on the attention branch (:272) and on the
                       MLP branch (:285) before the residual adds
    padded batches     core/transformer/attention.py:65-93 (unpad -> varlen causal -> pad_input: masked rows come back as zeros)
    loss               modeling_opt.py:497-505 (lm_head, shift, cross_entropy(ignore_index=-100)); models.py:191-197 (+ kl_weight * KL)

``forward_train`` restates that forward with plain differentiable torch ops over a dict of leaf tensors, so that ``loss.backward()`` IS the
reference's backward (autograd).  Pinning: with ``dropout_p = 0`` and no num-face dropout its loss must equal ``Oracle.forward_tf`` (itself pinned
against the reference modules by tests/golden) — tests/test_train_cpu.py checks that here on the CPU — and its GRADIENTS are pinned against the
reference's own `loss.backward()` executed on the CPU (oracle/gen_golden.py --only train -> tests/golden/train.npz: every parameter's gradient norm and
probe values, cond_mode 'point' with the encoder trained and 'point_latent').  The dropout mask cannot be torch's (the CUDA
path draws it from a counter-based generator, edgerunner_b200/csrc/backward.cu::drop_keep): ``dropout_keep`` restates that generator bit for bit so
that oracle and engine drop the same elements.
"""

from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from oracle.er_oracle import Oracle, quantize_num_faces

_M64 = (1 << 64) - 1


def dropout_keep(seed: int, site: int, n: int, p: float) -> np.ndarray:
    """keep[i] for i in [0, n): splitmix64 of (seed, site, i), upper 32 bits >= p * 2^32  (backward.cu::drop_keep)."""
    if p <= 0:
        return np.ones(n, dtype=bool)
    thr = min(int(float(np.float32(p)) * 4294967296.0), 4294967295)
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over='ignore'):
        z = np.uint64(seed & _M64) + np.uint64(0x9E3779B97F4A7C15) * idx + np.uint64((0xD1B54A32D192ED03 * (site + 1)) & _M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint64) >= np.uint64(thr)


def _dropout(y: torch.Tensor, seed: int, site: int, p: float) -> torch.Tensor:
    if p <= 0:
        return y
    keep = torch.from_numpy(dropout_keep(seed, site, y.numel(), p)).view(y.shape)
    return torch.where(keep, y / (1.0 - float(np.float32(p))), torch.zeros_like(y))


def trainable_leaves(state_dict: Dict[str, torch.Tensor], train_encoder: bool = False, round_fp16: bool = True) -> Dict[str, torch.Tensor]:
    """fp16-rounded copies (what the engine holds) of every trainable tensor as fp32 leaves with requires_grad: everything but the point encoder, or
    (train_encoder, opt.freeze_encoder = False) its parameters too — `point_embed.basis` is a buffer, not a parameter (point.py:50)."""
    return {k: (v.detach().to(torch.float16).float() if round_fp16 else v.detach().float().clone()).requires_grad_(True) for k, v in state_dict.items()
            if (train_encoder and k != 'point_encoder.point_embed.basis') or not k.startswith('point_encoder.')}


def encode_points_train(opt, state_dict, w: Dict[str, torch.Tensor], pc: torch.Tensor, round_embed: bool = True) -> torch.Tensor:
    """PointEncoderEmbed.forward (core/transformer/point.py:186-206; ResCrossAttBlock :117-126; GEGLU FFN :74-84) as differentiable fp32 ops over the
    leaves `w`.  The Fourier features are data (no parameter in front of them): they are taken with the fp16 rounding the autocast forward applies
    (Oracle.encode_points) — at |x * basis| up to 400 rad an unrounded product would be a different input, not a more accurate one."""
    pe = 'point_encoder.'
    r = _r16 if round_embed else (lambda t: t)          # round_embed = False: the reference's CPU / fp32 arithmetic (the pinning test)
    basis = r(state_dict[pe + 'point_embed.basis'].float())
    E = w[pe + 'query_embed'].shape[-1]
    Hh = opt.point_num_heads
    Dh = E // Hh

    def lin(x, name):
        return x @ w[pe + name + '.weight'].t() + w[pe + name + '.bias']

    def ln(x, name):
        return F.layer_norm(x, (x.shape[-1],), w[pe + name + '.weight'], w[pe + name + '.bias'], 1e-5)
    outs = []
    for b in range(pc.shape[0]):
        x = pc[b].float()
        with torch.no_grad():
            proj = r(r(x) @ basis)
            emb = torch.cat([r(torch.sin(proj)), r(torch.cos(proj)), r(x)], dim=1)
        kvx = ln(lin(emb, 'point_embed.mlp'), 'ln')
        q0 = w[pe + 'query_embed'][0]
        q = lin(ln(q0, 'cross_att.ln1'), 'cross_att.att.q_proj').view(-1, Hh, Dh).transpose(0, 1)
        k = lin(kvx, 'cross_att.att.k_proj').view(-1, Hh, Dh).transpose(0, 1)
        v = lin(kvx, 'cross_att.att.v_proj').view(-1, Hh, Dh).transpose(0, 1)
        a = (torch.softmax(q @ k.transpose(-1, -2) / Dh ** 0.5, -1) @ v).transpose(0, 1).reshape(-1, E)
        x1 = q0 + lin(a, 'cross_att.att.out_proj')
        h = lin(ln(x1, 'cross_att.ln2'), 'cross_att.mlp.net.0')
        a_, g_ = h.chunk(2, dim=-1)
        x2 = x1 + lin(a_ * F.gelu(g_), 'cross_att.mlp.net.2')
        outs.append(lin(x2, 'linear'))
    return torch.stack(outs)


def _r16(t):
    return t.to(torch.float16).to(torch.float32)


def forward_train(opt, state_dict, w: Dict[str, torch.Tensor], conds, tokens, labels, num_faces, masks: Optional[torch.Tensor] = None,
                  dropout_p: float = 0.0, seed: int = 0, train_encoder: bool = False, exact_fp32: bool = False):
    """-> dict(loss, loss_ce, loss_kl).  ``w``: trainable leaves (``trainable_leaves``); the frozen point encoder comes from ``state_dict``.
    fp32 arithmetic throughout (the gradient reference; the engine's fp16 rounding is what the test tolerance covers).  exact_fp32: no fp16 rounding
    of inputs either (Fourier features, latents) — the reference's own CPU arithmetic, used to pin this function against tests/golden/train.npz."""
    orc = Oracle(opt, state_dict, mode='fp32' if exact_fp32 else 'ledger')     # frozen encoder as the engine runs it (fp16 ledger), no graph
    B, T = tokens.shape
    C, H, NL = opt.hidden_dim, opt.num_heads, opt.num_layers
    D = C // H
    if train_encoder:                                                # opt.freeze_encoder = False: the encoder and the KL term are part of the graph
        assert opt.cond_mode == 'point'
        lat_all = encode_points_train(opt, state_dict, w, conds, round_embed=not exact_fp32)
    else:
        with torch.no_grad():
            lat_all = orc.encode_points(conds) if opt.cond_mode == 'point' else (conds.float() if exact_fp32 else orc.r(conds.float()))
    xs = []
    for b in range(B):
        ce = F.layer_norm(lat_all[b] @ w['proj_cond.weight'].t() + w['proj_cond.bias'], (C,), w['norm_cond.weight'], w['norm_cond.bias'], 1e-5)
        if opt.use_num_face_cond:
            ce = torch.cat([ce, w['embed_num_face.weight'][quantize_num_faces(int(num_faces[b]))][None]], 0)
        tok = w['mesh_decoder.model.embd.weight'][tokens[b].long()]
        x = torch.cat([ce, tok], 0)
        xs.append(x + w['mesh_decoder.model.embed_positions.weight'][:x.shape[0]])
    h = torch.stack(xs)                                               # [B, N, C]
    N = h.shape[1]
    causal = torch.triu(torch.ones(N, N, dtype=torch.bool), 1)
    valid = torch.ones(B, N, dtype=torch.bool) if masks is None else masks.bool()
    md = 'mesh_decoder.model.layers.'
    for i in range(NL):
        lp = md + '%d.' % i

        def lin(x, name):
            return x @ w[lp + name + '.weight'].t() + w[lp + name + '.bias']
        q = lin(h, 'self_attn.q_proj').view(B, N, H, D).transpose(1, 2)
        k = lin(h, 'self_attn.k_proj').view(B, N, H, D).transpose(1, 2)
        v = lin(h, 'self_attn.v_proj').view(B, N, H, D).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / (D ** 0.5)
        s = s.masked_fill(causal, float('-inf'))
        a = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, N, C)
        a = a * valid[..., None]                                      # pad_input: masked rows are zeros
        o = _dropout(lin(a, 'self_attn.out_proj'), seed, 2 * i, dropout_p)
        h = F.layer_norm(h + o, (C,), w[lp + 'self_attn_layer_norm.weight'], w[lp + 'self_attn_layer_norm.bias'], 1e-5)
        f = _dropout(lin(torch.relu(lin(h, 'fc1')), 'fc2'), seed, 2 * i + 1, dropout_p)
        h = F.layer_norm(h + f, (C,), w[lp + 'final_layer_norm.weight'], w[lp + 'final_layer_norm.bias'], 1e-5)
    logits = h @ w['mesh_decoder.lm_head.weight'].t()
    V = logits.shape[-1]
    loss_ce = F.cross_entropy(logits[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1).long(), ignore_index=-100)
    out = dict(loss_ce=loss_ce, loss=loss_ce)
    if opt.cond_mode == 'point':
        out['loss_kl'] = 0.5 * torch.sum(lat_all.float() ** 2)
        out['loss'] = loss_ce + opt.kl_weight * out['loss_kl']
    return out
