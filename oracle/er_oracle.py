"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product (``core.models.LMM`` -> ``edgerunner_b200`` -> CUDA) never
does and fails loudly when its CUDA library is missing.

What it is: a plain functional restatement (torch CPU tensor ops, no ``nn.Module``) of the reference's
mesh-token decode path, each function citing the reference file:line it follows:

    encode_points   core/transformer/point.py:186-206 (PointEncoderEmbed), :54-65 (PointEmbed),
                    :117-126 (ResCrossAttBlock), :74-84 (GEGLU FFN); core/transformer/attention.py:141-153
    encode_cond     core/models.py:101-144 ; core/utils.py:109-136 (quantize_num_faces)
    decoder_rows    core/transformer/modeling_opt.py:321-426 (ShapeOPTDecoder.forward),
                    :253-298 (post-LN OPTDecoderLayer), :172-237 (attention + KV cache), :497 (lm_head)
    attention       core/transformer/attention.py:27-62 (causal requires N==1 or N==M)
    constraint FSM  core/models.py:245-271 ; mask application core/utils.py:143-158
    greedy / sample HF transformers==4.46.2 GenerationMixin._sample (third-party, un-vendored;
                    requirements.lock.txt:16): logits[:, -1].float() -> prefix-constraint mask ->
                    (sample: top-k 10 keep-ties -> softmax -> multinomial) | argmax -> stop at EOS / max_new
    forward_tf      core/models.py:147-202 + modeling_opt.py:499-505 (shifted CE, ignore_index=-100)

Pinning: the reference has no golden vectors.  ``oracle/gen_golden.py`` executes the
REFERENCE modules themselves (imported from /root/reference in the build container, CPU, fp32, naive
attention path) on the seeded synthetic weights of ``edgerunner_b200.synth`` and commits their outputs to
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle in ``mode='fp32'`` against them.
``mode='ledger'`` re-runs the same algorithm with the fp16 rounding points that ``infer.py``'s
``model.half()`` + ``torch.autocast('cuda', fp16)`` produce on a GPU.  The
build container has no GPU, so the ledger is pinned ON THE GPU BOX instead (round 2): the reference's own
modules travel there in the git-ignored ``oracle/_ref/py`` and ``tests/test_gpu_reference.py`` /
``scripts/ref_gpu.py`` run them as ``infer.py`` does, record their forward-hook dtype ledger (asserted to be
the one implemented here) and compare logits teacher-forced (mean |d| 8.8e-4 over 4000 steps,
profiles/r02_reference_gpu_path.json).  The HF loop remains a restatement (parity unpinned at the HF
boundary: the installed transformers 5.5 cannot drive the reference's tuple KV cache; the pinned 4.46.2
is not installable offline).
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _r16(t: torch.Tensor) -> torch.Tensor:
    """Round to fp16 and return as fp32 (one fp16 rounding point of the ledger)."""
    return t.to(torch.float16).to(torch.float32)


def quantize_num_faces(n: int) -> int:
    """core/utils.py:109-125."""
    if n <= 0:
        return 0
    if n <= 1000:
        return 1
    if n <= 2000:
        return 2
    if n <= 4000:
        return 3
    if n <= 8000:
        return 4
    return 5


class ConstraintFSM:
    """core/models.py:245-271 (meto LR / LR_ABSCO grammar).  Tokens: 0 PAD 1 BOS 2 EOS 3 L 4 R 5 BOM 6.. coords."""

    def __init__(self, vocab: int, eos: int = 2):
        self.vocab, self.eos, self.counter = vocab, eos, 0

    def allowed(self, generated: List[int]) -> List[int]:
        if len(generated) == 0:
            return [5]
        last = generated[-1]
        if last == 5:
            self.counter = 9
        elif last in (3, 4):
            self.counter = 3
        elif last >= 6:
            self.counter -= 1
        if self.counter > 0:
            return list(range(6, self.vocab))
        return [3, 4, 5, self.eos]


class Oracle:
    def __init__(self, opt, state_dict: Dict[str, torch.Tensor], mode: str = 'ledger', threads: Optional[int] = None):
        assert mode in ('fp32', 'ledger')
        assert opt.cond_mode in ('point', 'point_latent')
        if threads:
            torch.set_num_threads(threads)
        self.opt, self.mode = opt, mode
        self.ledger = mode == 'ledger'
        # In ledger mode the model has been .half()'ed: every parameter/buffer holds an fp16 value.
        self.w = {k: (_r16(v.float()) if self.ledger else v.float().clone()) for k, v in state_dict.items()}
        self.C = opt.hidden_dim
        self.H = opt.num_heads
        self.D = self.C // self.H
        self.layers = opt.num_layers
        self.V = self.w['mesh_decoder.lm_head.weight'].shape[0]
        self.kv = None
        self.L = 0

    # ---- primitive ops with the ledger's rounding points --------------------------------------
    def r(self, t):
        return _r16(t) if self.ledger else t

    def linear(self, x, name, bias=True):
        """nn.Linear under autocast: inputs cast to fp16, fp32 accumulate, +bias, one rounding to fp16."""
        y = self.r(x) @ self.w[name + '.weight'].t()
        if bias:
            y = y + self.w[name + '.bias']
        return self.r(y)

    def layer_norm(self, x, name):
        """autocast runs layer_norm in fp32 and returns fp32."""
        return F.layer_norm(x, (x.shape[-1],), self.w[name + '.weight'], self.w[name + '.bias'], 1e-5)

    def attention(self, q, k, v, causal):
        """q [N,H,D], k/v [M,H,D] -> [N,H,D]; attention.py:47-62 (fp32 softmax; scale 1/sqrt(D))."""
        N, M = q.shape[0], k.shape[0]
        if causal and N > 1:
            assert N == M
            return self.attention_rows(q, k, v, 0)
        s = torch.einsum('nhd,mhd->hnm', q, k) / (self.dim_sqrt(q.shape[-1]))
        p = torch.softmax(s, dim=-1)
        return self.r(torch.einsum('hnm,mhd->nhd', p, v))

    def attention_rows(self, q, k, v, q0, chunk=512):
        """Causal attention of query rows q0 .. q0+N-1 over keys 0 .. (own position), in row chunks so that the score matrix of a long
        sequence fits in memory (same arithmetic per row as the one-shot form: fp32 scores + triu(-inf) mask, softmax, P @ V)."""
        N = q.shape[0]
        scale = self.dim_sqrt(q.shape[-1])
        outs = []
        for c0 in range(0, N, chunk):
            c1 = min(N, c0 + chunk)
            m = q0 + c1                                            # keys visible to the last row of the chunk
            s = torch.einsum('nhd,mhd->hnm', q[c0:c1], k[:m]) / scale
            rows = torch.arange(q0 + c0, q0 + c1)[:, None]
            s = s.masked_fill(torch.arange(m)[None, :] > rows, float('-inf'))
            p = torch.softmax(s, dim=-1)
            outs.append(torch.einsum('hnm,mhd->nhd', p, v[:m]))
        return self.r(torch.cat(outs, dim=0))

    @staticmethod
    def dim_sqrt(d):
        return d ** 0.5

    # ---- point encoder (prefill part 1) ---------------------------------------------------------
    def encode_points(self, pc: torch.Tensor) -> torch.Tensor:
        """pc [B,N,3] fp32 -> latent mean [B, latent_size, latent_dim]."""
        w, pe = self.w, 'point_encoder.'
        outs = []
        for b in range(pc.shape[0]):
            x = pc[b].float()
            # PointEmbed.embed: einsum under autocast -> fp16 operands, fp16 result; sin/cos in fp16
            proj = self.r(self.r(x) @ w[pe + 'point_embed.basis'])
            emb = torch.cat([self.r(torch.sin(proj)), self.r(torch.cos(proj)), x], dim=1)  # cat promotes to fp32
            kvx = self.layer_norm(self.linear(emb, pe + 'point_embed.mlp'), pe + 'ln')       # [N,E] fp32
            q0 = w[pe + 'query_embed'][0]                                                    # [Lq,E]
            E = q0.shape[-1]
            Hh = self.opt.point_num_heads
            Dh = E // Hh
            ca = pe + 'cross_att.'
            q = self.linear(self.layer_norm(q0, ca + 'ln1'), ca + 'att.q_proj').view(-1, Hh, Dh)
            k = self.linear(kvx, ca + 'att.k_proj').view(-1, Hh, Dh)
            v = self.linear(kvx, ca + 'att.v_proj').view(-1, Hh, Dh)
            a = self.attention(q, k, v, causal=False).reshape(-1, E)
            x1 = self.r(q0 + self.linear(a, ca + 'att.out_proj'))                             # fp16 + fp16
            h = self.linear(self.layer_norm(x1, ca + 'ln2'), ca + 'mlp.net.0')
            a_, g_ = h.chunk(2, dim=-1)
            h = self.r(a_ * self.r(F.gelu(g_)))                                                # GEGLU, erf gelu
            x2 = self.r(x1 + self.linear(h, ca + 'mlp.net.2'))
            outs.append(self.linear(x2, pe + 'linear'))
        return torch.stack(outs)

    def encode_cond(self, conds: torch.Tensor, num_faces: int) -> torch.Tensor:
        """-> cond_embeds [B, P, C] (fp32: LayerNorm output; the num-face row is an fp16 embedding)."""
        lat = self.encode_points(conds) if self.opt.cond_mode == 'point' else self.r(conds.float())
        ce = self.layer_norm(self.linear(lat, 'proj_cond'), 'norm_cond')
        if self.opt.use_num_face_cond:
            row = self.w['embed_num_face.weight'][quantize_num_faces(int(num_faces))]
            ce = torch.cat([ce, row.expand(ce.shape[0], 1, -1)], dim=1)
        return ce

    # ---- decoder ----------------------------------------------------------------------------------
    def reset_cache(self, max_len: int):
        self.kv = torch.zeros(self.layers, 2, max_len, self.H, self.D)
        self.L = 0

    def decoder_rows(self, hidden: torch.Tensor, first_residual_fp16: bool, all_logits: bool = False, replay: bool = False):
        """Run ``hidden`` [N,C] (already embeds + positions) through all layers, appending to the KV cache.

        first_residual_fp16: decode steps enter layer 0 with an fp16 hidden state, so the first
        residual add is an fp16 + fp16 add (Appendix B); the prefill enters in fp32."""
        md = 'mesh_decoder.model.layers.'
        N = hidden.shape[0]
        L0 = self.L
        h = hidden
        for i in range(self.layers):
            lp = md + '%d.' % i
            q = self.linear(h, lp + 'self_attn.q_proj').view(N, self.H, self.D)
            k = self.linear(h, lp + 'self_attn.k_proj').view(N, self.H, self.D)
            v = self.linear(h, lp + 'self_attn.v_proj').view(N, self.H, self.D)
            self.kv[i, 0, L0:L0 + N] = k
            self.kv[i, 1, L0:L0 + N] = v
            if N == 1:
                a = self.attention(q, self.kv[i, 0, :L0 + 1], self.kv[i, 1, :L0 + 1], causal=True)
            elif L0 == 0:
                a = self.attention(q, k, v, causal=True)
            else:
                # test-only replay (replay_steps): N consecutive decode steps evaluated together; row j attends to keys 0 .. L0+j.  The
                # reference itself never does this (attention.py:40-41 asserts N == 1 or N == M); per row it is the same arithmetic as N
                # calls of step().
                assert replay, 'multi-row pass must start from an empty cache (attention.py:40-41)'
                a = self.attention_rows(q, self.kv[i, 0, :L0 + N], self.kv[i, 1, :L0 + N], L0)
            o = self.linear(a.reshape(N, self.C), lp + 'self_attn.out_proj')
            h = h + o
            if first_residual_fp16 and i == 0:
                h = self.r(h)
            h = self.layer_norm(h, lp + 'self_attn_layer_norm')
            f = torch.relu(self.linear(h, lp + 'fc1'))
            h = self.layer_norm(h + self.linear(f, lp + 'fc2'), lp + 'final_layer_norm')
        self.L = L0 + N
        rows = h if all_logits else h[-1:]
        pre = self.r(rows) @ self.w['mesh_decoder.lm_head.weight'].t()   # fp32 value before the fp16 store
        return pre

    def prefill(self, cond_embeds: torch.Tensor, prompt_ids: List[int], all_logits=False):
        """cond_embeds [P,C] fp32, prompt ids (BOS [+ resume]) -> pre-rounding logits of the last row."""
        w = self.w
        tok = w['mesh_decoder.model.embd.weight'][torch.tensor(prompt_ids, dtype=torch.long)]
        x = torch.cat([cond_embeds, tok], dim=0)
        pos = w['mesh_decoder.model.embed_positions.weight'][:x.shape[0]]
        return self.decoder_rows(x + pos, first_residual_fp16=False, all_logits=all_logits)

    def step(self, token: int):
        w = self.w
        x = w['mesh_decoder.model.embd.weight'][token] + w['mesh_decoder.model.embed_positions.weight'][self.L]
        return self.decoder_rows(self.r(x)[None], first_residual_fp16=True)

    def replay_steps(self, tokens: List[int], chunk: int = 2048):
        """Teacher-forced replay of len(tokens) consecutive DECODE steps (fed tokens known in advance), evaluated chunk-wise instead of one
        by one: identical per-row arithmetic to step() (fp16 entry, fp16 first residual, causal attention over the cache), used by the
        long-context GPU tests to check every position of a 16k-token stream in a minute instead of hours.  -> logits_pre [n, V]."""
        w = self.w
        outs = []
        for c0 in range(0, len(tokens), chunk):
            ids = torch.tensor(tokens[c0:c0 + chunk], dtype=torch.long)
            x = w['mesh_decoder.model.embd.weight'][ids] + w['mesh_decoder.model.embed_positions.weight'][self.L:self.L + len(ids)]
            outs.append(self.decoder_rows(self.r(x), first_residual_fp16=True, all_logits=True, replay=True))
        return torch.cat(outs, dim=0)

    # ---- HF-4.46.2 _sample restated ---------------------------------------------------------------
    def generate(self, conds: torch.Tensor, num_faces: int = 1000, max_new_tokens: int = 64,
                 generate_mode: str = 'greedy', resume_ids: Optional[List[int]] = None,
                 use_tokenizer_fsm: bool = True, forced_tokens: Optional[List[int]] = None,
                 generator: Optional[torch.Generator] = None):
        """Returns dict(tokens, logits_pre [T,V] fp32 before the fp16 rounding, logits [T,V] as HF sees them).

        forced_tokens: teacher forcing — feed these ids instead of the oracle's own choice (the oracle's
        choice is still recorded in ``tokens``), used to check another implementation step by step."""
        assert conds.shape[0] == 1
        ce = self.encode_cond(conds, num_faces)[0]
        prompt = [self.opt.bos_token_id] + list(resume_ids or [])
        self.reset_cache(ce.shape[0] + len(prompt) + max_new_tokens + 1)
        fsm = ConstraintFSM(self.V, self.opt.eos_token_id)
        pre = self.prefill(ce, prompt)
        generated: List[int] = []   # what the FSM / next step sees (forced stream when teacher forcing)
        chosen: List[int] = []
        pres, outs = [], []
        for t in range(max_new_tokens):
            logits = self.r(pre[0]).float()                       # fp16 lm_head output, then .float()
            if use_tokenizer_fsm:
                allowed = fsm.allowed(generated)
            else:  # core/models.py:237-242
                allowed = list(range(3, self.V)) + ([self.opt.eos_token_id] if len(generated) % 9 == 1 else [])
            mask = torch.full_like(logits, -math.inf)
            mask[torch.tensor(allowed)] = 0
            scores = logits + mask
            if generate_mode == 'sample':
                kth = torch.topk(scores, min(10, scores.numel()))[0][-1]
                scores = scores.masked_fill(scores < kth, -math.inf)
                probs = torch.softmax(scores, dim=-1)
                nxt = int(torch.multinomial(probs, 1, generator=generator))
            else:
                nxt = int(torch.argmax(scores))
            pres.append(pre[0].clone())
            outs.append(scores)
            chosen.append(nxt)
            fed = nxt if forced_tokens is None else int(forced_tokens[t])
            generated.append(fed)
            if fed == self.opt.eos_token_id or t == max_new_tokens - 1:
                break
            pre = self.step(fed)
        return dict(tokens=np.asarray(chosen, dtype=np.int64), logits_pre=torch.stack(pres),
                    scores=torch.stack(outs), cond_embeds=ce)

    # ---- teacher-forced forward (config 4) ------------------------------------------------------------
    def forward_tf(self, conds, tokens, labels, num_faces):
        """core/models.py:147-202 in eval mode (no num-face dropout), dense causal (no padding).

        conds [B,N,3]; tokens [B,1+M+1] long; labels [B,P+1+M+1] long (-100 ignored); num_faces [B].
        Returns dict(loss, loss_ce, loss_kl, logits_pre [B,T,V])."""
        B = tokens.shape[0]
        w = self.w
        logits, kl = [], 0.0
        lat_all = self.encode_points(conds) if self.opt.cond_mode == 'point' else self.r(conds.float())
        for b in range(B):
            ce = self.layer_norm(self.linear(lat_all[b], 'proj_cond'), 'norm_cond')
            if self.opt.use_num_face_cond:
                ce = torch.cat([ce, w['embed_num_face.weight'][quantize_num_faces(int(num_faces[b]))][None]], 0)
            self.reset_cache(ce.shape[0] + tokens.shape[1])
            logits.append(self.prefill(ce, tokens[b].tolist(), all_logits=True))
        logits = torch.stack(logits)
        lg = self.r(logits)                                      # fp16 lm_head output
        # cross_entropy is on autocast's fp32 list
        loss_ce = F.cross_entropy(lg[:, :-1].reshape(-1, self.V).float(), labels[:, 1:].reshape(-1), ignore_index=-100)
        out = dict(loss_ce=loss_ce, logits_pre=logits)
        loss = loss_ce
        if self.opt.cond_mode == 'point':
            kl = 0.5 * torch.sum(lat_all.float() ** 2)           # point.py:33-35 (sum over the whole batch)
            out['loss_kl'] = kl
            loss = loss + self.opt.kl_weight * kl
        out['loss'] = loss
        return out
