"""``LMM`` — the large mesh model facade, drop-in for ``/root/reference/core/models.py:32-319``.

Same constructor, attributes (``opt``, ``vocab_size``, ``mesh_decoder.model.embd``, ``point_encoder`` ...), checkpoint
key schema and method signatures as the reference, so ``infer.py`` runs unmodified:

    model = LMM(opt); model.load_state_dict(ckpt, strict=False); model = model.half().eval().to('cuda')
    meshes, tokens = model.generate(cond, num_faces=..., max_new_tokens=..., tokenizer=tokenizer, clean=True)

What differs is where the work happens: ``encode_cond`` / ``generate`` / ``forward`` hand raw device pointers to the
sm_100a CUDA library (``edgerunner_b200``).  The whole auto-regressive loop — embedding, 24 decoder layers over an
in-place KV cache, lm_head, grammar-constrained argmax / top-k sampling, EOS test — is one persistent kernel; the host
is not involved per token.  There is no CPU or eager-PyTorch fallback: without CUDA these methods raise.

Numerics follow ``infer.py``'s ``model.half()`` + ``torch.autocast('cuda', fp16)`` regardless of
the dtype the parameters are stored in: weights are rounded to fp16 when they are uploaded to the engine.
"""

import numpy as np
import torch
import torch.nn as nn

from core.options import Options
from core.provider import save_mesh
from core.utils import quantize_num_faces
from core.transformer.point import DummyLatent


class LMM(nn.Module):
    def __init__(self, opt: Options):
        super().__init__()
        self.opt = opt

        # ---- conditioner (reference :38-73) ---------------------------------------------------------------------
        if opt.cond_mode == 'image':
            raise NotImplementedError("cond_mode='image' (CLIP-conditioned LMM) is outside the B200 decode path; "
                                      "use 'point' or 'point_latent' (SURVEY.md §8)")
        if opt.cond_mode == 'point':
            if opt.point_encoder_mode != 'embed':
                raise NotImplementedError("point_encoder_mode='downsample' needs torch_cluster FPS; no preset uses it")
            from core.transformer.point import PointEncoderEmbed
            self.point_encoder = PointEncoderEmbed(hidden_dim=opt.point_hidden_dim, num_heads=opt.point_num_heads,
                                                   latent_size=opt.point_latent_size, latent_dim=opt.point_latent_dim,
                                                   gradient_checkpointing=opt.checkpointing)
        if opt.cond_mode in ('point', 'point_latent'):
            self.proj_cond = nn.Linear(opt.point_latent_dim, opt.hidden_dim)
            self.norm_cond = nn.LayerNorm(opt.hidden_dim)
        if opt.use_num_face_cond:
            self.embed_num_face = nn.Embedding(10, opt.hidden_dim)

        # ---- mesh decoder (reference :75-99) -------------------------------------------------------------------------
        if opt.use_meto:
            self.vocab_size = (2 * opt.discrete_bins if opt.meto_backend == 'LR' else opt.discrete_bins) + 3 + 3
        else:
            self.vocab_size = opt.discrete_bins + 3
        from core.transformer.modeling_opt import ShapeOPTConfig, ShapeOPT
        self.config = ShapeOPTConfig(
            vocab_size=self.vocab_size, hidden_dim=opt.hidden_dim,
            intermediate_dim=opt.hidden_dim * 4 if opt.intermediate_dim is None else opt.intermediate_dim,
            num_hidden_layers=opt.num_layers, num_attention_heads=opt.num_heads,
            max_position_embeddings=opt.max_seq_length + opt.num_cond_tokens + 10, num_cond_tokens=opt.num_cond_tokens)
        self.mesh_decoder = ShapeOPT(self.config)
        if opt.checkpointing:
            self.mesh_decoder.model.gradient_checkpointing_enable()

        # engine state (not part of the module state)
        self._engine = None
        self._engine_key = None

    # ---- engine management ---------------------------------------------------------------------------------------------
    def _weights_fingerprint(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def get_engine(self, max_new_tokens=None, max_tf_rows=0):
        """Create (or refresh after a weight / capacity change) the CUDA engine holding the packed fp16 weights."""
        from edgerunner_b200.engine import Engine
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError('LMM: parameters are not on a CUDA device; the B200 path has no CPU fallback '
                               '(call model.to("cuda") first)')
        cap = int(max_new_tokens if max_new_tokens is not None else self.opt.max_seq_length)
        key = (p0.device, self._weights_fingerprint())
        e = self._engine
        if e is None or e.max_new_tokens < cap or e.cfg.max_tf_rows < max_tf_rows or e.device != p0.device:
            self._engine = None
            e = Engine(self.opt, p0.device, max_new_tokens=cap, max_points=max(self.opt.point_num, 8192), max_tf_rows=max_tf_rows)
            self._engine_key = None
        if self._engine_key != key:
            # infer_dit.py builds LMM with cond_mode='point' and then flips the shared Options to 'point_latent' (infer_dit.py:38-55): the engine
            # is then created without a point encoder and the module's point_encoder.* tensors are not its business
            e.load_state_dict({k: v for k, v in self.state_dict().items() if e.cfg.has_point_encoder or not k.startswith('point_encoder.')})
            self._engine_key = key
        self._engine = e
        return e

    # ---- reference API ------------------------------------------------------------------------------------------------------
    def encode_cond(self, conds, num_faces):
        """-> {'cond_embeds': [B, P, C] fp32 (LayerNorm output), 'posterior': DummyLatent}  (reference :101-144)."""
        if self.opt.cond_mode == 'none':
            raise NotImplementedError("cond_mode='none' is outside the B200 decode path")
        e = self.get_engine(max_new_tokens=getattr(self._engine, 'max_new_tokens', 64) if self._engine else 64)
        results, embeds, lats = {}, [], []
        for b in range(conds.shape[0]):
            nf = int(num_faces[b]) if torch.is_tensor(num_faces) or isinstance(num_faces, (list, tuple, np.ndarray)) else int(num_faces)
            emb, lat = e.encode_cond(conds[b], nf, want_embeds=True, want_latents=self.opt.cond_mode == 'point')
            embeds.append(emb)
            lats.append(lat)
        if self.opt.cond_mode == 'point':
            results['posterior'] = DummyLatent(torch.stack(lats))
        results['cond_embeds'] = torch.stack(embeds)
        return results

    def forward(self, data, step_ratio=1):
        """Teacher-forced forward -> {'loss', 'loss_ce', 'loss_kl', 'logits'}  (reference :147-202).

        ``model.eval()``: inference-mode numerics (no dropout, no num-face dropout), no autograd graph.  ``model.train()``: the training step of
        main.py:168-172 — see ``_forward_train``."""
        if self.training:
            return self._forward_train(data)
        masks = data.get('masks')                      # [B, P+T] bool; right-padded batches (collate_fn) take the masked path
        tokens, labels = data['tokens'], data['labels']
        B, T = tokens.shape
        e = self.get_engine(max_new_tokens=getattr(self._engine, 'max_new_tokens', 64) if self._engine else 64,
                            max_tf_rows=B * (self.opt.num_cond_tokens + T))
        losses, logits, sums = e.forward_tf(data['conds'], tokens, labels, data['num_faces'].tolist(), self.opt.kl_weight, want_logits=True,
                                            masks=masks, want_sums=True)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # data parallel over the batch (BASELINE configs[3]): ONE all-reduce of {sum of token CEs, token count, KL}; every rank gets
            # the loss of the concatenated batch
            from edgerunner_b200.dist import dp_reduce_losses
            loss, loss_ce, loss_kl = dp_reduce_losses(sums[0], sums[1], sums[2], self.opt.kl_weight if self.opt.cond_mode == 'point' else 0.0)
            losses = torch.stack([loss, loss_ce, loss_kl])
        out = {'loss': losses[0], 'loss_ce': losses[1], 'logits': logits}
        if self.opt.cond_mode == 'point':
            out['loss_kl'] = losses[2]
        return out

    def _forward_train(self, data):
        """Training-mode forward (reference :147-202 with self.training): num-face dropout (:161-164), F.dropout(config.dropout) on both branches of
        every decoder layer, and a loss that carries a graph: ``out['loss'].backward()`` fills ``.grad`` of the decoder, lm_head, embeddings,
        proj_cond / norm_cond / embed_num_face (and, unless frozen, point-encoder) parameters.  The forward AND the backward run inside one library call (er_train_step: checkpointed
        layers, tcgen05 dgrad / wgrad GEMMs, flash-attention backward); the autograd node only hands the stored gradients out.  As with
        ``opt.freeze_encoder = True`` (the Options default) the point encoder runs without a graph and loss_kl carries no gradient; with
        ``freeze_encoder = False`` (the ArAE preset; the only value the reference accepts in 'point' mode, reference :54) the library call also walks back
        through the point encoder and the KL term.  The dropout mask is a counter-based function of a seed drawn from torch's
        global generator (reproducible per seed; not torch's Philox stream).  'logits' is None in this mode (main.py never reads it while training)."""
        tokens, labels = data['tokens'], data['labels']
        B, T = tokens.shape
        num_faces = data['num_faces']
        if self.opt.use_num_face_cond:                                            # random num_faces dropout (reference :161-164, in place like it)
            unprog_mask = torch.rand((B,), device=num_faces.device) < self.opt.nof_dropout_ratio
            num_faces[unprog_mask] = -1
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        named = [(n, p) for n, p in self.named_parameters() if p.requires_grad]
        loss, loss_ce, loss_kl = _TrainStep.apply(self, data, seed, *[p for _, p in named])
        out = {'loss': loss, 'loss_ce': loss_ce, 'logits': None}
        if self.opt.cond_mode == 'point':
            out['loss_kl'] = loss_kl
        return out

    @torch.no_grad()
    def generate(self, conds, num_faces=1000, resume_ids=None, tokenizer=None, max_new_tokens=None, clean=True):
        """Reference :204-319.  Returns (list of meshes, list of np.int64 token arrays incl. EOS, +3 offset)."""
        B = conds.shape[0]
        assert B == 1, 'Batch size must be 1 for generation.'
        use_fsm = tokenizer is not None
        if tokenizer is not None and self.opt.meto_backend not in ('LR', 'LR_ABSCO'):
            print('[WARN] prefix_allowed_tokens_fn is not defined for meto backend:', self.opt.meto_backend)
            use_fsm = False                                   # reference :273-275: the constraint is disabled, not kept
        max_new_tokens = self.opt.max_seq_length if max_new_tokens is None else max_new_tokens
        prompt = [self.opt.bos_token_id]
        if resume_ids is not None:
            prompt += [int(x) for x in resume_ids[0].detach().cpu().tolist()]
        # the position table has max_seq_length + num_cond_tokens + 10 rows (reference :84): a resumed request with the default
        # max_new_tokens = max_seq_length simply runs until the table ends (the reference would index past it)
        max_new_tokens = min(int(max_new_tokens), self.opt.max_seq_length + 10 - len(prompt))
        e = self.get_engine(max_new_tokens=max_new_tokens + max(0, len(prompt) - 64))    # the engine keeps 64 rows of slack for the prompt
        e.encode_cond(conds[0], int(num_faces))
        e.prefill(prompt)
        mode = 'greedy' if self.opt.generate_mode == 'greedy' else 'sample'
        # sample mode draws its seed from torch's global generator, like HF's multinomial does
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if mode == 'sample' else 0
        out = e.decode(max_new_tokens, mode=mode, top_k=10, seed=seed, use_fsm=use_fsm)

        tokens = out['tokens']
        if resume_ids is not None:
            tokens = np.concatenate((resume_ids[0].detach().cpu().numpy(), tokens), axis=0)
        mesh = save_mesh(tokens, self.opt, tokenizer=tokenizer, clean=clean, verbose=True)
        return [mesh], [tokens]


class _TrainStep(torch.autograd.Function):
    """loss = er_train_step(batch); d loss / d parameter = the gradients the same call left in the engine (fp32, exported on backward)."""

    @staticmethod
    def forward(ctx, model, data, seed, *params):
        tokens = data['tokens']
        B, T = tokens.shape
        e = model.get_engine(max_new_tokens=getattr(model._engine, 'max_new_tokens', 64) if model._engine else 64,
                             max_tf_rows=B * (model.opt.num_cond_tokens + T))
        losses, _ = e.train_step(data['conds'], tokens, data['labels'], data['num_faces'].tolist(), model.opt.kl_weight, masks=data.get('masks'),
                                 dropout_p=float(model.config.dropout), seed=seed, loss_scale=getattr(model, 'loss_scale', None),
                                 train_encoder=model.opt.cond_mode == 'point' and not model.opt.freeze_encoder)
        ctx.engine = e
        ctx.names = [n for n, p in model.named_parameters() if p.requires_grad]
        ctx.meta = [(p.shape, p.dtype) for p in params]
        return losses[0].clone(), losses[1].clone(), losses[2].clone()

    @staticmethod
    def backward(ctx, g_loss, g_ce, g_kl):
        e = ctx.engine
        grads = []
        for name, (shape, dtype) in zip(ctx.names, ctx.meta):
            if not e.grad_has(name):
                grads.append(None)
                continue
            g = e.grad(name, shape)
            grads.append((g * g_loss).to(dtype))
        return (None, None, None, *grads)


ArAE = LMM   # the tyro preset name of the reference (core/options.py:158) — convenience alias
