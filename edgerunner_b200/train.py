"""The reference's training step (``main.py:160-185``) as one native loop over flat buffers:

    out = model(data); accelerator.backward(out['loss'])          -> Engine.train_step (er_train_step: training forward + backward)
    DDP gradient all-reduce (acc_configs/gpu8.yaml)                -> FlatGradAllReduce: the gradients are exported straight into ONE flat fp32
                                                                      buffer, all-reduced over NCCL in reverse-order slices, divided by world
    accelerator.clip_grad_norm_(..., opt.gradient_clip)            -> er_grad_norm_clip (coefficient applied inside the optimizer kernel)
    optimizer.step() [AdamW 0.9/0.95, wd 0.01]; scheduler.step()    -> FlatAdamW (er_adamw_step) + cosine_lr_lambda; the fp16 copy it writes is
                                                                      re-uploaded into the engine's weight arrays for the next forward

``core.models.LMM`` in ``train()`` mode offers the same backward through torch autograd (``loss.backward()`` + any torch optimizer); this
class is the B200-first form: no per-parameter tensors, no autograd graph, master weights / moments / gradients in four flat buffers.
Trainable: decoder, lm_head, embeddings, proj_cond, norm_cond, embed_num_face, and the point encoder unless ``opt.freeze_encoder``.
CUDA only — no CPU fallback.
"""

from __future__ import annotations

import torch

from .dist import FlatGradAllReduce, world
from .optim import FlatAdamW, cosine_lr_lambda


class FlatTrainer:
    def __init__(self, model, total_steps: int, max_batch: int, max_tokens: int, lr: float = None, warmup_ratio: float = None,
                 gradient_clip: float = None, n_slices: int = 4):
        opt = model.opt
        self.model, self.opt_cfg = model, opt
        self.engine = model.get_engine(max_new_tokens=64, max_tf_rows=max_batch * (opt.num_cond_tokens + max_tokens))
        e = self.engine
        self.train_encoder = opt.cond_mode == 'point' and not opt.freeze_encoder
        self.entries = []                                     # (name, offset, numel, shape) in registration order
        off = 0
        for name, p in model.named_parameters():
            if e.grad_has(name, self.train_encoder):
                self.entries.append((name, off, p.numel(), tuple(p.shape)))
                off += (p.numel() + 3) // 4 * 4               # 16-byte aligned slices
        self.numel = off
        dev = e.device
        self.param = torch.zeros(off, dtype=torch.float32, device=dev)
        params = dict(model.named_parameters())
        for name, o, n, shp in self.entries:
            self.param[o:o + n].copy_(params[name].detach().reshape(-1).float())
        self.param16 = self.param.half()
        self.reducer = FlatGradAllReduce(off, dev, n_slices=n_slices)
        self.grad = self.reducer.buf
        self.optim = FlatAdamW(self.param, lr=opt.lr if lr is None else lr, betas=(0.9, 0.95), weight_decay=0.01, param16=self.param16)
        self.total_steps = int(total_steps)
        self.warmup_ratio = opt.warmup_ratio if warmup_ratio is None else warmup_ratio
        self.gradient_clip = opt.gradient_clip if gradient_clip is None else gradient_clip
        self.step_count = 0
        self.dropout_p = float(model.config.dropout)
        self._rng = torch.Generator().manual_seed(1234 + world()[0])
        # overflow back-off of the fp16 activation gradients (what accelerate's GradScaler does for the reference's fp16 mode): a step whose global
        # gradient norm is not finite is skipped and the loss scale halved; it creeps back up after 200 good steps
        self.scale_mult, self.good_steps, self.skipped_steps = 1.0, 0, 0

    def step(self, data, loss_scale=None, check_finite=True):
        """One optimizer step on this rank's batch -> dict(loss, loss_ce, loss_kl (device scalars), grad_norm, lr, skipped).  check_finite: read the
        global gradient norm back (one host sync) and skip the update when it is inf / nan (fp16 overflow), halving the loss scale."""
        e, opt = self.engine, self.opt_cfg
        num_faces = data['num_faces'].clone()
        if opt.use_num_face_cond and opt.nof_dropout_ratio > 0:                      # models.py:161-164
            drop = torch.rand((num_faces.shape[0],), generator=self._rng) < opt.nof_dropout_ratio
            num_faces[drop.to(num_faces.device)] = -1
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=self._rng).item())
        losses, _ = e.train_step(data['conds'], data['tokens'], data['labels'], num_faces.tolist(), opt.kl_weight, masks=data.get('masks'),
                                 dropout_p=self.dropout_p, seed=seed, loss_scale=loss_scale, train_encoder=self.train_encoder,
                                 loss_scale_mult=self.scale_mult)
        for name, o, n, shp in self.entries:
            e.grad(name, out=self.grad[o:o + n])
        self.reducer.launch().wait()
        lr_scale = cosine_lr_lambda(self.step_count, self.total_steps, warmup_ratio=self.warmup_ratio)
        if check_finite:
            norm, _ = self.optim.grad_norm(self.grad, self.gradient_clip if self.gradient_clip else 0.0)
            if not bool(torch.isfinite(norm)):           # every rank sees the same all-reduced gradient, hence the same decision
                self.scale_mult *= 0.5
                self.good_steps = 0
                self.skipped_steps += 1
                return {'loss': losses[0], 'loss_ce': losses[1], 'loss_kl': losses[2], 'grad_norm': norm.clone(), 'lr': self.optim.lr * lr_scale, 'skipped': True}
            self.good_steps += 1
            if self.good_steps >= 200 and self.scale_mult < 1.0:
                self.scale_mult, self.good_steps = self.scale_mult * 2.0, 0
        norm = self.optim.step(self.grad, max_norm=self.gradient_clip, lr_scale=lr_scale)
        self.step_count += 1
        # the forward kernels read the engine's fp16 weight arrays: refresh them from the optimizer's fp16 copy
        e.load_state_dict({name: self.param16[o:o + n].view(shp) for name, o, n, shp in self.entries}, persistent=True)
        return {'loss': losses[0], 'loss_ce': losses[1], 'loss_kl': losses[2], 'grad_norm': norm, 'lr': self.optim.lr * lr_scale, 'skipped': False}

    def sync_to_model(self):
        """copy the master weights back into the module's parameters (checkpointing: accelerator.save_state / safetensors of main.py)"""
        params = dict(self.model.named_parameters())
        with torch.no_grad():
            for name, o, n, shp in self.entries:
                params[name].copy_(self.param[o:o + n].view(shp).to(params[name].dtype))
