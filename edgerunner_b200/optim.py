"""Optimizer half of the reference's training step (``main.py:133-185``) over flat fp32 buffers, on the sm_100a library:

* ``FlatAdamW``       — ``torch.optim.AdamW(params, lr, weight_decay=0.01, betas=(0.9, 0.95))`` (main.py:133) + ``accelerator.clip_grad_norm_``
  (main.py:175-177): one fused streaming kernel per step (``er_adamw_step``; the clipping coefficient of ``er_grad_norm_clip`` is applied as
  the gradient is read), parameters / moments in one flat buffer each, an fp16 copy of the updated parameters for the forward kernels.
* ``cosine_lr_lambda`` — the LambdaLR function of main.py:136-141 (linear warm-up, cosine decay to ``min_ratio``).

Gradients come from ``Engine.train_step`` (er_train_step; edgerunner_b200/train.py::FlatTrainer drives the whole step); tests also feed random ones and compare with
``torch.optim.AdamW`` / ``torch.nn.utils.clip_grad_norm_``).  CUDA only — no CPU fallback.
"""

import ctypes as C
import math

import torch

from . import _lib


def cosine_lr_lambda(current_step, total_steps, warmup_ratio=0.01, num_cycles=0.5, min_ratio=0.1):
    """main.py:136-141."""
    progress = current_step / max(1, total_steps)
    if warmup_ratio > 0 and progress < warmup_ratio:
        return progress / warmup_ratio
    progress = (progress - warmup_ratio) / (1 - warmup_ratio)
    return max(min_ratio, min_ratio + (1 - min_ratio) * 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatAdamW:
    def __init__(self, param: torch.Tensor, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01, param16: torch.Tensor = None):
        """param: flat fp32 CUDA tensor updated in place; param16 (optional): fp16 tensor of the same length that receives the rounded copy."""
        if not (param.is_cuda and param.dtype == torch.float32 and param.dim() == 1 and param.is_contiguous()):
            raise RuntimeError('FlatAdamW needs a flat contiguous fp32 CUDA tensor (no CPU fallback)')
        self.lib = _lib.load()
        self.param, self.param16 = param, param16
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(param), torch.zeros_like(param)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self._scratch = torch.empty(592, dtype=torch.float64, device=param.device)
        self._norm = torch.empty(2, dtype=torch.float32, device=param.device)          # [norm, clip coefficient]

    def grad_norm(self, grad: torch.Tensor, max_norm: float):
        """-> (total norm, clip coefficient) as 0-dim device tensors (no host sync)."""
        with torch.cuda.device(self.param.device):
            _lib.check(self.lib.er_grad_norm_clip(grad.data_ptr(), grad.numel(), float(max_norm), self._scratch.data_ptr(), self._norm.data_ptr(),
                                                  self._norm.data_ptr() + 4, _stream()))
        return self._norm[0], self._norm[1]

    def step(self, grad: torch.Tensor, max_norm: float = None, lr_scale: float = 1.0):
        """One update from the flat fp32 gradient; max_norm: clip the global gradient norm first (accelerator.clip_grad_norm_)."""
        assert grad.is_cuda and grad.dtype == torch.float32 and grad.numel() == self.param.numel() and grad.is_contiguous()
        self.step_count += 1
        norm = None
        scale_ptr = None
        if max_norm is not None:
            norm, _ = self.grad_norm(grad, max_norm)
            scale_ptr = self._norm.data_ptr() + 4
        with torch.cuda.device(self.param.device):
            _lib.check(self.lib.er_adamw_step(self.param.data_ptr(), grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                              self.param16.data_ptr() if self.param16 is not None else None, self.param.numel(),
                                              float(self.lr * lr_scale), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                              float(self.weight_decay), self.step_count, scale_ptr, _stream()))
        return norm
