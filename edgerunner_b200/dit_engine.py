"""Python handle of the native DiT denoiser engine (include/edgerunner_b200.h, `er_dit_*`; csrc/dit.cu).

Owns the fp16 weights of ``MDiT``'s ``dit.*`` / ``proj_cond.*`` / ``norm_cond.*`` tensors on the device and runs
``DiT.forward`` (core/transformer/dit.py:165-196 of the reference), ``MDiT.get_cond``'s adaptor (core/models_dit.py:116) and the whole
guided DDIM loop of ``MDiT.run`` (core/models_dit.py:209-227).  No CPU fallback: everything here needs a CUDA device.
"""

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _lib

PRED = {'epsilon': 0, 'v_prediction': 1}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class DiTEngine:
    def __init__(self, device: torch.device, hidden_dim=1024, num_heads=16, num_layers=24, latent_size=2048, latent_dim=64,
                 cond_tokens=257, cond_dim=1280):
        if device.type != 'cuda':
            raise RuntimeError('DiTEngine needs a CUDA device (there is no CPU fallback)')
        self.lib = _lib.load()
        self.device = device
        self.cfg = _lib.ErDitConfig(device.index or 0, hidden_dim, num_heads, num_layers, latent_size, latent_dim, cond_tokens, cond_dim)
        self.h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.er_dit_create(C.byref(self.cfg), C.byref(self.h)))

    def __del__(self):
        h, self.h = getattr(self, 'h', None), None
        if h:
            self.lib.er_dit_destroy(h)

    # names the engine holds; everything else in an MDiT state dict (image_encoder.*, point_encoder.*) is not the engine's business
    @staticmethod
    def wants(name: str) -> bool:
        return name.startswith(('dit.', 'proj_cond.', 'norm_cond.'))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        with torch.cuda.device(self.device):
            for name, t in sd.items():
                if not self.wants(name):
                    continue
                t = t.detach().to(self.device)
                if t.dtype not in (torch.float16, torch.float32):
                    t = t.float()
                t = t.contiguous()
                _lib.check(self.lib.er_dit_load_weight(self.h, name.encode(), t.data_ptr(), 0 if t.dtype == torch.float16 else 1, t.numel(), _stream()))
            _lib.check(self.lib.er_dit_finalize_weights(self.h, _stream()))

    def cond(self, clip_hidden: torch.Tensor) -> torch.Tensor:
        """norm_cond(proj_cond(h)): [B, M, cond_dim] (any float dtype) -> [B, M, C] fp32."""
        h = clip_hidden.to(self.device, torch.float16).contiguous()
        B = h.shape[0]
        assert h.shape[1:] == (self.cfg.cond_tokens, self.cfg.cond_dim), h.shape
        out = torch.empty(B, self.cfg.cond_tokens, self.cfg.hidden_dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_dit_cond(self.h, h.data_ptr(), B, out.data_ptr(), _stream()))
        return out

    def forward(self, x: torch.Tensor, c: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """DiT.forward: x [B, N, Dl], c [B, M, C], t [B] -> fp16 [B, N, Dl]."""
        B = x.shape[0]
        x = x.to(self.device, torch.float32).contiguous()
        c = c.to(self.device, torch.float32).contiguous()
        t = t.to(self.device, torch.float32).contiguous()
        assert x.shape[1:] == (self.cfg.latent_size, self.cfg.latent_dim) and c.shape == (B, self.cfg.cond_tokens, self.cfg.hidden_dim) and t.shape == (B,)
        out = torch.empty(B, self.cfg.latent_size, self.cfg.latent_dim, dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_dit_forward(self.h, x.data_ptr(), c.data_ptr(), t.data_ptr(), B, out.data_ptr(), _stream()))
        return out

    def run(self, cond: torch.Tensor, latents: torch.Tensor, timesteps: np.ndarray, coef: np.ndarray, guidance_scale=7.5, guided=True,
            prediction_type='v_prediction') -> torch.Tensor:
        """The sampling loop: latents fp32 [R, N, Dl] are updated IN PLACE and returned; cond fp32 [R, M, C]."""
        assert latents.dtype == torch.float32 and latents.is_contiguous() and latents.device == self.device
        cond = cond.to(self.device, torch.float32).contiguous()
        R = latents.shape[0]
        ts = np.ascontiguousarray(timesteps, dtype=np.float32)
        cf = np.ascontiguousarray(coef, dtype=np.float32).reshape(-1, 4)
        assert len(ts) == len(cf) and cond.shape[0] == R
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_dit_run(self.h, cond.data_ptr(), latents.data_ptr(), R, len(ts), _fp(ts), _fp(cf), float(guidance_scale),
                                           1 if guided else 0, PRED[prediction_type], _stream()))
        return latents

    def run_host(self, cond: np.ndarray, latents: np.ndarray, timesteps: np.ndarray, coef: np.ndarray, guidance_scale=7.5, guided=True,
                 prediction_type='v_prediction') -> np.ndarray:
        """Host buffers in, host buffers out, synchronous (bench.py e2e): latents fp32 [R, N, Dl] updated in place."""
        assert cond.dtype == np.float32 and latents.dtype == np.float32 and cond.flags.c_contiguous and latents.flags.c_contiguous
        ts = np.ascontiguousarray(timesteps, dtype=np.float32)
        cf = np.ascontiguousarray(coef, dtype=np.float32).reshape(-1, 4)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_dit_run_host(self.h, _fp(cond), _fp(latents), latents.shape[0], len(ts), _fp(ts), _fp(cf), float(guidance_scale),
                                                1 if guided else 0, PRED[prediction_type]))
        return latents

    def kernel_launches(self) -> int:
        return int(self.lib.er_dit_kernel_launches(self.h))

    def flops_per_forward(self, batch: int) -> float:
        return float(self.lib.er_dit_flops_per_forward(self.h, batch))

    def debug_set(self, key: str, value: int):
        _lib.check(self.lib.er_dit_debug_set(self.h, key.encode(), int(value)))
