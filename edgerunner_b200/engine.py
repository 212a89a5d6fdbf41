"""Host-side handle on the CUDA engine (ctypes over include/edgerunner_b200.h).

PyTorch is used for device memory and the current stream only; every computation happens inside the library.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .synth import vocab_size_of


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Engine:
    def __init__(self, opt, device: torch.device, max_new_tokens: int, max_points: Optional[int] = None, max_tf_rows: int = 0,
                 debug: Optional[Dict[str, int]] = None):
        if not torch.cuda.is_available():
            raise RuntimeError('edgerunner_b200: no CUDA device; this path has no CPU fallback')
        if opt.cond_mode not in ('point', 'point_latent'):
            raise NotImplementedError(f"cond_mode '{opt.cond_mode}' is outside the B200 decode path (point / point_latent only)")
        self.lib = _lib.load()
        self.opt = opt
        self.device = torch.device(device)
        self.V = vocab_size_of(opt)
        self.P = opt.num_cond_tokens
        self.C = opt.hidden_dim
        cfg = _lib.ErConfig()
        cfg.device = self.device.index or 0
        cfg.hidden_dim, cfg.num_heads, cfg.num_layers = opt.hidden_dim, opt.num_heads, opt.num_layers
        cfg.ffn_dim = opt.hidden_dim * 4 if opt.intermediate_dim is None else opt.intermediate_dim
        cfg.vocab_size = self.V
        cfg.max_positions = opt.max_seq_length + opt.num_cond_tokens + 10
        cfg.num_cond_tokens = opt.num_cond_tokens
        cfg.use_num_face_cond = int(opt.use_num_face_cond)
        cfg.bos_token_id, cfg.eos_token_id, cfg.pad_token_id = opt.bos_token_id, opt.eos_token_id, opt.pad_token_id
        cfg.has_point_encoder = int(opt.cond_mode == 'point')
        cfg.point_hidden_dim, cfg.point_num_heads = opt.point_hidden_dim, opt.point_num_heads
        cfg.point_latent_size, cfg.point_latent_dim = opt.point_latent_size, opt.point_latent_dim
        self.max_new_tokens = int(max_new_tokens)
        cfg.max_seq_rows = min(self.P + 64 + self.max_new_tokens, cfg.max_positions)
        cfg.max_points = int(max_points or opt.point_num)
        cfg.max_tf_rows = int(max_tf_rows)
        self.cfg = cfg
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._keep = []
        for k, v in (debug or {}).items():      # experiment switches (scripts/, tests): explicit, never from the environment
            self.debug_set(k, v)

    def debug_set(self, key: str, value: int):
        _lib.check(self.lib.er_debug_set(self.h, key.encode(), int(value)))

    def __del__(self):
        h, self.h = getattr(self, 'h', None), None
        if h:
            try:
                self.lib.er_destroy(h)
            except Exception:
                pass

    # ---- weights ----------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], persistent: bool = False):
        """Upload a checkpoint in the reference key schema (fp32 or fp16 tensors on any device).  persistent=True: the tensors outlive the call
        (device-resident views of a long-lived buffer, as the training loop passes), so the stream is not synchronised after every entry."""
        with torch.cuda.device(self.device):
            for name, t in sd.items():
                t = t.detach()
                if t.dtype not in (torch.float16, torch.float32):
                    t = t.float()
                t = t.to(self.device, non_blocking=False).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
                dt = _lib.C.c_int32(0 if t.dtype == torch.float16 else 1)
                _lib.check(self.lib.er_load_weight(self.h, name.encode(), t.data_ptr(), dt, shape, max(t.dim(), 1), _stream()))
                if not persistent:
                    torch.cuda.current_stream().synchronize()   # t may be a temporary
            _lib.check(self.lib.er_finalize_weights(self.h, _stream()))

    # ---- generate pieces --------------------------------------------------------------------------------------------------
    def encode_cond(self, conds: torch.Tensor, num_faces: int, want_embeds=False, want_latents=False):
        """conds: [n,3] fp32 points ('point') or [latent_size, latent_dim] fp32 latents ('point_latent'), on the device."""
        is_latent = int(self.opt.cond_mode == 'point_latent')
        conds = conds.to(self.device, torch.float32).contiguous()
        emb = torch.empty((self.P, self.C), dtype=torch.float32, device=self.device) if want_embeds else None
        lat = torch.empty((self.opt.point_latent_size, self.opt.point_latent_dim), dtype=torch.float16, device=self.device) if want_latents else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_encode_cond(self.h, conds.data_ptr(), conds.shape[0], is_latent, int(num_faces),
                                               emb.data_ptr() if emb is not None else None,
                                               lat.data_ptr() if lat is not None else None, _stream()))
        return emb, lat

    def prefill(self, prompt_ids: Sequence[int]):
        arr = (C.c_int32 * len(prompt_ids))(*[int(x) for x in prompt_ids])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_prefill(self.h, arr, len(prompt_ids), _stream()))

    def decode(self, max_new_tokens: int, mode: str = 'greedy', top_k: int = 10, seed: int = 0, use_fsm: bool = True,
               tokens_per_launch: int = 0, want_logits: bool = False, forced: Optional[Sequence[int]] = None, sync: bool = True):
        """-> dict(tokens np.int64 [T], logits_pre torch [T,V] | None).  With sync=False returns device tensors."""
        T = int(max_new_tokens)
        ids = torch.empty(T, dtype=torch.int32, device=self.device)
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        logits = torch.empty((T, self.V), dtype=torch.float32, device=self.device) if want_logits else None
        f = torch.as_tensor(np.asarray(forced, dtype=np.int32), device=self.device) if forced is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_decode(self.h, T, 0 if mode == 'greedy' else 1, int(top_k), C.c_uint64(int(seed) & (2 ** 64 - 1)),
                                          int(bool(use_fsm)), int(tokens_per_launch), ids.data_ptr(), n.data_ptr(),
                                          logits.data_ptr() if logits is not None else None,
                                          f.data_ptr() if f is not None else None, _stream()))
        if not sync:
            self._keep = [f]
            return dict(ids=ids, n=n, logits_pre=logits)
        cnt = int(n.item())
        return dict(tokens=ids[:cnt].cpu().numpy().astype(np.int64), logits_pre=None if logits is None else logits[:cnt])

    def generate_host(self, conds_host: np.ndarray, num_faces: int, max_new_tokens: int, mode='greedy', top_k=10, seed=0,
                      use_fsm=True, resume_ids: Optional[Sequence[int]] = None) -> np.ndarray:
        """Host buffers in, host ids out (the e2e entry: H2D / D2H copies happen inside)."""
        conds_host = np.ascontiguousarray(conds_host, dtype=np.float32)
        is_latent = int(self.opt.cond_mode == 'point_latent')
        out = np.empty(max_new_tokens, dtype=np.int32)
        n = C.c_int32(0)
        res = np.asarray(resume_ids if resume_ids is not None else [], dtype=np.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_generate_host(self.h, conds_host.ctypes.data, conds_host.shape[0], is_latent, int(num_faces),
                                                 res.ctypes.data if len(res) else None, len(res), int(max_new_tokens),
                                                 0 if mode == 'greedy' else 1, int(top_k), C.c_uint64(int(seed) & (2 ** 64 - 1)),
                                                 int(bool(use_fsm)), out.ctypes.data, C.byref(n)))
        return out[:n.value].astype(np.int64)

    def forward_tf(self, conds: torch.Tensor, tokens: torch.Tensor, labels: torch.Tensor, num_faces, kl_weight: float, want_logits=False,
                   masks: Optional[torch.Tensor] = None, want_sums=False):
        """Teacher-forced forward (eval-mode LMM.forward).  conds [B,n,3] | [B,Lq,Ld]; tokens [B,T]; labels [B,P+T]; masks [B,P+T] bool
        (right-padded) or None.  -> (losses[3], logits | None[, sums[3] = ce_sum, n_tokens, kl])."""
        B, T = tokens.shape
        is_latent = int(self.opt.cond_mode == 'point_latent')
        conds = conds.to(self.device, torch.float32).contiguous()
        tok = tokens.to(self.device, torch.int32).contiguous()
        lab = labels.to(self.device, torch.int64).contiguous()
        nf = (C.c_int32 * B)(*[int(x) for x in num_faces])
        losses = torch.zeros(3, dtype=torch.float32, device=self.device)
        sums = torch.zeros(3, dtype=torch.float64, device=self.device)
        logits = torch.empty((B, self.P + T, self.V), dtype=torch.float32, device=self.device) if want_logits else None
        m8 = None
        if masks is not None:
            m = masks.to(self.device).bool()
            if m.shape != (B, self.P + T):
                raise ValueError(f'masks must be [B, P+T] = {(B, self.P + T)}, got {tuple(m.shape)}')
            if bool((m[:, 1:] & ~m[:, :-1]).any()):
                raise NotImplementedError('only right-padded attention masks are supported (collate_fn pads at the end)')
            if not bool(m.all()):
                m8 = m.to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_forward_tf2(self.h, conds.data_ptr(), conds.shape[1], is_latent, tok.data_ptr(), lab.data_ptr(),
                                               m8.data_ptr() if m8 is not None else None, nf, B, T, C.c_float(kl_weight), losses.data_ptr(),
                                               sums.data_ptr(), logits.data_ptr() if logits is not None else None, _stream()))
        self._keep = [m8]
        return (losses, logits, sums) if want_sums else (losses, logits)

    # ---- training step ------------------------------------------------------------------------------------------
    def train_step(self, conds: torch.Tensor, tokens: torch.Tensor, labels: torch.Tensor, num_faces, kl_weight: float,
                   masks: Optional[torch.Tensor] = None, dropout_p: float = 0.1, seed: int = 0, loss_scale: Optional[float] = None,
                   train_encoder: bool = False, loss_scale_mult: float = 1.0):
        """Training-mode forward + backward of LMM.forward on this batch (er_train_step): -> (losses[3] = loss, mean CE, KL; sums[3]).
        The gradients stay in the engine until ``grad(name, ...)`` exports them.  Arguments as ``forward_tf``.  loss_scale: static scale of the fp16
        activation gradients (removed again on export); default: the power of two that puts 4..8 on each supervised row's d loss / d logits
        (the mean over n rows carries 1/n), so the scale does not depend on the batch size.  train_encoder: also back-propagate through the point
        encoder of cond_mode 'point' and the KL term (opt.freeze_encoder = False); otherwise both are constants (the Options default).
        loss_scale_mult: factor on the default scale (a trainer's overflow back-off)."""
        B, T = tokens.shape
        is_latent = int(self.opt.cond_mode == 'point_latent')
        conds = conds.to(self.device, torch.float32).contiguous()
        tok = tokens.to(self.device, torch.int32).contiguous()
        lab = labels.to(self.device, torch.int64).contiguous()
        nf = (C.c_int32 * B)(*[int(x) for x in num_faces])
        if loss_scale is None:
            n_sup = max(1, int((lab[:, 1:] >= 0).sum().item()))
            loss_scale = float(2 ** int(np.ceil(np.log2(4.0 * n_sup)))) * float(loss_scale_mult)
        losses = torch.zeros(3, dtype=torch.float32, device=self.device)
        sums = torch.zeros(3, dtype=torch.float64, device=self.device)
        m8 = None
        if masks is not None:
            m = masks.to(self.device).bool()
            if m.shape != (B, self.P + T):
                raise ValueError(f'masks must be [B, P+T] = {(B, self.P + T)}, got {tuple(m.shape)}')
            if bool((m[:, 1:] & ~m[:, :-1]).any()):
                raise NotImplementedError('only right-padded attention masks are supported (collate_fn pads at the end)')
            if not bool(m.all()):
                m8 = m.to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_train_step(self.h, conds.data_ptr(), conds.shape[1], is_latent, tok.data_ptr(), lab.data_ptr(),
                                              m8.data_ptr() if m8 is not None else None, nf, B, T, C.c_float(kl_weight), C.c_float(dropout_p),
                                              C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_float(loss_scale), int(bool(train_encoder)), losses.data_ptr(), sums.data_ptr(),
                                              _stream()))
        self.encoder_trained = bool(train_encoder)
        self._keep = [m8, conds, tok, lab]
        return losses, sums

    def grad_has(self, name: str, train_encoder: Optional[bool] = None) -> bool:
        """does ``name`` receive a gradient?  point_encoder.* entries only when the encoder is trained (default: as in the last train_step)"""
        te = getattr(self, 'encoder_trained', False) if train_encoder is None else train_encoder
        if name.startswith('point_encoder.') and not te:
            return False
        return bool(self.lib.er_grad_has(self.h, name.encode()))

    def grad(self, name: str, shape=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 gradient of one state-dict entry from the last ``train_step`` (loss scale removed); ``out``: a contiguous fp32 CUDA tensor to fill."""
        if out is None:
            out = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.er_grad_get(self.h, name.encode(), out.data_ptr(), out.numel(), _stream()))
        return out

    # ---- introspection ----------------------------------------------------------------------------------------------------
    def weight_bytes_per_token(self) -> int:
        return int(self.lib.er_weight_bytes_per_token(self.h))

    def kv_bytes_per_row(self) -> int:
        return int(self.lib.er_kv_bytes_per_row(self.h))

    def kernel_launches(self) -> int:
        return int(self.lib.er_kernel_launches(self.h))
