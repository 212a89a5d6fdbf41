"""The attention op seam.  Mirrors ``/root/reference/core/transformer/attention.py:27-95``:
``attention(q, k, v, mask_q=None, mask_kv=None, dropout=0, causal=False)`` with ``[B, N, H, D]`` tensors.

Unmasked fp16 CUDA inputs run on the sm_100a flash-style kernel of ``edgerunner_b200`` (head_dim 64 or 96) instead of
flash-attn; like the reference without flash-attn, padding masks raise ``NotImplementedError``.  There is no eager /
CPU fallback: non-CUDA inputs raise.
"""

import torch

from edgerunner_b200 import _lib

FLASH_ATTN_AVAILABLE = False   # this implementation never calls flash-attn


def attention(q, k, v, mask_q=None, mask_kv=None, dropout=0, causal=False):
    B, N, H, D = q.shape
    M = k.shape[1]
    if causal:
        assert N == 1 or N == M, 'Causal mask only supports self-attention'
    if mask_q is not None or mask_kv is not None:
        raise NotImplementedError('masked (varlen) attention is not part of the B200 decode path')
    if dropout:
        raise NotImplementedError('attention dropout is a training-time op; not part of the B200 decode path')
    if not q.is_cuda:
        raise RuntimeError('edgerunner_b200 attention needs CUDA tensors (no CPU fallback)')
    in_dtype = q.dtype
    q16, k16, v16 = (t.to(torch.float16).contiguous() for t in (q, k, v))
    out = torch.empty_like(q16)
    lib = _lib.load()
    _lib.check(lib.er_attention_bnhd(q16.data_ptr(), k16.data_ptr(), v16.data_ptr(), out.data_ptr(), B, N, M, H, D,
                                     1 if (causal and N > 1) else 0, torch.cuda.current_stream().cuda_stream))
    return out.to(in_dtype)
