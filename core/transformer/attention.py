"""The attention op seam.  Mirrors ``/root/reference/core/transformer/attention.py:27-95``:
``attention(q, k, v, mask_q=None, mask_kv=None, dropout=0, causal=False)`` with ``[B, N, H, D]`` tensors.

fp16 CUDA inputs run on the sm_100a tcgen05 flash kernel of ``edgerunner_b200`` (head_dim 64 or 96) instead of flash-attn.
Padding masks (the varlen branch, reference ``:65-93``: unpad -> ``flash_attn_varlen_func`` -> ``pad_input``): sample b attends over its kept
keys with its kept queries and the masked query rows come back as zeros (``pad_input``).  RIGHT-padded masks — what ``collate_fn`` produces —
run in place on the prefix; masks with holes are gathered / scattered per sample like ``unpad_input`` / ``pad_input`` do.
Like the reference's (flash-attn's) op, the unmasked call is differentiable: when an input requires grad the result carries an autograd node whose
backward is the library's flash-attention backward (``er_attention_bwd_bnhd``: fp16 gradients, deterministic).  Masked calls return no graph.
There is no eager / CPU fallback: non-CUDA inputs raise.
"""

import torch

from edgerunner_b200 import _lib

FLASH_ATTN_AVAILABLE = False   # this implementation never calls flash-attn


class _AttentionFn(torch.autograd.Function):
    """dense / causal attention with the library's backward (q, k, v: [B, N|M, H, D]; the result and the gradients are computed in fp16)"""

    @staticmethod
    def forward(ctx, q, k, v, causal):
        q16, k16, v16 = (t.detach().to(torch.float16).contiguous() for t in (q, k, v))
        B, N, H, D = q16.shape
        M = k16.shape[1]
        out = torch.empty_like(q16)
        _lib.check(_lib.load().er_attention_bnhd(q16.data_ptr(), k16.data_ptr(), v16.data_ptr(), out.data_ptr(), B, N, M, H, D, int(causal),
                                                 torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(q16, k16, v16, out)
        ctx.causal = int(causal)
        ctx.dtypes = (q.dtype, k.dtype, v.dtype)
        return out.to(q.dtype)

    @staticmethod
    def backward(ctx, dout):
        q16, k16, v16, out = ctx.saved_tensors
        B, N, H, D = q16.shape
        M = k16.shape[1]
        do16 = dout.detach().to(torch.float16).contiguous()
        dq, dk, dv = torch.empty_like(q16), torch.empty_like(k16), torch.empty_like(v16)
        _lib.check(_lib.load().er_attention_bwd_bnhd(q16.data_ptr(), k16.data_ptr(), v16.data_ptr(), out.data_ptr(), do16.data_ptr(), dq.data_ptr(),
                                                     dk.data_ptr(), dv.data_ptr(), B, N, M, H, D, ctx.causal, torch.cuda.current_stream().cuda_stream))
        return dq.to(ctx.dtypes[0]), dk.to(ctx.dtypes[1]), dv.to(ctx.dtypes[2]), None


def attention(q, k, v, mask_q=None, mask_kv=None, dropout=0, causal=False):
    B, N, H, D = q.shape
    M = k.shape[1]
    if causal:
        assert N == 1 or N == M, 'Causal mask only supports self-attention'
    if dropout:
        raise NotImplementedError('attention dropout is a training-time op; not part of the B200 decode path')
    if not q.is_cuda:
        raise RuntimeError('edgerunner_b200 attention needs CUDA tensors (no CPU fallback)')
    if mask_q is None and mask_kv is None and torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        return _AttentionFn.apply(q, k, v, bool(causal and N > 1))
    in_dtype = q.dtype
    q16, k16, v16 = (t.to(torch.float16).contiguous() for t in (q, k, v))
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    if mask_q is None and mask_kv is None:
        out = torch.empty_like(q16)
        _lib.check(lib.er_attention_bnhd(q16.data_ptr(), k16.data_ptr(), v16.data_ptr(), out.data_ptr(), B, N, M, H, D,
                                         1 if (causal and N > 1) else 0, stream))
        return out.to(in_dtype)
    # varlen branch for right-padded masks (reference :65-93: a missing mask counts as all-True)
    mq = torch.ones(B, N, dtype=torch.bool, device=q.device) if mask_q is None else mask_q.bool()
    mk = torch.ones(B, M, dtype=torch.bool, device=q.device) if mask_kv is None else mask_kv.bool()
    right_padded = not any(bool((m[:, 1:] & ~m[:, :-1]).any()) for m in (mq, mk))
    len_q, len_k = mq.sum(1).tolist(), mk.sum(1).tolist()
    out = torch.zeros_like(q16)                                   # pad_input: masked query rows are zero
    row = H * D * 2                                              # bytes per (batch, position) row
    for b in range(B):
        nq, nk = int(len_q[b]), int(len_k[b])
        if nq == 0 or nk == 0:
            continue
        if causal and nq > 1 and nq != nk:
            raise NotImplementedError('causal varlen attention needs equal query / key lengths per sample')
        c = 1 if (causal and nq > 1) else 0
        if right_padded:                                         # collate_fn's case: the kept rows are a prefix, no copies
            _lib.check(lib.er_attention_bnhd(q16.data_ptr() + b * N * row, k16.data_ptr() + b * M * row, v16.data_ptr() + b * M * row,
                                             out.data_ptr() + b * N * row, 1, nq, nk, H, D, c, stream))
        else:                                                    # masks with holes: unpad_input (gather) -> attention -> pad_input (scatter)
            iq, ik = mq[b].nonzero().squeeze(1), mk[b].nonzero().squeeze(1)
            qb, kb, vb = q16[b].index_select(0, iq), k16[b].index_select(0, ik), v16[b].index_select(0, ik)
            ob = torch.empty_like(qb)
            _lib.check(lib.er_attention_bnhd(qb.data_ptr(), kb.data_ptr(), vb.data_ptr(), ob.data_ptr(), 1, nq, nk, H, D, c, stream))
            out[b].index_copy_(0, iq, ob)
    return out.to(in_dtype)
