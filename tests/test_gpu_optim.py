"""Optimizer half of the training step against PyTorch's own AdamW / clip_grad_norm_ on the GPU (main.py:133,175-177)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flat_adamw_matches_torch_adamw_with_clipping():
    from edgerunner_b200.optim import FlatAdamW
    torch.manual_seed(0)
    shapes = [(257, 129), (1000,), (33, 7, 5), (3,)]                   # ragged total (not a multiple of 4)
    n = sum(int(np.prod(s)) for s in shapes)
    flat = torch.randn(n, device='cuda') * 0.05
    ref_params = []
    off = 0
    for s in shapes:
        k = int(np.prod(s))
        ref_params.append(torch.nn.Parameter(flat[off:off + k].clone().view(s)))
        off += k
    opt_ref = torch.optim.AdamW(ref_params, lr=1e-3, weight_decay=0.01, betas=(0.9, 0.95))
    p16 = torch.empty(n, dtype=torch.float16, device='cuda')
    opt = FlatAdamW(flat.clone(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01, param16=p16)
    for step in range(6):
        g = torch.randn(n, device='cuda') * (3.0 if step % 2 else 0.01)      # alternately above / below the clipping threshold
        off = 0
        for p in ref_params:
            p.grad = g[off:off + p.numel()].clone().view_as(p)
            off += p.numel()
        ref_norm = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        opt_ref.step()
        norm = opt.step(g, max_norm=1.0)
        assert abs(float(norm) - float(ref_norm)) <= 1e-5 * float(ref_norm)
        ref_flat = torch.cat([p.detach().reshape(-1) for p in ref_params])
        d = (opt.param - ref_flat).abs()
        assert float(d.max()) <= 2e-7 + 1e-6 * float(ref_flat.abs().max()), (step, float(d.max()))
        assert torch.equal(p16, opt.param.half())
    m_ref = torch.cat([opt_ref.state[p]['exp_avg'].reshape(-1) for p in ref_params])
    v_ref = torch.cat([opt_ref.state[p]['exp_avg_sq'].reshape(-1) for p in ref_params])
    assert float((opt.exp_avg - m_ref).abs().max()) <= 1e-6 * float(m_ref.abs().max()) + 1e-9
    assert float((opt.exp_avg_sq - v_ref).abs().max()) <= 1e-6 * float(v_ref.abs().max()) + 1e-12
    # no clipping requested -> plain AdamW
    g = torch.randn(n, device='cuda')
    before = opt.param.clone()
    assert opt.step(g) is None and not torch.equal(before, opt.param)
    with pytest.raises(RuntimeError):
        FlatAdamW(torch.zeros(8))                                       # CPU tensor: no fallback


def test_adamw_streaming_rate():
    """the fused step moves 30 bytes per parameter; at 200 M parameters it should run near the HBM roofline (reported, loosely asserted)"""
    from edgerunner_b200.optim import FlatAdamW
    n = 200_000_000
    p = torch.zeros(n, device='cuda'); g = torch.full((n,), 1e-3, device='cuda'); p16 = torch.empty(n, dtype=torch.float16, device='cuda')
    opt = FlatAdamW(p, param16=p16)
    for _ in range(2):
        opt.step(g, max_norm=1.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        opt.step(g, max_norm=1.0)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    gbs = n * (30 + 4) / (ms * 1e-3) / 1e9                              # + 4: the norm pass reads the gradient once more
    print(f'adamw + clip: {ms:.3f} ms for {n / 1e6:.0f} M parameters = {gbs:.0f} GB/s')
    assert gbs > 2500
