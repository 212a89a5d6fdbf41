"""Training step (forward in training mode + backward) against torch autograd over the oracle's restatement of LMM.forward
(oracle/train_oracle.py; reference main.py:168-172, models.py:147-202, modeling_opt.py:253-298, 464-517).

Tolerances: the engine's activations and activation gradients are fp16 (the reference trains under bf16 autocast), the oracle is fp32 on the
same fp16-rounded weights, so a gradient tensor is accepted when ||g - g_ref|| <= GRAD_RTOL * ||g_ref|| (+ a floor for tensors whose gradient
is itself at the fp16 noise level)."""

import ctypes as C

import numpy as np
import pytest
import torch

from edgerunner_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

GRAD_RTOL = 3e-2


def _batch(opt, B=2, T=40, seed=0, pad=0):
    g = torch.Generator().manual_seed(seed)
    V = synth.vocab_size_of(opt)
    P = opt.num_cond_tokens
    tokens = torch.randint(6, V, (B, T), generator=g)
    tokens[:, 0] = opt.bos_token_id
    labels = torch.cat([torch.full((B, P), -100, dtype=torch.long), tokens.long()], 1)
    masks = torch.ones(B, P + T, dtype=torch.bool)
    if pad:
        masks[B - 1, -pad:] = False
        labels[B - 1, -pad:] = -100
        tokens[B - 1, -pad:] = opt.pad_token_id
    if opt.cond_mode == 'point':
        conds = torch.cat([synth.synth_point_cloud(seed + b, opt.point_num) for b in range(B)])
    else:
        conds = torch.randn(B, opt.point_latent_size, opt.point_latent_dim, generator=g) * 0.5
    return conds, tokens, labels, masks, [1500, 5000, 300, 9000][:B]


class _debug:
    """process-wide experiment switches of the library (er_debug_set(NULL, key, value)), restored on exit"""

    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        from edgerunner_b200 import _lib
        for k, v in self.kv.items():
            _lib.check(_lib.load().er_debug_set(None, k.encode(), int(v)))

    def __exit__(self, *exc):
        from edgerunner_b200 import _lib
        for k in self.kv:
            _lib.check(_lib.load().er_debug_set(None, k.encode(), DEFAULTS[k]))


DEFAULTS = {'attn_bwd_wmma': 0, 'train_fwd_lse': 1, 'train_recompute': 0}


@pytest.mark.parametrize('impl', ['mma', 'wmma'])
def test_attention_backward_matches_autograd(impl):
    """the attention() seam's backward (er_attention_bwd_bnhd) against autograd of the fp32 softmax attention on the same fp16 inputs; both
    implementations: register-resident mma.sync (attention_bwd_mma.cu, the default) and wmma through shared memory (backward.cu)"""
    with _debug(attn_bwd_wmma=impl == 'wmma'):
        _attention_backward_cases()


def _attention_backward_cases():
    from core.transformer.attention import attention
    from edgerunner_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    for (B, N, M, H, D, causal) in [(2, 150, 150, 2, 96, True), (1, 70, 333, 2, 64, False), (1, 64, 64, 1, 96, True), (1, 257, 257, 3, 96, True)]:
        q = torch.randn(B, N, H, D, device='cuda', dtype=torch.float16)
        k = torch.randn(B, M, H, D, device='cuda', dtype=torch.float16)
        v = torch.randn(B, M, H, D, device='cuda', dtype=torch.float16)
        do = torch.randn(B, N, H, D, device='cuda', dtype=torch.float16)
        out = attention(q, k, v, causal=causal).contiguous()
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        _lib.check(lib.er_attention_bwd_bnhd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                             dv.data_ptr(), B, N, M, H, D, int(causal), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        qf, kf, vf = (t.float().transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
        w = qf @ kf.transpose(-1, -2) / D ** 0.5
        if causal:
            w = w + torch.triu(torch.full((N, M), float('-inf'), device='cuda'), diagonal=1)
        ref = torch.softmax(w, -1) @ vf
        ref.backward(do.float().transpose(1, 2))
        for name, got, want in (('dq', dq, qf.grad), ('dk', dk, kf.grad), ('dv', dv, vf.grad)):
            want = want.transpose(1, 2)
            err = (got.float() - want).norm() / want.norm()
            assert float(err) < 1e-2, (B, N, M, H, D, causal, name, float(err))
            assert float((got.float() - want).abs().max()) < 2e-2 * float(want.abs().max()) + 2e-3, (name, B, N, M, H, D)


def _engine_grads(opt, sd, batch, dropout_p, seed, use_masks, train_encoder=False):
    from edgerunner_b200.engine import Engine
    conds, tokens, labels, masks, nf = batch
    B, T = tokens.shape
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=64, max_points=opt.point_num, max_tf_rows=B * (opt.num_cond_tokens + T))
    eng.load_state_dict(sd)
    losses, sums = eng.train_step(conds.cuda(), tokens, labels, nf, opt.kl_weight, masks=masks if use_masks else None, dropout_p=dropout_p, seed=seed,
                                   train_encoder=train_encoder)
    grads = {k: eng.grad(k, v.shape).cpu() for k, v in sd.items() if eng.grad_has(k)}
    return eng, losses.cpu().numpy(), sums.cpu().numpy(), grads


def _oracle_grads(opt, sd, batch, dropout_p, seed, use_masks, train_encoder=False):
    from oracle.train_oracle import forward_train, trainable_leaves
    conds, tokens, labels, masks, nf = batch
    w = trainable_leaves(sd, train_encoder)
    out = forward_train(opt, sd, w, conds, tokens, labels, nf, masks=masks if use_masks else None, dropout_p=dropout_p, seed=seed,
                        train_encoder=train_encoder)
    out['loss'].backward()
    return {k: float(v) for k, v in out.items()}, {k: v.grad for k, v in w.items()}


def _compare(grads, ref, rtol=GRAD_RTOL):
    assert set(grads) == set(ref), (sorted(set(ref) - set(grads))[:5], sorted(set(grads) - set(ref))[:5])
    gmax = max(float(v.norm()) for v in ref.values())
    bad = []
    for k, want in ref.items():
        got = grads[k]
        assert bool(torch.isfinite(got).all()), k
        err = float((got - want).norm())
        if err > rtol * float(want.norm()) + 2e-4 * gmax:
            bad.append((k, err / max(float(want.norm()), 1e-30), float(want.norm())))
    assert not bad, bad[:12]


@pytest.mark.parametrize('cond_mode', ['point', 'point_latent'])
def test_train_step_gradients_match_autograd(cond_mode):
    opt = synth.tiny_options(cond_mode=cond_mode)
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    batch = _batch(opt)
    eng, losses, sums, grads = _engine_grads(opt, sd, batch, 0.0, 0, False)
    ref_losses, ref = _oracle_grads(opt, sd, batch, 0.0, 0, False)
    np.testing.assert_allclose(losses[1], ref_losses['loss_ce'], rtol=2e-3)
    assert not any(k.startswith('point_encoder.') for k in grads)
    _compare(grads, ref)
    # dropout off: the training forward is the eval forward
    conds, tokens, labels, masks, nf = batch
    l_eval, _ = eng.forward_tf(conds.cuda(), tokens, labels, nf, opt.kl_weight)
    np.testing.assert_allclose(losses, l_eval.cpu().numpy(), rtol=1e-6)
    # bit-reproducible (no atomics anywhere)
    l2, _ = eng.train_step(conds.cuda(), tokens, labels, nf, opt.kl_weight, dropout_p=0.0, seed=0)
    assert np.array_equal(l2.cpu().numpy(), losses)
    k = 'mesh_decoder.model.layers.0.fc1.weight'
    assert torch.equal(eng.grad(k, sd[k].shape).cpu(), grads[k])


def test_train_step_trains_the_point_encoder():
    """opt.freeze_encoder = False (the ArAE preset): the backward continues through proj_cond into the point encoder (cross-attention with D = 64,
    GEGLU, three LayerNorms, the Fourier-feature Linear, query_embed) and the KL term kl_weight * 0.5 sum(latent^2) contributes (kl_weight raised
    from the preset's 1e-8 so that its gradient is visible); three clouds, so the per-cloud accumulation of the encoder's weight gradients is covered"""
    opt = synth.tiny_options(kl_weight=3e-3)
    sd = synth.synth_state_dict(opt, seed=2, eos_logit=-30.0)
    batch = _batch(opt, B=3, T=30, seed=4, pad=5)
    eng, losses, sums, grads = _engine_grads(opt, sd, batch, 0.1, 99, True, train_encoder=True)
    ref_losses, ref = _oracle_grads(opt, sd, batch, 0.1, 99, True, train_encoder=True)
    np.testing.assert_allclose(losses[1], ref_losses['loss_ce'], rtol=3e-3)
    np.testing.assert_allclose(losses[2], ref_losses['loss_kl'], rtol=5e-3)
    assert sum(k.startswith('point_encoder.') for k in grads) == 23 and 'point_encoder.point_embed.basis' not in grads
    _compare(grads, ref)
    # with the encoder frozen the same engine refuses to hand out encoder gradients and the others lose the encoder-side KL / nothing else
    conds, tokens, labels, masks, nf = batch
    eng.train_step(conds.cuda(), tokens, labels, nf, opt.kl_weight, masks=masks, dropout_p=0.1, seed=99, train_encoder=False)
    assert not eng.grad_has('point_encoder.linear.weight')
    from edgerunner_b200 import _lib
    with pytest.raises(_lib.ErError):
        eng.grad('point_encoder.linear.weight', sd['point_encoder.linear.weight'].shape)
    k = 'mesh_decoder.model.layers.1.fc2.weight'
    assert torch.equal(eng.grad(k, sd[k].shape).cpu(), grads[k])          # the decoder's gradients do not depend on whether the encoder is trained


def test_train_step_with_dropout_and_padding():
    """dropout 0.1 (the reference's config.dropout) through the restated counter-based mask, one right-padded sample (collate_fn)"""
    opt = synth.tiny_options()
    sd = synth.synth_state_dict(opt, seed=1, eos_logit=-30.0)
    batch = _batch(opt, B=3, T=33, seed=5, pad=7)
    _, losses, sums, grads = _engine_grads(opt, sd, batch, 0.1, 1234567, True)
    ref_losses, ref = _oracle_grads(opt, sd, batch, 0.1, 1234567, True)
    np.testing.assert_allclose(losses[1], ref_losses['loss_ce'], rtol=3e-3)
    assert int(sums[1]) == int((batch[2][:, 1:] >= 0).sum())
    _compare(grads, ref)
    # a different seed gives a different mask, hence a different loss
    _, losses2, _, _ = _engine_grads(opt, sd, batch, 0.1, 7654321, True)
    assert losses2[1] != losses[1]


@pytest.mark.parametrize('variant', ['default', 'stats_pass', 'wmma', 'recompute', 'recompute+stats_pass', 'recompute+wmma'])
def test_train_step_mid_size(variant):
    """hidden 768 / 8 heads / 3 layers, 2 x 300 rows: several attention tiles per head, padded transposes; every variant of the backward:
    activations kept from the forward pass (default) | per-layer recomputation (the reference's opt.checkpointing); attention backward with
    mma.sync (default) | wmma kernels; row log-sum-exp from the forward kernel (default) | from a statistics pass"""
    with _debug(attn_bwd_wmma='wmma' in variant, train_fwd_lse='stats_pass' not in variant, train_recompute='recompute' in variant):
        _mid_size()


def _mid_size():
    opt = synth.tiny_options(hidden_dim=768, num_heads=8, num_layers=3, cond_mode='point_latent')
    sd = synth.synth_state_dict(opt, seed=3, eos_logit=-30.0)
    batch = _batch(opt, B=2, T=235, seed=2, pad=11)
    _, losses, _, grads = _engine_grads(opt, sd, batch, 0.0, 0, True)
    ref_losses, ref = _oracle_grads(opt, sd, batch, 0.0, 0, True)
    np.testing.assert_allclose(losses[1], ref_losses['loss_ce'], rtol=3e-3)
    _compare(grads, ref)


def test_lmm_train_mode_backward_and_optimizer_step():
    """the drop-in surface: model.train(); out = model(data); out['loss'].backward(); clip; AdamW step (main.py:160-181) — the loss of a fixed
    batch goes down, the frozen encoder gets no gradient, eval mode still works afterwards"""
    from core.models import LMM
    opt = synth.tiny_options(freeze_encoder=True, nof_dropout_ratio=0.0)
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    model = LMM(opt)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    model.config.dropout = 0.0
    conds, tokens, labels, masks, nf = _batch(opt, B=2, T=40)
    data = {'conds': conds.cuda(), 'tokens': tokens.cuda(), 'labels': labels.cuda(), 'masks': masks.cuda(), 'num_faces': torch.tensor(nf).cuda(),
            'num_tokens': torch.full((2,), 40).cuda()}
    optim = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, betas=(0.9, 0.95))
    hist = []
    for step in range(5):
        optim.zero_grad()
        out = model(data)
        out['loss'].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        optim.step()
        hist.append(float(out['loss']))
    assert hist[-1] < hist[0] - 0.05, hist
    assert all(p.grad is None for n, p in model.named_parameters() if n.startswith('point_encoder.'))
    assert model.mesh_decoder.lm_head.weight.grad is not None and float(model.mesh_decoder.lm_head.weight.grad.abs().max()) > 0
    model.eval()
    with torch.no_grad():
        ev = model(data)
    assert float(ev['loss']) < hist[0]
    # freeze_encoder = False (the ArAE preset): the encoder's parameters receive gradients through the same autograd node
    opt2 = synth.tiny_options(freeze_encoder=False, nof_dropout_ratio=0.0)
    m2 = LMM(opt2)
    m2.load_state_dict(sd, strict=True)
    m2 = m2.cuda().train()
    out2 = m2(data)
    out2['loss'].backward()
    enc = [p for n, p in m2.named_parameters() if n.startswith('point_encoder.')]
    assert enc and all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in enc)
    assert float(m2.point_encoder.query_embed.grad.abs().max()) > 0


def test_flat_trainer_steps_reduce_the_loss():
    """the native loop (edgerunner_b200/train.py): er_train_step -> flat gradient buffer -> clip -> fused AdamW -> fp16 weights back into the engine"""
    from core.models import LMM
    from edgerunner_b200.train import FlatTrainer
    opt = synth.tiny_options(freeze_encoder=True, nof_dropout_ratio=0.0, lr=1e-3, warmup_ratio=0.0)
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    model = LMM(opt)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    conds, tokens, labels, masks, nf = _batch(opt, B=2, T=40)
    data = {'conds': conds.cuda(), 'tokens': tokens.cuda(), 'labels': labels.cuda(), 'masks': masks.cuda(), 'num_faces': torch.tensor(nf).cuda()}
    tr = FlatTrainer(model, total_steps=100, max_batch=2, max_tokens=40)
    assert tr.numel >= sum(p.numel() for n, p in model.named_parameters() if not n.startswith('point_encoder.'))
    hist = [float(tr.step(data)['loss']) for _ in range(6)]
    assert all(np.isfinite(hist)) and hist[-1] < hist[0] - 0.05, hist
    # the eval forward reads the updated weights
    tr.sync_to_model()
    model.eval()
    with torch.no_grad():
        ev = model(data)
    assert float(ev['loss']) < hist[0]


def test_attention_seam_is_differentiable():
    """core.transformer.attention.attention() with inputs that require grad (what the reference gets from flash-attn's autograd function)"""
    from core.transformer.attention import attention
    torch.manual_seed(1)
    for (B, N, M, H, D, causal, dt) in [(2, 100, 100, 2, 96, True, torch.float16), (1, 50, 130, 2, 64, False, torch.float32)]:
        q = torch.randn(B, N, H, D, device='cuda', dtype=dt).requires_grad_(True)
        k = torch.randn(B, M, H, D, device='cuda', dtype=dt).requires_grad_(True)
        v = torch.randn(B, M, H, D, device='cuda', dtype=dt).requires_grad_(True)
        do = torch.randn(B, N, H, D, device='cuda', dtype=dt)
        out = attention(q, k, v, causal=causal)
        assert out.dtype == dt and out.requires_grad
        out.backward(do)
        qf, kf, vf = (t.detach().half().float().transpose(1, 2).requires_grad_(True) for t in (q, k, v))
        w = qf @ kf.transpose(-1, -2) / D ** 0.5
        if causal:
            w = w + torch.triu(torch.full((N, M), float('-inf'), device='cuda'), diagonal=1)
        ref = torch.softmax(w, -1) @ vf
        ref.backward(do.half().float().transpose(1, 2))
        assert float((out.float() - ref.transpose(1, 2)).abs().max()) < 4e-3
        for name, got, want in (('dq', q.grad, qf.grad), ('dk', k.grad, kf.grad), ('dv', v.grad, vf.grad)):
            assert got is not None and got.dtype == dt
            err = (got.float() - want.transpose(1, 2)).norm() / want.norm()
            assert float(err) < 1e-2, (name, float(err))
    with torch.no_grad():                                               # no graph requested: the plain path
        assert not attention(q, k, v, causal=False).requires_grad


def test_flat_trainer_skips_a_step_whose_gradient_overflowed():
    """fp16 overflow back-off: a step run at an absurd loss scale produces a non-finite global gradient norm; the update is skipped (weights and
    optimizer state untouched), the loss scale halves, and the next ordinary step goes through"""
    from core.models import LMM
    from edgerunner_b200.train import FlatTrainer
    opt = synth.tiny_options(freeze_encoder=True, nof_dropout_ratio=0.0, lr=1e-3, warmup_ratio=0.0)
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    model = LMM(opt)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    conds, tokens, labels, masks, nf = _batch(opt, B=2, T=40)
    data = {'conds': conds.cuda(), 'tokens': tokens.cuda(), 'labels': labels.cuda(), 'masks': masks.cuda(), 'num_faces': torch.tensor(nf).cuda()}
    tr = FlatTrainer(model, total_steps=100, max_batch=2, max_tokens=40)
    before = tr.param.clone()
    out = tr.step(data, loss_scale=1e30)
    assert out['skipped'] and not bool(torch.isfinite(out['grad_norm']))
    assert torch.equal(tr.param, before) and tr.optim.step_count == 0 and tr.step_count == 0 and tr.scale_mult == 0.5 and tr.skipped_steps == 1
    out = tr.step(data)
    assert not out['skipped'] and bool(torch.isfinite(out['grad_norm'])) and not torch.equal(tr.param, before)
