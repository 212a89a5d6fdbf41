"""Training-step oracle (oracle/train_oracle.py) pinned on the CPU: its forward equals the pinned teacher-forced oracle when dropout is off, its
dropout generator is the restated counter-based one, autograd reaches every trainable tensor."""

import numpy as np
import torch

from edgerunner_b200 import synth
from oracle.er_oracle import Oracle
from oracle.train_oracle import dropout_keep, forward_train, trainable_leaves


def _batch(opt, B=2, T=24, seed=0):
    g = torch.Generator().manual_seed(seed)
    V = synth.vocab_size_of(opt)
    P = opt.num_cond_tokens
    tokens = torch.randint(6, V, (B, T), generator=g)
    tokens[:, 0] = opt.bos_token_id
    labels = torch.cat([torch.full((B, P), -100, dtype=torch.long), tokens.long()], 1)
    if opt.cond_mode == 'point':
        conds = torch.cat([synth.synth_point_cloud(seed + b, opt.point_num) for b in range(B)])
    else:
        conds = torch.randn(B, opt.point_latent_size, opt.point_latent_dim, generator=g) * 0.5
    return conds, tokens, labels, [1500, 5000][:B]


def test_forward_train_equals_pinned_forward_tf_without_dropout():
    opt = synth.tiny_options(cond_mode='point_latent')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    sd16 = {k: v.to(torch.float16).float() for k, v in sd.items()}
    conds, tokens, labels, nf = _batch(opt)
    ref = Oracle(opt, sd16, mode='fp32').forward_tf(conds.to(torch.float16).float(), tokens, labels, nf)
    w = trainable_leaves(sd)
    out = forward_train(opt, sd, w, conds, tokens, labels, nf)
    assert abs(float(out['loss']) - float(ref['loss'])) < 2e-5 * max(1.0, abs(float(ref['loss'])))


def test_autograd_reaches_every_trainable_tensor():
    opt = synth.tiny_options()
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    conds, tokens, labels, nf = _batch(opt)
    w = trainable_leaves(sd)
    assert not any(k.startswith('point_encoder.') for k in w)
    masks = torch.ones(2, opt.num_cond_tokens + tokens.shape[1], dtype=torch.bool)
    masks[1, -5:] = False
    labels[1, -5:] = -100
    out = forward_train(opt, sd, w, conds, tokens, labels, nf, masks=masks, dropout_p=0.1, seed=7)
    out['loss'].backward()
    for k, v in w.items():
        assert v.grad is not None and bool(torch.isfinite(v.grad).all()), k
    used = w['mesh_decoder.model.embed_positions.weight'].grad.abs().sum(1) > 0
    N = opt.num_cond_tokens + tokens.shape[1]
    assert bool(used[:N - 1].all()) and not bool(used[N:].any())
    assert 'loss_kl' in out and not out['loss_kl'].requires_grad          # frozen encoder: the KL term has no graph


def test_dropout_generator():
    a = dropout_keep(1234, 0, 200000, 0.1)
    assert abs(a.mean() - 0.9) < 3e-3
    assert np.array_equal(a, dropout_keep(1234, 0, 200000, 0.1))
    assert not np.array_equal(a, dropout_keep(1234, 1, 200000, 0.1)) and not np.array_equal(a, dropout_keep(1235, 0, 200000, 0.1))
    assert dropout_keep(1, 0, 10, 0.0).all()
    # known answers of the restated splitmix64 (seed 0, site 0, first elements): guards the CUDA / numpy pair against drifting apart
    z = []
    for i in range(4):
        x = (0x9E3779B97F4A7C15 * (i + 1) + 0xD1B54A32D192ED03) & (2 ** 64 - 1)
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        x ^= x >> 31
        z.append((x >> 32) >= int(0.5 * 2 ** 32))
    assert list(dropout_keep(0, 0, 4, 0.5)) == z


def test_oracle_gradients_match_the_reference_backward(golden_dir):
    """tests/golden/train.npz was produced by executing the REFERENCE's training step (model.train(); model(data)['loss'].backward(), fp32 CPU, dropout
    off) — oracle/gen_golden.py --only train.  The oracle's autograd over its restated forward must reproduce every parameter's gradient."""
    import os
    g = np.load(os.path.join(golden_dir, 'train.npz'))
    tokens, labels, nf = torch.from_numpy(g['tokens']), torch.from_numpy(g['labels']), g['num_faces'].tolist()
    for tag, cond_mode in (('point', 'point'), ('latent', 'point_latent')):
        opt = synth.tiny_options(cond_mode=cond_mode, kl_weight=3e-3)
        sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
        if cond_mode == 'point':
            conds = torch.cat([synth.synth_point_cloud(b, opt.point_num) for b in range(2)])
        else:
            conds = torch.from_numpy(g['latent_conds'])
        w = trainable_leaves(sd, train_encoder=cond_mode == 'point', round_fp16=False)
        out = forward_train(opt, sd, w, conds, tokens, labels, nf, train_encoder=cond_mode == 'point', exact_fp32=True)
        out['loss'].backward()
        ref_loss = g[f'{tag}_loss']
        assert abs(float(out['loss'].detach()) - ref_loss[0]) < 2e-5 * abs(ref_loss[0])
        names = [str(n) for n in g[f'{tag}_names']]
        assert set(names) == set(w), (set(names) ^ set(w))
        gmax = max(float(g[f'{tag}|{n}|norm']) for n in names)
        for n in names:
            got = w[n].grad.double().reshape(-1).numpy()
            norm = float(g[f'{tag}|{n}|norm'])
            assert abs(np.linalg.norm(got) - norm) <= 1e-4 * norm + 1e-7 * gmax, (tag, n, np.linalg.norm(got), norm)
            idx = np.unique(np.linspace(0, got.size - 1, num=min(48, got.size)).astype(np.int64))
            np.testing.assert_allclose(got[idx], g[f'{tag}|{n}|probe'], rtol=2e-3, atol=2e-6 * gmax, err_msg=f'{tag} {n}')
