"""CPU checks of the DiT path: the oracle against the reference module's recorded output, the parameter containers against the reference's
state-dict schema, the scheduler restatement (diffusers is not installed: its algebraic identities are what can be checked here)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))


def test_oracle_fp32_matches_reference_dit(golden_dir):
    from dit_oracle import DitOracle, synth_dit_state
    g = np.load(os.path.join(golden_dir, 'dit.npz'))
    cfg = json.loads(str(g['cfg']))
    o = DitOracle(synth_dit_state(**cfg, seed=3), cfg['num_heads'], mode='fp32')
    out = o.forward(torch.from_numpy(g['x']), torch.from_numpy(g['c']), torch.from_numpy(g['t']))
    assert float(np.abs(g['out']).mean()) > 0.05
    np.testing.assert_allclose(out.numpy(), g['out'], atol=2e-5, rtol=1e-4)
    # the ledger differs from fp32 only by fp16 rounding noise
    led = DitOracle(synth_dit_state(**cfg, seed=3), cfg['num_heads'], mode='ledger').forward(torch.from_numpy(g['x']), torch.from_numpy(g['c']), torch.from_numpy(g['t']))
    assert float((led - out).abs().max()) < 3e-2 and float((led - out).abs().mean()) < 3e-3


def test_dit_containers_have_the_reference_schema(golden_dir):
    from core.transformer.dit import DiT
    g = np.load(os.path.join(golden_dir, 'dit.npz'))
    cfg = json.loads(str(g['cfg']))
    sd = DiT(**cfg).state_dict()
    assert sorted(sd.keys()) == [str(k) for k in g['keys']]
    for k, shp in zip(g['keys'], g['shapes']):
        assert ','.join(map(str, sd[str(k)].shape)) == str(shp), k
    with pytest.raises(RuntimeError):                      # no CPU fallback
        DiT(**cfg).eval()(torch.zeros(1, cfg['latent_size'], cfg['latent_dim']), torch.zeros(1, 257, cfg['hidden_dim']), torch.zeros(1))


def test_ddim_scheduler_restatement():
    from core.models_dit import DDIMScheduler, DDPMScheduler
    from dit_oracle import ddim_tables, ddim_step
    s = DDIMScheduler(prediction_type='v_prediction')
    s.set_timesteps(100)
    ts = s.timesteps
    assert ts[0] == 991 and ts[-1] == 1 and len(ts) == 100 and bool((ts[:-1] - ts[1:] == 10).all())     # leading spacing, steps_offset 1
    coef = s.step_coefficients(ts)
    ts2, coef2 = ddim_tables(100)
    assert np.array_equal(ts.numpy(), ts2) and torch.equal(coef, coef2)
    ac = s.alphas_cumprod
    assert abs(float(ac[0]) - (1 - 0.00085)) < 1e-6 and abs(float(1 - ac[999] / ac[998]) - 0.012) < 1e-5                # scaled-linear end points
    # known answer: this is the Stable Diffusion v1 schedule (scaled_linear 0.00085 .. 0.012, 1000 steps); its final alpha-bar is the widely
    # published 0.0047 (terminal SNR 0.0047 / 0.9953), its first 0.99915
    assert abs(float(ac[999]) - 0.004660) < 2e-6 and abs(float(ac[0]) - 0.99915) < 1e-6
    np.testing.assert_allclose(coef[:, 0] ** 2 + coef[:, 1] ** 2, 1.0, atol=1e-6)
    np.testing.assert_allclose(coef[:-1, 2], coef[1:, 0], atol=0)                   # alpha_prev of step i is alpha_t of step i + 1
    assert float(coef[-1, 2]) == float(ac[0] ** 0.5)                                # set_alpha_to_one=False: the last step lands on alphas_cumprod[0]
    # exactness property of DDIM (eta = 0): if the model predicts the true v / eps of x_t = a x0 + b e, one step lands on a' x0 + b' e
    g = torch.Generator().manual_seed(0)
    x0, e = torch.randn(64, generator=g), torch.randn(64, generator=g)
    for i in (0, 50, 99):
        a, b, ap, bp = [float(v) for v in coef[i]]
        xt = a * x0 + b * e
        for ptype, target in (('v_prediction', a * e - b * x0), ('epsilon', e)):
            nxt = ddim_step(target, xt, coef[i], ptype, ledger=False)
            np.testing.assert_allclose(nxt.numpy(), (ap * x0 + bp * e).numpy(), atol=2e-5)
    # add_noise / get_velocity of the training-side scheduler
    n = DDPMScheduler(prediction_type='v_prediction')
    t = torch.tensor([0, 500, 999])
    xs, es = torch.randn(3, 4, 2, generator=g), torch.randn(3, 4, 2, generator=g)
    noisy = n.add_noise(xs, es, t)
    a = (n.alphas_cumprod[t] ** 0.5).view(3, 1, 1)
    b = ((1 - n.alphas_cumprod[t]) ** 0.5).view(3, 1, 1)
    assert torch.allclose(noisy, a * xs + b * es) and torch.allclose(n.get_velocity(xs, es, t), a * es - b * xs)
    with pytest.raises(ValueError):
        DDIMScheduler(prediction_type='sample')


def test_mdit_schema_and_refusals():
    """MDiT with a two-layer CLIP tower: checkpoint keys of the reference (dit.*, proj_cond.*, norm_cond.*, image_encoder.*, point_encoder.*);
    the training loss raises; run() on CPU raises (no fallback)."""
    import dataclasses
    from core.options import config_defaults
    from core.models_dit import MDiT
    opt = dataclasses.replace(config_defaults['DiT'], dit_hidden_dim=128, dit_num_heads=2, dit_num_layers=1, point_latent_size=16, point_latent_dim=16,
                              point_hidden_dim=64, point_num_heads=1)
    tiny_clip = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, image_size=28, patch_size=14, hidden_act='gelu')
    m = MDiT(opt, image_encoder_config=tiny_clip).eval()
    keys = set(m.state_dict().keys())
    for k in ('dit.proj_in.weight', 'dit.layers.0.attn2.k_proj.bias', 'dit.layers.0.ff.net.2.weight', 'dit.scale_shift_table', 'proj_cond.weight',
              'norm_cond.bias', 'point_encoder.query_embed', 'image_encoder.vision_model.embeddings.class_embedding'):
        assert k in keys, k
    assert m.cond_tokens == 5 and m.proj_cond.in_features == 32
    with pytest.raises(NotImplementedError):
        m.forward({})
    with pytest.raises(RuntimeError):
        m.run(torch.rand(1, 3, 32, 32), num_inference_steps=2)


def test_cosine_lr_lambda_matches_main_py():
    """main.py:136-141: linear warm-up to 1 over warmup_ratio of the run, cosine decay to min_ratio = 0.1 at the end."""
    from edgerunner_b200.optim import cosine_lr_lambda as f
    T = 1000
    assert f(0, T, 0.01) == 0.0 and abs(f(5, T, 0.01) - 0.5) < 1e-12 and abs(f(10, T, 0.01) - 1.0) < 1e-12
    assert abs(f(T, T, 0.01) - 0.1) < 1e-12 and abs(f(505, T, 0.01) - 0.55) < 1e-9
    assert f(0, T, 0) == 1.0                                           # ArAE preset: warmup_ratio = 0
    vals = [f(s, T, 0.01) for s in range(10, T + 1)]
    assert all(a >= b for a, b in zip(vals, vals[1:]))
