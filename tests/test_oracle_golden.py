"""Pin the CPU oracle against golden vectors produced by EXECUTING THE REFERENCE (oracle/gen_golden.py).

CPU-only.  ``mode='fp32'`` of the oracle must reproduce the reference's own fp32 CPU run: same greedy /
sampled token ids and logits to 1e-4; the ledger mode (fp16 rounding points) must stay within 2e-2 of it
(it models a different — the GPU — precision of the same algorithm).
"""

import json
import os
from dataclasses import asdict, replace

import numpy as np
import pytest
import torch

from core.options import config_defaults
from edgerunner_b200 import synth
from oracle.er_oracle import ConstraintFSM, Oracle, quantize_num_faces


def _load(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        pytest.skip(f'{name} not generated')
    return np.load(p)


def test_options_match_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, 'options.json')))['options']
    for preset, fields in ref.items():
        mine = asdict(config_defaults[preset])
        mine = json.loads(json.dumps(mine))  # tuples -> lists like the golden
        assert mine == fields, preset


def test_quantize_num_faces(golden_dir):
    from core.utils import quantize_num_faces as qprod
    ref = json.load(open(os.path.join(golden_dir, 'quantize_num_faces.json')))
    for n, q in ref.items():
        assert quantize_num_faces(int(n)) == q
        assert qprod(int(n)) == q
    t = torch.tensor([int(n) for n in ref])
    assert qprod(t).tolist() == list(ref.values())


def test_fsm_grammar():
    fsm = ConstraintFSM(518)
    gen = []
    assert fsm.allowed(gen) == [5]
    gen.append(5)
    for i in range(9):
        assert fsm.allowed(gen) == list(range(6, 518))
        gen.append(100 + i)
    assert fsm.allowed(gen) == [3, 4, 5, 2]
    gen.append(3)
    for i in range(3):
        assert fsm.allowed(gen) == list(range(6, 518))
        gen.append(7)
    assert fsm.allowed(gen) == [3, 4, 5, 2]


@pytest.fixture(scope='module')
def tiny():
    opt = synth.tiny_options()
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    return opt, sd


def test_tiny_fp32_matches_reference(golden_dir, tiny):
    g = _load(golden_dir, 'tiny.npz')
    opt, sd = tiny
    orc = Oracle(opt, sd, mode='fp32')
    cond = synth.synth_point_cloud(0, opt.point_num)
    lat = orc.encode_points(cond)[0].numpy()
    np.testing.assert_allclose(lat, g['latents'], atol=2e-5, rtol=1e-4)
    ce = orc.encode_cond(cond, 1000)[0].numpy()
    np.testing.assert_allclose(ce[:4], g['cond_embeds_head'], atol=1e-4)
    np.testing.assert_allclose(ce[-2:], g['cond_embeds_tail'], atol=1e-4)
    T = len(g['greedy_tokens'])
    out = orc.generate(cond, 1000, max_new_tokens=T, generate_mode='greedy')
    np.testing.assert_array_equal(out['tokens'], g['greedy_tokens'])
    np.testing.assert_allclose(out['logits_pre'].numpy(), g['greedy_logits'], atol=1e-4)


def test_tiny_sample_matches_reference(golden_dir, tiny):
    g = _load(golden_dir, 'tiny.npz')
    opt, sd = tiny
    orc = Oracle(opt, sd, mode='fp32')
    cond = synth.synth_point_cloud(0, opt.point_num)
    T = 64
    torch.manual_seed(1234)                     # same global-RNG protocol as gen_golden.run_generate
    out = orc.generate(cond, 1000, max_new_tokens=T, generate_mode='sample')
    np.testing.assert_array_equal(out['tokens'], g['sample_tokens'])
    np.testing.assert_allclose(out['logits_pre'].numpy(), g['sample_logits'], atol=1e-4)


def test_tiny_ledger_close_to_fp32(golden_dir, tiny):
    g = _load(golden_dir, 'tiny.npz')
    opt, sd = tiny
    orc = Oracle(opt, sd, mode='ledger')
    cond = synth.synth_point_cloud(0, opt.point_num)
    T = 48
    out = orc.generate(cond, 1000, max_new_tokens=T, generate_mode='greedy', forced_tokens=list(g['greedy_tokens'][:T]))
    err = np.abs(out['logits_pre'].numpy() - g['greedy_logits'][:T]).max()
    assert err < 2e-2, err


def test_tiny_teacher_forced_forward(golden_dir, tiny):
    g = _load(golden_dir, 'tiny.npz')
    opt, sd = tiny
    orc = Oracle(opt, sd, mode='fp32')
    conds = torch.cat([synth.synth_point_cloud(b, opt.point_num) for b in range(2)])
    res = orc.forward_tf(conds, torch.from_numpy(g['tf_tokens']), torch.from_numpy(g['tf_labels']), g['tf_num_faces'])
    np.testing.assert_allclose(float(res['loss']), g['tf_loss'][0], rtol=1e-5)
    np.testing.assert_allclose(float(res['loss_ce']), g['tf_loss'][1], rtol=1e-5)
    np.testing.assert_allclose(float(res['loss_kl']), g['tf_loss'][2], rtol=1e-4)
    np.testing.assert_allclose(res['logits_pre'][:, -8:].numpy(), g['tf_logits_tail'], atol=1e-4)


@pytest.mark.slow
def test_arae_fp32_matches_reference(golden_dir):
    g = _load(golden_dir, 'arae.npz')
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    orc = Oracle(opt, sd, mode='fp32')
    del sd
    cond = synth.synth_point_cloud(0, opt.point_num)
    T = 12
    out = orc.generate(cond, 1000, max_new_tokens=T, generate_mode='greedy')
    np.testing.assert_allclose(out['cond_embeds'][:4].numpy(), g['cond_embeds_head'], atol=2e-4)
    np.testing.assert_array_equal(out['tokens'], g['greedy_tokens'][:T])
    np.testing.assert_allclose(out['logits_pre'].numpy(), g['greedy_logits'][:T], atol=3e-4)
