"""GPU parity at the context lengths the metric is measured on (VERDICT r1 "weak" 1): ArAE WIDTH (C = 1536, 16 heads x 96, FFN 6144,
vocabulary 518, 2049-token condition prefix) at reduced DEPTH (2 layers) so that the CPU ledger oracle finishes in about a minute.

  * a free-running 16 000-token greedy decode — cache rows 2050 .. 18 049, i.e. every `L & 31` phase of the blocked K cache, every
    per-split block count from 8 to 63, thousands of ring wraps — checked at EVERY position, logits and ids, against the oracle
    teacher-forced on the CUDA stream (oracle.replay_steps: the same per-row arithmetic as 16 000 calls of step(), evaluated in chunks);
  * a decode that starts from an 8 050-row cache built by er_prefill from a long `resume_ids` prompt (reference:
    LMM.generate(resume_ids=...), core/models.py:222-223) against the oracle's own multi-row prefill + steps;
  * `point_latent` conditioning (core/models.py:126-129, infer_dit.py:111-113) and a C3-shaped sample run at full ArAE size.

Tolerances: same rule as tests/test_gpu_parity.py (ArAE width): max |dlogit| <= 8e-3, mean <= 1.5e-3 (north-star 1e-3 holds in the
mean), ids bit-exact outside the 2*tol + 1 fp16-ulp band of the oracle's own decision margin."""
import os
from dataclasses import replace

import numpy as np
import pytest
import torch

from core.options import config_defaults
from edgerunner_b200 import synth

pytestmark = pytest.mark.gpu

MAX_TOL, MEAN_TOL = 8e-3, 1.5e-3


def _fp16_ulp(x):
    return float(np.spacing(np.float16(abs(x))))


def grammar_stream(n, seed=0):
    """grammar-valid ids: BOM + 9 coords, then (L|R) + 3 coords ...  (models.py:245-271)"""
    rng = np.random.RandomState(seed)
    out = [5] + list(rng.randint(6, 518, 9))
    while len(out) < n:
        out += [int(rng.randint(3, 5))] + list(rng.randint(6, 518, 3))
    return [int(x) for x in out[:n]]


def fsm_masks(tokens, V, eos=2):
    """allowed-token mask [T, V] of the constraint FSM along a fed stream (models.py:252-268)"""
    m = np.zeros((len(tokens), V), dtype=bool)
    counter = 0
    for t in range(len(tokens)):
        if t == 0:
            m[t, 5] = True
        else:
            last = tokens[t - 1]
            if last == 5:
                counter = 9
            elif last in (3, 4):
                counter = 3
            elif last >= 6:
                counter -= 1
            if counter > 0:
                m[t, 6:] = True
            else:
                m[t, [3, 4, 5, eos]] = True
    return m


def check_stream(cuda_tokens, cuda_logits_pre, ref_logits_pre, V, label):
    """logits within tolerance at every position; ids equal to the oracle's argmax outside the near-tie band"""
    d = (cuda_logits_pre - ref_logits_pre).abs()
    T = len(cuda_tokens)
    q = [float(d[i * T // 4:(i + 1) * T // 4].mean()) for i in range(4)]
    print(f'[{label}] T={T} max |dlogit| {float(d.max()):.3e} mean {float(d.mean()):.3e} per-quarter mean {q}')
    assert not torch.isnan(cuda_logits_pre).any()
    assert float(d.max()) <= MAX_TOL and float(d.mean()) <= MEAN_TOL, (float(d.max()), float(d.mean()))
    sc = ref_logits_pre.to(torch.float16).float().numpy()           # what HF sees: the fp16 lm_head output
    sc = np.where(fsm_masks([int(x) for x in cuda_tokens], V), sc, -np.inf)
    ref_ids = sc.argmax(axis=1)
    mism = np.nonzero(ref_ids != np.asarray(cuda_tokens))[0]
    for t in mism:
        margin = sc[t, ref_ids[t]] - sc[t, cuda_tokens[t]]
        band = 2 * MAX_TOL + _fp16_ulp(sc[t, ref_ids[t]])
        assert margin <= band, f'{label} step {t}: cuda chose {cuda_tokens[t]}, oracle {ref_ids[t]}, margin {margin} > band {band}'
    print(f'[{label}] ids differing from the oracle argmax inside the near-tie band: {len(mism)} of {T}')
    return len(mism)


@pytest.fixture(scope='module', params=['five_exchange', 'tensor_parallel'])
def wide_setup(request):
    """both decode-layer variants of the kernel (er_debug_set decode_fuse)"""
    from edgerunner_b200.engine import Engine
    from oracle.er_oracle import Oracle
    opt = replace(config_defaults['ArAE'], generate_mode='greedy', num_layers=2)
    sd = synth.synth_state_dict(opt, seed=5, eos_logit=-30.0)
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=16000, debug={'decode_fuse': int(request.param == 'tensor_parallel')})
    eng.load_state_dict(sd)
    return opt, sd, eng, Oracle(opt, sd, mode='ledger'), synth.synth_point_cloud(2, opt.point_num)


def test_free_running_16k_every_position_against_oracle(wide_setup):
    opt, sd, eng, orc, cond = wide_setup
    T = 16000
    eng.encode_cond(cond[0].cuda(), 4000)
    eng.prefill([1])
    out = eng.decode(T, mode='greedy', want_logits=True)
    toks = [int(x) for x in out['tokens']]
    assert len(toks) == T
    ce = orc.encode_cond(cond, 4000)[0]
    orc.reset_cache(ce.shape[0] + 1 + T + 1)
    pre0 = orc.prefill(ce, [opt.bos_token_id])
    rep = orc.replay_steps(toks[:T - 1])
    ref = torch.cat([pre0, rep], dim=0)
    n = check_stream(toks, out['logits_pre'].cpu(), ref, eng.V, 'free-running 16k, 2 layers, ArAE width')
    assert n <= T // 200


def test_decode_from_long_resume_prompt(wide_setup):
    """cache built by er_prefill from BOS + 5 999 resume ids (8 049 rows; the dense workspace grows on demand), then 16 decode steps"""
    opt, sd, eng, orc, cond = wide_setup
    resume = grammar_stream(5999, seed=7)
    forced = grammar_stream(16, seed=8)                      # the FSM restarts at idx 0 after a resume (models.py:252): BOM first
    eng.encode_cond(cond[0].cuda(), 4000)
    eng.prefill([1] + resume)
    assert eng.lib.er_cache_rows(eng.h) == opt.num_cond_tokens + 6000
    out = eng.decode(16, mode='greedy', forced=forced, want_logits=True)
    ref = orc.generate(cond, 4000, max_new_tokens=16, generate_mode='greedy', resume_ids=resume, forced_tokens=forced)
    d = (out['logits_pre'].cpu() - ref['logits_pre']).abs()
    print(f'[long resume prompt, L=8049..] max |dlogit| {float(d.max()):.3e} mean {float(d.mean()):.3e}')
    assert float(d.max()) <= MAX_TOL and float(d.mean()) <= MEAN_TOL
    sc = ref['scores'].numpy()
    for t, (a, b) in enumerate(zip(out['tokens'], ref['tokens'])):
        if a != b:
            assert sc[t, b] - sc[t, a] <= 2 * MAX_TOL + _fp16_ulp(sc[t, b]), (t, a, b)


def test_point_latent_generate():
    """cond_mode='point_latent' (the DiT hand-over, infer_dit.py:111-113): latents in, same decode path"""
    from edgerunner_b200.engine import Engine
    from oracle.er_oracle import Oracle
    opt = synth.tiny_options(cond_mode='point_latent')
    sd = synth.synth_state_dict(opt, seed=1, eos_logit=-30.0)
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=128)
    eng.load_state_dict(sd)
    orc = Oracle(opt, sd, mode='ledger')
    lat = torch.randn(1, opt.point_latent_size, opt.point_latent_dim, generator=torch.Generator().manual_seed(3))
    emb, _ = eng.encode_cond(lat[0].cuda(), 2000, want_embeds=True)
    ref_emb = orc.encode_cond(lat, 2000)[0]
    assert (emb.cpu() - ref_emb).abs().max().item() < 2e-2
    eng.prefill([1])
    out = eng.decode(96, mode='greedy', want_logits=True)
    ref = orc.generate(lat, 2000, max_new_tokens=96, generate_mode='greedy', forced_tokens=list(out['tokens']))
    d = (out['logits_pre'].cpu() - ref['logits_pre']).abs()
    assert float(d.max()) <= 2.5e-3 and float(d.mean()) <= 3e-4, (float(d.max()), float(d.mean()))
    # and through the public API
    from core.models import LMM
    from core.utils import get_tokenizer
    model = LMM(opt)
    model.load_state_dict(sd, strict=True)
    model = model.half().eval().to('cuda:0')
    tok, _ = get_tokenizer(opt)
    meshes, toks = model.generate(lat.cuda(), num_faces=2000, max_new_tokens=96, tokenizer=tok, clean=True)
    np.testing.assert_array_equal(toks[0][:len(out['tokens'])], out['tokens'][:len(toks[0])])


def test_arae_sample_c3_shape():
    """BASELINE configs[2] shape on one replica: full ArAE, test_num_face=2000, sample top-k 10.  Sampled ids are grammar-valid,
    reproducible per seed, differ across seeds, each lies in the oracle's top-10 set, and the logits follow the ledger oracle."""
    from edgerunner_b200.engine import Engine
    from oracle.er_oracle import Oracle, ConstraintFSM
    opt = replace(config_defaults['ArAE'], generate_mode='sample')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=8000)
    eng.load_state_dict(sd)
    cond = synth.synth_point_cloud(4, opt.point_num)
    runs = []
    for seed in (21, 21, 22):
        eng.encode_cond(cond[0].cuda(), 2000)
        eng.prefill([1])
        runs.append(eng.decode(8000 if seed == 22 else 48, mode='sample', top_k=10, seed=seed, want_logits=(seed == 21)))
    np.testing.assert_array_equal(runs[0]['tokens'], runs[1]['tokens'])
    assert not np.array_equal(runs[0]['tokens'], runs[2]['tokens'][:48])
    long_toks = runs[2]['tokens']
    assert len(long_toks) == 8000
    fsm, gen = ConstraintFSM(eng.V), []
    for t in long_toks:
        assert int(t) in fsm.allowed(gen)
        gen.append(int(t))
    toks = runs[0]['tokens']
    orc = Oracle(opt, sd, mode='ledger')
    ref = orc.generate(cond, 2000, max_new_tokens=len(toks), generate_mode='greedy', forced_tokens=list(toks))
    d = (runs[0]['logits_pre'].cpu() - ref['logits_pre']).abs()
    print(f'[C3 sample, ArAE] max |dlogit| {float(d.max()):.3e} mean {float(d.mean()):.3e}')
    assert float(d.max()) <= MAX_TOL and float(d.mean()) <= MEAN_TOL
    sc = ref['scores'].numpy()
    for t, tok in enumerate(toks):
        kth = np.sort(sc[t])[::-1][9]
        assert np.isfinite(sc[t, tok])
        if np.isfinite(kth):
            assert sc[t, tok] >= kth - (2 * MAX_TOL + _fp16_ulp(kth)), (t, tok)
