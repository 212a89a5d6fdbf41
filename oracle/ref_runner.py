"""REFERENCE RUNNER — TEST INFRASTRUCTURE ONLY (tests/, bench.py's reference legs, scripts/).  Never imported by the product.

Executes the UNMODIFIED reference modules copied by ``make -C oracle refpy`` into the git-ignored ``oracle/_ref/py/core``
(``core.models.LMM``, ``core.transformer.*``) — on the GPU exactly as ``infer.py:56,104-106`` runs them (``model.half().eval()``,
``torch.autocast('cuda', fp16)``, flash-attn kernels when the installed flash_attn supports the device) or on the host cores
(fp32, ``flash_attn`` masked so that ``core/transformer/attention.py:19-25`` takes its naive path).

Only one thing is restated: HF ``GenerationMixin._sample`` (transformers==4.46.2, un-vendored; the installed 5.5 hands
``prepare_inputs_for_generation`` a DynamicCache that ``modeling_opt.py:524`` cannot index).  ``hf_sample`` below follows the 4.46.2
loop (``logits[:, -1].float()`` -> prefix-constraint mask (core/utils.py:143-158) -> [top-k 10 keep-ties -> softmax -> multinomial] |
argmax -> append -> stop at EOS / max_new_tokens) and calls the reference's own ``prepare_inputs_for_generation`` + ``forward``.

Import-time stubs exist only for packages that are absent from the image and never touched by the arithmetic (kiui, trimesh, megfile).
"""

from __future__ import annotations

import math
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_PY = os.path.join(HERE, '_ref', 'py')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_PY, 'core'))


def install_stubs(mask_flash: bool):
    if mask_flash:
        sys.modules['flash_attn'] = None            # attention.py:19-25 -> naive bmm path
    kiui = types.ModuleType('kiui')
    kiui.lo = lambda *a, **k: None
    kiui.seed_everything = lambda s: (torch.manual_seed(s), np.random.seed(s))
    mu = types.ModuleType('kiui.mesh_utils')
    mu.clean_mesh = mu.decimate_mesh = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    op = types.ModuleType('kiui.op')
    op.recenter = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    kiui.mesh_utils, kiui.op = mu, op
    sys.modules.update({'kiui': kiui, 'kiui.mesh_utils': mu, 'kiui.op': op})
    tm = types.ModuleType('trimesh')

    class Trimesh:      # holder only (the clean-up calls are third-party trimesh: parity unpinned)
        def __init__(self, vertices=None, faces=None, **k):
            self.vertices, self.faces = np.asarray(vertices), np.asarray(faces)

        def merge_vertices(self): pass
        def unique_faces(self): return np.ones(len(self.faces), dtype=bool)
        def update_faces(self, m): pass
        def fix_normals(self): pass

    tm.Trimesh = Trimesh
    sys.modules['trimesh'] = tm
    sys.modules['megfile'] = types.ModuleType('megfile')


def import_reference(mask_flash: bool):
    """Put the reference's `core` package first on sys.path (this repository's own `core/` mirror must not shadow it) and import it.
    Must be called before anything imported `core` in this process.  Returns (LMM class, config_defaults)."""
    if not available():
        raise RuntimeError('oracle/_ref/py/core is missing: run `make -C oracle refpy` in the build container')
    assert 'core' not in sys.modules or os.path.abspath(sys.modules['core'].__file__).startswith(REF_PY), \
        'this process already imported the repository\'s own core/ package'
    install_stubs(mask_flash)
    sys.path[:] = [REF_PY] + [p for p in sys.path if os.path.abspath(p or '.') not in (REF_PY,)]
    from core.models import LMM
    from core.options import config_defaults
    return LMM, config_defaults


def flash_usable(device) -> bool:
    """Does the installed flash_attn run on this GPU?  (the reference falls back to its naive path only on ImportError)"""
    try:
        from flash_attn import flash_attn_func
        q = torch.randn(1, 4, 2, 64, device=device, dtype=torch.float16)
        flash_attn_func(q, q, q, 0.0, causal=True)
        torch.cuda.synchronize()
        return True
    except Exception as e:      # noqa
        print('[ref_runner] flash_attn not usable on this device: %s: %s' % (type(e).__name__, str(e)[:200]), flush=True)
        return False


def build_model(opt, sd, device, half: bool):
    """LMM(opt) with the given state dict, as infer.py:41-56 builds it (setup() must have run)."""
    LMM, _ = _REF
    torch.manual_seed(0)
    model = LMM(opt).eval()
    model.load_state_dict(sd, strict=True)
    if half:
        model = model.half()
    return model.eval().to(device)


_REF = None


def setup(mask_flash: bool):
    """import the reference once per process -> (LMM, config_defaults)"""
    global _REF
    if _REF is None:
        _REF = import_reference(mask_flash)
    return _REF


def fsm_fn(vocab_size, eos):
    """the closure LMM.generate builds (core/models.py:245-271), taken out so that the loop below can be driven with a tokenizer"""
    state = {'counter': 0}

    def fn(batch_id, input_ids):
        idx = input_ids.shape[0]
        if idx == 0:
            return [5]
        last = int(input_ids[-1])
        if last == 5:
            state['counter'] = 9
        elif last in (3, 4):
            state['counter'] = 3
        elif last >= 6:
            state['counter'] -= 1
        if state['counter'] > 0:
            return list(range(6, vocab_size))
        return [3, 4, 5, eos]
    return fn


@torch.no_grad()
def hf_sample(decoder, inputs_embeds, eos_token_id, max_new_tokens, prefix_allowed_tokens_fn, do_sample=False, top_k=10,
              forced=None, record_logits=None, past=None, on_step=None):
    """transformers==4.46.2 GenerationMixin._sample for B == 1, num_beams == 1, inputs_embeds-only prompt (see module docstring).
    forced: teacher-forcing stream (the model's own choice is still what is returned).  record_logits: list receiving logits[0,-1]."""
    dev = inputs_embeds.device
    input_ids = torch.ones((1, 0), dtype=torch.long, device=dev)      # what the FSM / the next step sees
    chosen = []
    attention_mask = torch.ones(inputs_embeds.shape[:2], dtype=torch.long, device=dev)
    while True:
        mi = decoder.prepare_inputs_for_generation(input_ids, past_key_values=past, attention_mask=attention_mask,
                                                   inputs_embeds=inputs_embeds, use_cache=True)
        out = decoder(**mi, return_dict=True)
        past = out.past_key_values
        logits = out.logits[:, -1, :].clone().float()
        mask = torch.full_like(logits, -math.inf)
        allowed = prefix_allowed_tokens_fn(0, input_ids[0])
        mask[0, allowed] = 0
        scores = logits + mask
        if do_sample:
            k = min(top_k, scores.size(-1))
            remove = scores < torch.topk(scores, k)[0][..., -1, None]
            scores = scores.masked_fill(remove, -float('inf'))
            probs = torch.softmax(scores, dim=-1)
            nxt = torch.multinomial(probs, num_samples=1).squeeze(1)
        else:
            nxt = torch.argmax(scores, dim=-1)
        if record_logits is not None:
            record_logits.append(out.logits[0, -1].detach().float().cpu())
        chosen.append(int(nxt))
        fed = nxt if forced is None else torch.tensor([int(forced[len(chosen) - 1])], device=dev)
        input_ids = torch.cat([input_ids, fed[:, None]], dim=-1)
        attention_mask = torch.cat([attention_mask, attention_mask.new_ones((1, 1))], dim=-1)
        if on_step is not None:
            on_step(len(chosen))
        if int(fed) == eos_token_id or input_ids.shape[1] >= max_new_tokens:
            return np.asarray(chosen, dtype=np.int64), past


@torch.no_grad()
def prefix_embeds(model, cond, num_faces):
    """LMM.generate lines 219-231: cond embeds ++ BOS embed"""
    dev = cond.device
    nf = torch.full((1,), num_faces, dtype=torch.long, device=dev)
    ce = model.encode_cond(cond, nf)['cond_embeds']
    bos = torch.full((1, 1), model.opt.bos_token_id, dtype=torch.long, device=dev)
    return torch.cat((ce, model.mesh_decoder.model.embd(bos)), dim=1)


@torch.no_grad()
def make_past(model, L, randn=True):
    """A FABRICATED cache of L rows in the model's dtype (timing only): the tuple-of-tuples layout ShapeOPTDecoder returns
    (modeling_opt.py:400-426), 24 x (K, V) each [1, H, L, D]."""
    p = next(model.mesh_decoder.parameters())
    H, D, NL = model.opt.num_heads, model.opt.hidden_dim // model.opt.num_heads, model.opt.num_layers
    mk = (lambda: torch.randn(1, H, L, D, device=p.device, dtype=p.dtype) * 0.1) if randn else (lambda: torch.zeros(1, H, L, D, device=p.device, dtype=p.dtype))
    return tuple((mk(), mk()) for _ in range(NL))


@torch.no_grad()
def decode_window(model, L, n_steps, warm=2, autocast=True, past=None):
    """Time n_steps cached decode steps of the reference decoder from a cache of L rows (fabricated unless `past` is given) -> seconds
    per token.  One step = ShapeOPT.forward with past_key_values (modeling_opt.py:464-517: 24 x [q/k/v Linear, torch.cat of the whole
    cache :191-192, attention(), out_proj, LN, fc1, ReLU, fc2, LN], lm_head) + the argmax and its host sync, i.e. what HF's loop runs
    per generated token."""
    dec = model.mesh_decoder
    p = next(dec.parameters())
    dev = p.device
    if past is None:
        past = make_past(model, L)
    ids = torch.full((1, 1), 100, dtype=torch.long, device=dev)
    am = torch.ones((1, L + 1), dtype=torch.long, device=dev)

    def step(past, am):
        if dev.type == 'cuda' and autocast:
            with torch.autocast('cuda', dtype=torch.float16):
                out = dec(input_ids=ids, past_key_values=past, attention_mask=am, use_cache=True, return_dict=True)
        else:
            out = dec(input_ids=ids, past_key_values=past, attention_mask=am, use_cache=True, return_dict=True)
        nxt = torch.argmax(out.logits[:, -1, :].float(), dim=-1)      # the per-token host sync HF's loop has (stopping criteria)
        int(nxt)
        return out.past_key_values, torch.cat([am, am.new_ones((1, 1))], dim=-1)

    for _ in range(warm):
        past, am = step(past, am)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        past, am = step(past, am)
    if dev.type == 'cuda':
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n_steps


def extrapolate_request(samples, L0, T):
    """samples: [(L, seconds_per_token)] at >= 2 context lengths -> seconds for T tokens generated from a cache of L0 rows, with the
    per-token time modelled as a + b * L (weights + KV bytes) fitted by least squares.  Clearly an extrapolation."""
    Ls = np.array([s[0] for s in samples], dtype=np.float64)
    ts = np.array([s[1] for s in samples], dtype=np.float64)
    if len(samples) == 1:
        a, b = ts[0], 0.0
    else:
        b, a = np.polyfit(Ls, ts, 1)
    n = T - 1
    total = n * a + b * (n * L0 + n * (n - 1) / 2.0)
    return float(total), float(a), float(b)
