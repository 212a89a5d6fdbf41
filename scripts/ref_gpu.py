"""GPU: run the REFERENCE's own GPU path (oracle/_ref/py/core copied by `make -C oracle refpy`: model.half() + autocast(fp16) +
flash-attn, infer.py:56,104-106; HF loop restated in oracle/ref_runner.py) next to this repository's CUDA engine on the same B200,
same synthetic weights, same point cloud.  Delivers:
  (a) forward-hook dtype ledger of the reference (one prefill + one cached step)  -> pins SURVEY Appendix B / oracle mode='ledger'
  (b) teacher-forced |dlogit| (ours vs reference, on the reference's greedy stream), id mismatches, FIRST DIVERGENCE index of the
      two free-running greedy streams
  (c) tokens/s of the reference GPU path: the free run, and cached-step windows at L ~ 2k / 8k / 16k (+ 16k-request extrapolation)
Usage: ref_gpu.py [T=4000] [out=gpurun_out/ref_gpu.json]"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from dataclasses import replace

from oracle import ref_runner as rr


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    out_path = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/ref_gpu.json'
    DRY = not torch.cuda.is_available()          # build-container dry run of the reference half (tiny config, CPU)
    dev = torch.device('cpu' if DRY else 'cuda:0')
    use_flash = (not DRY) and rr.flash_usable(dev)
    LMM, cfgs = rr.setup(mask_flash=not use_flash)
    from edgerunner_b200 import synth
    from edgerunner_b200.engine import Engine
    opt = synth.tiny_options() if DRY else replace(cfgs['ArAE'], generate_mode='greedy')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    cond = synth.synth_point_cloud(0, opt.point_num).to(dev)
    model = rr.build_model(opt, sd, dev, half=not DRY)
    sync = (lambda: None) if DRY else torch.cuda.synchronize
    ac = (lambda: torch.autocast('cpu', enabled=False)) if DRY else (lambda: torch.autocast('cuda', dtype=torch.float16))
    V = model.vocab_size
    res = {'flash_attn_used': use_flash, 'T': T, 'torch': torch.__version__, 'gpu': 'none (dry run)' if DRY else torch.cuda.get_device_name(0)}

    # ---- (a) dtype ledger through forward hooks ------------------------------------------------------------------------------
    ledger = {}
    phase = ['prefill']

    def hook(name):
        def fn(mod, inp, out):
            i = inp[0] if isinstance(inp, tuple) and len(inp) else inp
            o = out[0] if isinstance(out, tuple) else out
            key = (phase[0], type(mod).__name__, name.split('.')[-1] if not name.split('.')[-1].isdigit() else name)
            if torch.is_tensor(i) and torch.is_tensor(o):
                ledger.setdefault(str(key), set()).add(f'{str(i.dtype)[6:]}->{str(o.dtype)[6:]}')
        return fn
    hs = [m.register_forward_hook(hook(n)) for n, m in model.named_modules() if len(list(m.children())) == 0]
    with torch.no_grad(), ac():
        emb = rr.prefix_embeds(model, cond, 4000)
        res['inputs_embeds_dtype'] = str(emb.dtype)

        def on_step(n):
            phase[0] = 'decode'
        rec = []
        rr.hf_sample(model.mesh_decoder, emb, opt.eos_token_id, 3, rr.fsm_fn(V, opt.eos_token_id), record_logits=rec, on_step=on_step)
    for h in hs:
        h.remove()
    res['dtype_ledger'] = {k: sorted(v) for k, v in sorted(ledger.items())}
    print('dtype ledger:', json.dumps(res['dtype_ledger'], indent=0)[:3000], flush=True)

    # ---- reference free-running greedy, T tokens, with logits ------------------------------------------------------------------
    rec = []
    marks = {}

    def on_step2(n):
        if n in (1, 257, T):
            sync(); marks[n] = time.perf_counter()
    with torch.no_grad(), ac():
        emb = rr.prefix_embeds(model, cond, 4000)
        sync(); t0 = time.perf_counter()
        ref_tokens, _ = rr.hf_sample(model.mesh_decoder, emb, opt.eos_token_id, T, rr.fsm_fn(V, opt.eos_token_id), record_logits=rec, on_step=on_step2)
        sync(); t1 = time.perf_counter()
    ref_logits = torch.stack(rec)                    # [T, V] fp16 values as float
    res['ref_free_run'] = {'tokens': int(len(ref_tokens)), 'seconds': t1 - t0, 'tok_s': len(ref_tokens) / (t1 - t0),
                           'tok_s_first256_after_prefill': 256 / (marks[257] - marks[1]) if 257 in marks else None}
    print('reference GPU free run:', json.dumps(res['ref_free_run']), flush=True)

    if DRY:
        print('dry run: reference half ok', ref_tokens[:12]); return
    # ---- ours, teacher-forced on the reference's stream ---------------------------------------------------------------------------------
    eng = Engine(opt, dev, max_new_tokens=T + 8)
    eng.load_state_dict(sd)
    del sd
    eng.encode_cond(cond[0], 4000); eng.prefill([1])
    ours = eng.decode(len(ref_tokens), mode='greedy', forced=[int(x) for x in ref_tokens], want_logits=True)
    ol = ours['logits_pre'].cpu()
    ol16 = ol.to(torch.float16).float()
    d = (ol16 - ref_logits).abs()
    mism = np.nonzero(ours['tokens'] != ref_tokens)[0]
    # decision margins of the reference at the mismatching steps (top-1 minus our choice, in the reference's own fp16 logits)
    margins = [float(ref_logits[t, ref_tokens[t]] - ref_logits[t, ours['tokens'][t]]) for t in mism[:50]]
    res['teacher_forced'] = {'max_abs_dlogit': float(d.max()), 'mean_abs_dlogit': float(d.mean()),
                             'p99_abs_dlogit': float(d.flatten().kthvalue(int(0.99 * d.numel())).values),
                             'logit_std': float(ref_logits[ref_logits > -1e4].std()),
                             'id_mismatches': int(len(mism)), 'first_mismatch': int(mism[0]) if len(mism) else -1,
                             'mismatch_margins_ref_fp16': margins,
                             'per_quarter_mean': [float(d[i * len(d) // 4:(i + 1) * len(d) // 4].mean()) for i in range(4)]}
    print('ours vs reference (teacher-forced):', json.dumps(res['teacher_forced']), flush=True)
    # ---- ours, free-running: first divergence -----------------------------------------------------------------------------------------------
    eng.encode_cond(cond[0], 4000); eng.prefill([1])
    free = eng.decode(len(ref_tokens), mode='greedy')['tokens']
    n = min(len(free), len(ref_tokens))
    neq = np.nonzero(free[:n] != ref_tokens[:n])[0]
    res['free_running'] = {'first_divergence': int(neq[0]) if len(neq) else -1, 'compared': int(n)}
    print('free-running first divergence:', json.dumps(res['free_running']), flush=True)
    # run-to-run identity of the REFERENCE itself (cuBLAS/flash are deterministic here?) on the first 300 tokens
    rec2 = []
    with torch.no_grad(), ac():
        emb = rr.prefix_embeds(model, cond, 4000)
        t2, _ = rr.hf_sample(model.mesh_decoder, emb, opt.eos_token_id, 300, rr.fsm_fn(V, opt.eos_token_id), record_logits=rec2)
    res['ref_run_to_run_identical_300'] = bool(torch.equal(torch.stack(rec2), ref_logits[:300]))
    del eng
    torch.cuda.empty_cache()

    # ---- (c) reference GPU tokens/s by context length ------------------------------------------------------------------------------------------
    win = []
    for L in (2050, 8000, 16000):
        s = rr.decode_window(model, L, 48, warm=4)
        win.append((L, s))
        print(f'reference GPU cached step at L={L}: {1 / s:.1f} tok/s', flush=True)
    total, a, b = rr.extrapolate_request(win, 2050, 16000)
    res['ref_gpu_windows'] = {'tok_s': {str(L): 1 / s for L, s in win}, 'extrapolated_16k_request_tok_s': 16000 / total,
                              'model': f't(L) = {a * 1e3:.3f} ms + {b * 1e6:.4f} us * L'}
    print(json.dumps(res['ref_gpu_windows']), flush=True)
    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    json.dump(res, open(out_path, 'w'), indent=1)


if __name__ == '__main__':
    main()
