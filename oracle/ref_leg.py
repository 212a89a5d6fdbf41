"""REFERENCE LEGS of bench.py — TEST / MEASUREMENT INFRASTRUCTURE (executes oracle/_ref/py, the reference's own modules).

    python oracle/ref_leg.py cpu [--steps K --warmup W --tokens n --threads t]     the reference's CPU path on the host cores
    python oracle/ref_leg.py gpu [--steps K --warmup W --tokens n]                 the reference's GPU path (fp16 + autocast + flash-attn)

Workload = BASELINE.json configs[1]: ArAE greedy, 16 000 new tokens from a 2 050-row prefix (cache 2 050 .. 18 049 rows).  The whole
request takes the reference minutes (GPU) to hours (CPU), so one "step" is a BOUNDED SAMPLE of it: `--tokens` cached decode steps at
each of three context lengths (2 050, 10 000, 18 000 rows; caches fabricated — contents do not change the work), from which the time of
the full request is EXTRAPOLATED with t(L) = a + b L (weights + KV bytes are linear in L).  Prints ONE JSON line:
{"tok_s": 16000 / extrapolated seconds, "windows": {L: tok/s}, ...}.  Runs as its own process because the reference's package is also
called `core` (this repository's drop-in mirror has the same name by design)."""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
import torch

from oracle import ref_runner as rr

L0, T = 2050, 16000
WINDOWS = (2050, 10000, 18000)


def run(kind, steps, warmup, tokens, threads=None, tiny=False):
    from dataclasses import replace
    gpu = kind == 'gpu'
    dev = torch.device('cuda:0' if gpu else 'cpu')
    use_flash = gpu and rr.flash_usable(dev)
    LMM, cfgs = rr.setup(mask_flash=not use_flash)
    from edgerunner_b200 import synth
    opt = synth.tiny_options() if tiny else replace(cfgs['ArAE'], generate_mode='greedy')
    windows = (80, 200, 400) if tiny else WINDOWS
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    model = rr.build_model(opt, sd, dev, half=gpu)
    del sd
    pasts = {L: rr.make_past(model, L, randn=gpu) for L in windows}
    if not gpu:
        # a chain of small GEMVs + a torch.cat of the whole cache per layer: more threads is not faster (128 threads on this box: 16 s per
        # token; 16 threads: tens of ms).  Time one cached step per candidate at the shortest window and keep the best.
        n_host = os.cpu_count() or 1
        cands = [threads] if threads else sorted({c for c in (8, 16, 32, 64) if c <= n_host} | ({n_host} if n_host <= 16 else set()))
        best, best_t = cands[0], float('inf')
        for c in cands:
            torch.set_num_threads(c)
            t = rr.decode_window(model, windows[0], 2, warm=1, past=pasts[windows[0]])
            if t < best_t:
                best, best_t = c, t
        torch.set_num_threads(best)
    per_step = []
    last = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        win = [(L, rr.decode_window(model, L, tokens, warm=1 if i else 2, past=pasts[L])) for L in windows]
        dt = time.perf_counter() - t0
        if i >= warmup:
            per_step.append((dt, win))
        last = win
    # average seconds/token per window over the timed steps
    avg = [(L, sum(w[j][1] for _, w in per_step) / len(per_step)) for j, L in enumerate(windows)]
    total, a, b = rr.extrapolate_request(avg, 80 if tiny else L0, 300 if tiny else T)
    n_req = 300 if tiny else T
    return {
        'kind': 'reference', 'path': 'gpu: model.half() + autocast(fp16) + %s' % ('flash_attn' if use_flash else 'naive attention (flash_attn not usable here)') if gpu
        else 'cpu: fp32, naive attention (flash_attn masked), torch CPU ops',
        'tok_s': n_req / total, 'extrapolated_request_s': total, 'model': 't(L) = %.4f ms + %.6f us * L' % (a * 1e3, b * 1e6),
        'windows_tok_s': {str(L): 1.0 / s for L, s in avg}, 'tokens_per_window': tokens, 'steps': steps, 'warmup': warmup,
        'sample_s_per_step': sum(d for d, _ in per_step) / len(per_step),
        'threads': None if gpu else torch.get_num_threads(), 'host_threads': os.cpu_count(),
        'sample': f'{tokens} cached decode steps of the reference ShapeOPT.forward (+argmax, host sync) at each of L={list(windows)} '
                  f'(fabricated caches), extrapolated to the {n_req}-token request with t(L)=a+bL',
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('kind', choices=['cpu', 'gpu'])
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--tokens', type=int, default=0)
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--tiny', action='store_true')
    a = ap.parse_args()
    tokens = a.tokens or (32 if a.kind == 'gpu' else 2)
    with torch.no_grad():
        r = run(a.kind, a.steps, a.warmup, tokens, a.threads or None, a.tiny)
    print('REF_LEG ' + json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
