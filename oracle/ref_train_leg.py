"""REFERENCE LEG of `bench.py --workload train` — TEST / MEASUREMENT INFRASTRUCTURE (executes oracle/_ref/py, the reference's own modules).

    python oracle/ref_train_leg.py cpu [--steps K --warmup W --tokens T --threads t]   the reference's training step on the host cores (fp32, naive attention)
    python oracle/ref_train_leg.py gpu [--steps K --warmup W --batch B]                the reference's training step on the GPU as main.py runs it

One step = what main.py:160-181 does for a batch: ``model.train(); out = model(data); out['loss'].backward(); clip_grad_norm_(1.0);
AdamW(lr, wd 0.01, betas (0.9, 0.95)).step()`` on the ArAE preset (point encoder trained, dropout 0.1, opt.checkpointing as the preset says).

* gpu: ``torch.autocast('cuda', bf16)`` (acc_configs/gpu8.yaml), flash-attn when the installed build supports the device, at the BASELINE configs[3]
  shape (B samples of 8 194 tokens + 2 049 condition rows).  The denominator for "times the reference's GPU path" of the training workload.
* cpu: the same step in fp32 with the naive attention path on a BOUNDED SAMPLE — one sample of `--tokens` tokens (+ 2 049 condition rows) — whose
  time is extrapolated to the configs[3] step by the ratio of model FLOPs (GEMMs linear in the rows, causal attention quadratic).

Prints ONE line: REF_TRAIN_LEG {json}.  Own process: the reference's package is also called `core`."""
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

C4_B, C4_T = 4, 8194


def model_flops(B, N, C, NL, per_row):
    """3 x GEMM + 3.5 x causal attention of the forward (the bench line's definition) for B samples of N rows"""
    return 3 * 2 * per_row * B * N + 3.5 * 2 * N * N * C * NL * B


def batch(opt, V, B, T, dev, seed=0):
    from edgerunner_b200 import synth
    g = torch.Generator().manual_seed(100 + seed)
    tokens = torch.randint(6, V, (B, T), generator=g)
    tokens[:, 0] = opt.bos_token_id
    P = opt.num_cond_tokens
    return {'conds': torch.cat([synth.synth_point_cloud(b, opt.point_num) for b in range(B)]).to(dev), 'tokens': tokens.to(dev),
            'labels': torch.cat([torch.full((B, P), -100, dtype=torch.long), tokens.long()], dim=1).to(dev),
            'masks': torch.ones((B, P + T), dtype=torch.bool, device=dev), 'num_faces': torch.tensor([4000] * B, device=dev),
            'num_tokens': torch.full((B,), T, device=dev)}


def run(kind, steps, warmup, tokens, bsz, threads=None, tiny=False):
    from dataclasses import replace
    gpu = kind == 'gpu'
    dev = torch.device('cuda:0' if gpu else 'cpu')
    use_flash = gpu and rr.flash_usable(dev)
    LMM, cfgs = rr.setup(mask_flash=not use_flash)
    from edgerunner_b200 import synth
    opt = synth.tiny_options() if tiny else replace(cfgs['ArAE'])
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    torch.manual_seed(0)
    model = LMM(opt)
    model.load_state_dict(sd, strict=True)
    del sd
    model = model.to(dev).train()
    if not gpu:
        torch.set_num_threads(threads or (os.cpu_count() or 1))
    B = bsz if gpu else 1
    T = (C4_T if gpu else tokens) if not tiny else 40
    data = batch(opt, model.vocab_size, B, T, dev)
    optim = torch.optim.AdamW(model.parameters(), lr=opt.lr, weight_decay=0.01, betas=(0.9, 0.95))
    times, losses = [], []
    for i in range(warmup + steps):
        if gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        optim.zero_grad()
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=gpu):
            out = model(data)
        out['loss'].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), opt.gradient_clip)
        optim.step()
        loss = float(out['loss'])
        if gpu:
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
            losses.append(loss)
    s_step = sum(times) / len(times)
    P, C, NL = opt.num_cond_tokens, opt.hidden_dim, opt.num_layers
    per_row = sum(p.numel() for n, p in model.named_parameters() if n.startswith('mesh_decoder.model.layers.') and p.dim() == 2) + model.vocab_size * C
    f_sample, f_c4 = model_flops(B, P + T, C, NL, per_row), model_flops(C4_B, P + C4_T, C, NL, per_row)
    s_c4 = s_step * f_c4 / f_sample
    return {'kind': 'reference', 's_per_step_sample': s_step, 'sample_rows': B * (P + T), 'extrapolated_c4_step_s': s_c4,
            'tok_s': C4_B * C4_T / s_c4, 'losses': losses, 'steps': steps, 'warmup': warmup,
            'threads': None if gpu else torch.get_num_threads(), 'host_threads': os.cpu_count(), 'flash_attn': bool(use_flash),
            'path': ('gpu: torch autograd, autocast(bf16), %s, opt.checkpointing=%s' % ('flash_attn' if use_flash else 'naive attention', opt.checkpointing)) if gpu
            else 'cpu: torch autograd, fp32, naive attention (flash_attn masked), opt.checkpointing=%s' % opt.checkpointing,
            'sample': f'{steps} full training step(s) (forward, backward, clip, AdamW) of the reference LMM on {B} sample(s) of {T} tokens + {P} condition rows'
                      + ('' if gpu and T == C4_T and B == C4_B else f'; extrapolated to {C4_B} x {C4_T} tokens by model FLOPs (x{f_c4 / f_sample:.1f})')}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('kind', choices=['cpu', 'gpu'])
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--tokens', type=int, default=128)
    ap.add_argument('--batch', type=int, default=C4_B)
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--tiny', action='store_true')
    a = ap.parse_args()
    r = run(a.kind, a.steps, a.warmup, a.tokens, a.batch, a.threads or None, a.tiny)
    print('REF_TRAIN_LEG ' + json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
