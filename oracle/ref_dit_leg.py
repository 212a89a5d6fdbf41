"""TEST / BENCH INFRASTRUCTURE.  Times the REFERENCE's own DiT module (oracle/_ref/py/core/transformer/dit.py) at the preset size for
bench.py's `--workload dit` legs, in its own process (its package is called `core`, like this repository's mirror):
  gpu: .half() + autocast(fp16) + flash-attn on cuda:0, guided batch (2 x images)   -> ms per denoiser forward
  cpu: fp32, naive attention, torch CPU ops on the host cores (bounded: `--images 1` = batch 2) -> s per denoiser forward
Falls back to the oracle port (oracle/dit_oracle.py, fp32) when oracle/_ref/py is absent.  Prints one line `REF_DIT_LEG {json}`."""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('kind', choices=['cpu', 'gpu'])
    ap.add_argument('--images', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--layers', type=int, default=24)
    ap.add_argument('--threads', type=int, default=0)
    args = ap.parse_args()
    from oracle import ref_runner as rr
    from oracle import dit_oracle as do
    cfg = dict(hidden_dim=1024, num_heads=16, latent_size=2048, latent_dim=64, num_layers=args.layers)
    M, B = 257, 2 * args.images
    dev = torch.device('cuda:0' if args.kind == 'gpu' else 'cpu')
    sd = do.synth_dit_state(**cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cfg['latent_size'], cfg['latent_dim'], generator=g).to(dev)
    c = torch.randn(B, M, cfg['hidden_dim'], generator=g).to(dev)
    t = torch.full((B,), 991.0).to(dev)
    out = {'kind': args.kind, 'batch': B, 'layers': args.layers}
    if rr.available():
        use_flash = args.kind == 'gpu' and rr.flash_usable(dev)
        rr.setup(mask_flash=not use_flash)
        from core.transformer.dit import DiT
        m = DiT(**cfg, gradient_checkpointing=False).eval()
        m.load_state_dict(sd, strict=True)
        m = (m.half() if args.kind == 'gpu' else m).to(dev)
        fwd = lambda: m(x, c, t)
        out.update(impl='reference', flash_attn=use_flash)
    else:
        o = do.DitOracle(sd, cfg['num_heads'], mode='fp32', device=dev)
        fwd = lambda: o.forward(x, c, t)
        out.update(impl='port')
    if args.kind == 'cpu':
        nt = args.threads or min(os.cpu_count() or 1, 64)
        torch.set_num_threads(nt)
        out['cores'] = nt
        ctx = torch.autocast('cpu', enabled=False)
    else:
        ctx = torch.autocast('cuda', dtype=torch.float16)
    sync = torch.cuda.synchronize if args.kind == 'gpu' else (lambda: None)
    with torch.no_grad(), ctx:
        for _ in range(args.warmup):
            fwd()
        sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd()
        sync()
    out['s_per_forward'] = (time.perf_counter() - t0) / args.steps
    print('REF_DIT_LEG ' + json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
