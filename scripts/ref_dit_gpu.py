"""GPU: the REFERENCE's own DiT module (oracle/_ref/py/core/transformer/dit.py, copied by `make -C oracle refpy`) run as infer_dit.py runs it
— .half(), torch.autocast(fp16), the installed flash-attn — next to this repository's CUDA DiT engine, same synthetic weights, same inputs.
Also: the reference's MDiT.run loop body (guidance + a restated diffusers DDIM step, oracle/dit_oracle.py) driven with the reference module
as the denoiser, against the engine's device-side loop; and the time of both per denoiser forward.
Usage: ref_dit_gpu.py [layers=24] [out=gpurun_out/ref_dit_gpu.json] [batch=2]"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch

from oracle import ref_runner as rr
from oracle import dit_oracle as do


def main():
    NL = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    out_path = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/ref_dit_gpu.json'
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    DRY = not torch.cuda.is_available()
    dev = torch.device('cpu' if DRY else 'cuda:0')
    use_flash = (not DRY) and rr.flash_usable(dev)
    rr.setup(mask_flash=not use_flash)                          # reference `core` first on sys.path, kiui / trimesh stubs
    from core.transformer.dit import DiT                        # the reference's module
    cfg = dict(hidden_dim=128, num_heads=2, latent_size=40, latent_dim=16, num_layers=2) if DRY else \
        dict(hidden_dim=1024, num_heads=16, latent_size=2048, latent_dim=64, num_layers=NL)
    M = 9 if DRY else 257
    sd = do.synth_dit_state(**cfg, seed=1)
    ref = DiT(**cfg, gradient_checkpointing=False).eval()
    ref.load_state_dict(sd, strict=True)
    ref = (ref if DRY else ref.half()).to(dev)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, cfg['latent_size'], cfg['latent_dim'], generator=g).to(dev)
    c = torch.randn(B, M, cfg['hidden_dim'], generator=g).to(dev)
    t = torch.tensor([991.0, 501.0, 11.0, 1.0][:B] if B <= 4 else [float(1 + 10 * i) for i in range(B)]).to(dev)
    ac = (lambda: torch.autocast('cpu', enabled=False)) if DRY else (lambda: torch.autocast('cuda', dtype=torch.float16))
    res = {'flash_attn_used': use_flash, 'cfg': cfg, 'batch': B, 'gpu': 'none (dry run)' if DRY else torch.cuda.get_device_name(0)}

    # dtype ledger of the reference through forward hooks
    ledger = {}

    def hook(name):
        def fn(mod, inp, out):
            i = inp[0] if isinstance(inp, tuple) and len(inp) else inp
            if torch.is_tensor(i) and torch.is_tensor(out):
                ledger.setdefault(f'{type(mod).__name__}:{name.split(".")[-1]}', set()).add(f'{str(i.dtype)[6:]}->{str(out.dtype)[6:]}')
        return fn
    hs = [m.register_forward_hook(hook(n)) for n, m in ref.named_modules() if len(list(m.children())) == 0]
    with torch.no_grad(), ac():
        y_ref = ref(x, c, t)
    for h in hs:
        h.remove()
    res['dtype_ledger'] = {k: sorted(v) for k, v in ledger.items()}
    res['ref_out_dtype'] = str(y_ref.dtype)
    y_ref = y_ref.float()

    # oracle (ledger mode) on the same device
    orc = do.DitOracle(sd, cfg['num_heads'], mode='fp32' if DRY else 'ledger', device=dev)
    y_orc = orc.forward(x, c, t)
    res['oracle_vs_ref'] = {'max': float((y_orc - y_ref).abs().max()), 'mean': float((y_orc - y_ref).abs().mean())}
    res['ref_abs_mean'] = float(y_ref.abs().mean())
    if DRY:
        print(json.dumps(res)[:600])
        return

    from edgerunner_b200.dit_engine import DiTEngine
    eng = DiTEngine(dev, cfg['hidden_dim'], cfg['num_heads'], cfg['num_layers'], cfg['latent_size'], cfg['latent_dim'], M, 1280)
    full = {'dit.' + k: v for k, v in sd.items()}
    C = cfg['hidden_dim']
    full.update({'proj_cond.weight': torch.zeros(C, 1280), 'proj_cond.bias': torch.zeros(C), 'norm_cond.weight': torch.ones(C), 'norm_cond.bias': torch.zeros(C)})
    eng.load_state_dict(full)
    y = eng.forward(x, c, t).float()
    res['engine_vs_ref'] = {'max': float((y - y_ref).abs().max()), 'mean': float((y - y_ref).abs().mean())}
    res['engine_vs_oracle'] = {'max': float((y - y_orc).abs().max()), 'mean': float((y - y_orc).abs().mean())}

    # sampling loop: reference module as denoiser + restated scheduler step vs the engine's device loop (R = 1, guided)
    S = 8
    ts, coef = do.ddim_tables(S)
    lat0 = torch.randn(1, cfg['latent_size'], cfg['latent_dim'], generator=g).to(dev)
    cond = c[:1]
    lat = lat0.clone()
    cc = torch.cat([torch.zeros_like(cond), cond], dim=0)
    with torch.no_grad(), ac():
        for i, tt in enumerate(ts.tolist()):
            t_in = torch.tensor([tt] * 2, device=dev, dtype=lat.dtype)
            pred = ref(torch.cat([lat] * 2, dim=0), cc, t_in)
            u, cnd = pred.chunk(2)
            m = u + 7.5 * (cnd - u)
            assert m.dtype == torch.float16
            lat = do.ddim_step(m.float(), lat, coef[i], 'v_prediction', ledger=True)
    lat_eng = eng.run(cond, lat0.clone(), ts.astype(np.float32), coef.numpy(), 7.5, True, 'v_prediction')
    torch.cuda.synchronize()
    res['loop_engine_vs_ref'] = {'steps': S, 'max': float((lat_eng - lat).abs().max()), 'mean': float((lat_eng - lat).abs().mean()), 'lat_abs_mean': float(lat.abs().mean())}

    # time per guided denoiser forward (batch 2): reference module vs engine loop
    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n
    with torch.no_grad(), ac():
        t_in = torch.tensor([991.0] * 2, device=dev)
        xin = torch.cat([lat0] * 2, dim=0)
        res['ref_ms_per_forward_b2'] = 1e3 * timed(lambda: ref(xin, cc, t_in), 5)
    ts20, coef20 = do.ddim_tables(20)
    res['engine_ms_per_step_b2'] = 1e3 * timed(lambda: eng.run(cond, lat0.clone(), ts20.astype(np.float32), coef20.numpy(), 7.5, True, 'v_prediction'), 2) / 20
    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    json.dump(res, open(out_path, 'w'), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != 'dtype_ledger'}))


if __name__ == '__main__':
    main()
