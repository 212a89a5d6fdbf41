"""GPU diagnostic: stage-by-stage error of the CUDA path against the ledger oracle (tiny config, optional full ArAE)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine
from oracle.er_oracle import Oracle
from core.options import config_defaults

def stats(name, a, b):
    d = (a - b).abs()
    print(f'{name:28s} max {d.max().item():.3e} mean {d.mean().item():.3e} frac_neq {(d > 0).float().mean().item():.4f} ref_absmax {b.abs().max().item():.3f} ref_std {b.std().item():.3f}', flush=True)

def run(opt, T, tag, point_latent=False):
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    orc = Oracle(opt, sd, mode='ledger')
    cond = synth.synth_point_cloud(0, opt.point_num)
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=T + 8, max_points=opt.point_num)
    eng.load_state_dict(sd)
    t0 = time.time()
    emb, lat = eng.encode_cond(cond[0].cuda(), 1000, want_embeds=True, want_latents=True)
    torch.cuda.synchronize()
    rl = orc.encode_points(cond)[0]
    re = orc.encode_cond(cond, 1000)[0]
    stats(tag + ' latents', lat.float().cpu(), rl)
    stats(tag + ' cond_embeds', emb.cpu(), re)
    # decoder with IDENTICAL latents on both sides (isolates the decoder): point_latent engine fed the oracle latents
    opt2 = replace(opt, cond_mode='point_latent')
    sd2 = {k: v for k, v in sd.items() if not k.startswith('point_encoder.')}
    eng2 = Engine(opt2, torch.device('cuda:0'), max_new_tokens=T + 8)
    eng2.load_state_dict(sd2)
    orc2 = Oracle(opt2, sd2, mode='ledger')
    emb2, _ = eng2.encode_cond(rl.cuda(), 1000, want_embeds=True)
    re2 = orc2.encode_cond(rl[None], 1000)[0]
    stats(tag + ' cond_embeds(same lat)', emb2.cpu(), re2)
    eng2.prefill([1])
    out = eng2.decode(T, mode='greedy', want_logits=True)
    ref = orc2.generate(rl[None], 1000, max_new_tokens=T, generate_mode='greedy', forced_tokens=list(out['tokens']))
    lp, rp = out['logits_pre'].cpu(), ref['logits_pre']
    stats(tag + ' logits step0 (prefill)', lp[0], rp[0])
    stats(tag + ' logits all steps', lp, rp)
    per = (lp - rp).abs().max(dim=1).values
    print(tag, 'per-step max err first 12:', [f'{x:.1e}' for x in per[:12].tolist()], flush=True)
    print(tag, 'token agreement', float((out['tokens'] == ref['tokens']).mean()), 'tokens[:12]', out['tokens'][:12], flush=True)
    # end-to-end (encoder on both sides)
    eng.prefill([1])
    out = eng.decode(T, mode='greedy', want_logits=True)
    ref = orc.generate(cond, 1000, max_new_tokens=T, generate_mode='greedy', forced_tokens=list(out['tokens']))
    stats(tag + ' e2e logits', out['logits_pre'].cpu(), ref['logits_pre'])
    return out

if __name__ == '__main__':
    run(synth.tiny_options(), 64, 'tiny')
    if len(sys.argv) > 1 and sys.argv[1] == 'arae':
        opt = replace(config_defaults['ArAE'], generate_mode='greedy')
        t0 = time.time()
        out = run(opt, 16, 'arae')
        g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'arae.npz'))
        print('arae vs reference-fp32 golden: tokens equal', (out['tokens'] == g['greedy_tokens'][:16]).mean(),
              'max logit diff', np.abs(out['logits_pre'].cpu().numpy() - g['greedy_logits'][:16]).max(), 'time', time.time() - t0)
