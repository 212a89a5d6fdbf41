"""GPU (round 2, first call): validate and time the fused decode phases (er_debug_set decode_fuse=1) against the default kernel.
The fused kernel sums the K-split partials in a different order (exact fixed-point sum of fp32 partials), so logits agree within the
fp16-rounding noise band, not bit for bit; each mode must be bit-reproducible run to run.  Prints tokens/s of both at the given lengths."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine


def make(opt, sd, T, fuse):
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=T, debug={'decode_fuse': fuse})
    eng.load_state_dict(sd)
    return eng


def run(eng, cond, T, want_logits, forced=None):
    eng.encode_cond(cond, 4000); eng.prefill([1])
    torch.cuda.synchronize(); t0 = time.time()
    r = eng.decode(T, mode='greedy', want_logits=want_logits, forced=forced)
    torch.cuda.synchronize()
    return r, time.time() - t0


def main():
    lens = [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['4000'])]
    for name, opt in (('tiny', synth.tiny_options()), ('arae', replace(config_defaults['ArAE'], generate_mode='greedy'))):
        sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
        cond = synth.synth_point_cloud(0, opt.point_num)[0].cuda()
        T = 400 if name == 'tiny' else max(lens)
        engs = {f: make(opt, sd, T, f) for f in (0, 1)}
        del sd
        base = run(engs[0], cond, 200, True)[0]
        forced = [int(x) for x in base['tokens']]                       # teacher-force the default kernel's stream through the fused one
        a = run(engs[1], cond, 200, True, forced=forced)[0]
        b = run(engs[1], cond, 200, True, forced=forced)[0]
        d = (a['logits_pre'] - base['logits_pre']).abs()
        print(name, 'fused vs default: max |dlogit| %.3e mean %.3e nan %d | fused run-to-run identical %s | free-running ids equal %s' % (
            float(d.max()), float(d.mean()), int(torch.isnan(a['logits_pre']).sum()), bool(torch.equal(a['logits_pre'], b['logits_pre'])),
            bool(np.array_equal(run(engs[1], cond, 200, False)[0]['tokens'], base['tokens']))), flush=True)
        if name == 'arae':
            for T in lens:
                for f in (0, 1):
                    run(engs[f], cond, 64, False)
                    best = min(run(engs[f], cond, T, False)[1] for _ in range(2 if T <= 4000 else 1))
                    print(f'arae T={T} fuse={f}: {T / best:.1f} tok/s', flush=True)
        del engs
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
