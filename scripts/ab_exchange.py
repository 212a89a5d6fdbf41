"""GPU: A/B of the decode kernel's cross-CTA exchange — flagged words (ER_DECODE_LL=1, default) vs grid barriers (=0).
Same arithmetic in both modes, so ids AND logits must be bit-identical; then tokens/s of each variant at a few lengths."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine


def make(opt, sd, T, ll):
    os.environ['ER_DECODE_LL'] = str(ll)
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=T)
    eng.load_state_dict(sd)
    return eng


def run(eng, cond, T, want_logits, chunk=0, mode='greedy'):
    eng.encode_cond(cond, 4000); eng.prefill([1])
    torch.cuda.synchronize(); t0 = time.time()
    r = eng.decode(T, mode=mode, want_logits=want_logits, tokens_per_launch=chunk, seed=7)
    torch.cuda.synchronize(); dt = time.time() - t0
    return r, dt


def main():
    out = {}
    lens = [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['2000'])]
    variants = [('barrier', 0, {}), ('ll_all', 1, {'ER_POLL_ROUNDS': '1000000000'}), ('ll_r4', 1, {'ER_POLL_ROUNDS': '4'}),
                ('ll_r1', 1, {'ER_POLL_ROUNDS': '1'}), ('ll_r4_h0', 1, {'ER_POLL_ROUNDS': '4', 'ER_SPLIT_HANDICAP': '0'}),
                ('ll_r4_h2', 1, {'ER_POLL_ROUNDS': '4', 'ER_SPLIT_HANDICAP': '2'}), ('ll_r16', 1, {'ER_POLL_ROUNDS': '16'})]
    for name, opt in (('tiny', synth.tiny_options()), ('arae', replace(config_defaults['ArAE'], generate_mode='greedy'))):
        sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
        cond = synth.synth_point_cloud(0, opt.point_num)[0].cuda()
        T = 400 if name == 'tiny' else max(lens)
        engs = {ll: make(opt, sd, T, ll) for ll in (1, 0)}
        del sd
        n_cmp = 300
        res = {}
        for key, ll, chunk in (('ll_a', 1, 97), ('ll_b', 1, 0), ('bar_a', 0, 0), ('bar_b', 0, 53)):
            res[key] = run(engs[ll], cond, n_cmp, True, chunk=chunk)[0]
        eq = lambda a, b: (bool(np.array_equal(res[a]['tokens'], res[b]['tokens'])), bool(torch.equal(res[a]['logits_pre'], res[b]['logits_pre'])),
                           float((res[a]['logits_pre'] - res[b]['logits_pre']).abs().max()))
        out[name] = {'ll_vs_ll': eq('ll_a', 'll_b'), 'bar_vs_bar': eq('bar_a', 'bar_b'), 'll_vs_bar': eq('ll_a', 'bar_a')}
        print(name, json.dumps(out[name]), flush=True)
        if name == 'arae':
            for T in lens:
                for vname, ll, env in variants:
                    for k in ('ER_POLL_ROUNDS', 'ER_SPLIT_HANDICAP'):
                        os.environ.pop(k, None)
                    os.environ.update(env)
                    run(engs[ll], cond, 64, False)
                    best = min(run(engs[ll], cond, T, False)[1] for _ in range(2 if T <= 4000 else 1))
                    out[f'arae_T{T}_{vname}_tok_s'] = T / best
                    print(f'arae T={T} {vname}: {T / best:.1f} tok/s', flush=True)
        del engs
        torch.cuda.empty_cache()
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/ab_exchange.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
