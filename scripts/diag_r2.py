"""GPU (round 2): where does a decode step's time go?  One process, two engines (default / fused phases), every knob set through
er_debug_set: L2 run-ahead distance sweep, barrier-free streaming rate (`nosync`, results garbage), at three context lengths
(`cache_rows` pretends the cache is that long: contents irrelevant for timing).  Prints tokens/s + achieved HBM GB/s per setting and
writes gpurun_out/diag_r2.json.  Usage: diag_r2.py [quick]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine

QUICK = len(sys.argv) > 1 and sys.argv[1] == 'quick'
OUT = 'gpurun_out/diag_r2.json'


def grammar_stream(n, seed=0):
    """a grammar-valid forced stream: BOM + 9 coords, then (L|R) + 3 coords ..."""
    rng = np.random.RandomState(seed)
    out = [5] + list(rng.randint(6, 518, 9))
    while len(out) < n:
        out += [int(rng.randint(3, 5))] + list(rng.randint(6, 518, 3))
    return [int(x) for x in out[:n]]


def timed_decode(eng, cond, T, L0, forced):
    eng.encode_cond(cond, 4000)
    eng.prefill([1])
    if L0 != 2050:
        eng.debug_set('cache_rows', L0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    eng.decode(T, mode='greedy', forced=forced[:T], sync=False)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    W, kv = eng.weight_bytes_per_token(), eng.kv_bytes_per_row()
    n = T - 1
    alg = n * (W + kv) + kv * (n * L0 + n * (n - 1) // 2)
    return dict(tok_s=T / (ms / 1e3), ms=ms, gbs=alg / (ms / 1e3) / 1e9, us_per_layer=ms * 1e3 / n / 24)


def main():
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    cond = synth.synth_point_cloud(0, opt.point_num)[0].cuda()
    Tcap = 18200
    res = []
    forced = grammar_stream(4096)
    lens = [2050, 10000, 17000]
    dists = [0, 128, 256] if not QUICK else [128]
    for name, dbg in (('default', {}), ('fuse', {'decode_fuse': 1})):
        eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=Tcap, debug=dbg)
        eng.load_state_dict(sd)
        timed_decode(eng, cond, 64, 2050, forced)      # warm-up
        for L0 in lens:
            T = 768 if L0 == 2050 else 384
            for nosync in (0, 1):
                for d in dists:
                    if nosync and d != 128:
                        continue
                    eng.debug_set('nosync', nosync)
                    eng.debug_set('pf_dist', d * 1024)
                    r = min((timed_decode(eng, cond, T, L0, forced) for _ in range(2)), key=lambda x: x['ms'])
                    r.update(variant=name, L0=L0, T=T, pf_kb=d, nosync=nosync)
                    res.append(r)
                    print(json.dumps(r), flush=True)
        eng.debug_set('nosync', 0)
        # parity of this variant with pf on vs off (same kernel, same order: must be bit-identical), free-running 300 tokens
        outs = []
        for d in (0, 128 * 1024):
            eng.debug_set('pf_dist', d)
            eng.encode_cond(cond, 4000); eng.prefill([1])
            outs.append(eng.decode(300, mode='greedy', want_logits=True))
        same = bool(torch.equal(outs[0]['logits_pre'], outs[1]['logits_pre']))
        print(f'{name}: run-ahead on/off bit-identical logits over 300 free-running tokens: {same}', flush=True)
        res.append(dict(variant=name, pf_bit_identical=same))
        del eng
        torch.cuda.empty_cache()
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open(OUT, 'w'), indent=1)


if __name__ == '__main__':
    main()
