"""GPU: per-phase timeline of the persistent decode kernel (one CTA, one token) -> profiles/*.json + table."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine

def main():
    tokens = [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['200', '8000'])]
    out_path = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/phase_timeline.json'
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    T = max(tokens) + 8
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=T)
    eng.load_state_dict(sd); del sd
    cond = synth.synth_point_cloud(0, opt.point_num)
    res = {}
    NL = opt.num_layers
    nslots = 1 + NL * 10 + 2
    names = ['qkv', 'attn', 'out_proj', 'ln1+fc1', 'fc2']
    for tok in tokens:
        for cta in (0, 73, 147):
            eng.encode_cond(cond[0].cuda(), 4000); eng.prefill([1])
            eng.lib.er_debug_phase_timeline(eng.h, tok, cta)
            eng.decode(tok + 4, mode='greedy')
            buf = (C.c_uint64 * nslots)()
            eng.lib.er_debug_read_timeline(eng.h, buf, nslots)
            ts = np.array(list(buf), dtype=np.float64)
            d = np.diff(ts) / 1e3   # us
            # d[0] = sample+embed -> end of P1(layer0) is slot1.. ; per layer: phase i end = slot 1+10l+2i, barrier end = +1
            per = {n: [] for n in names}; bar = {n: [] for n in names}
            for l in range(NL):
                for i, n in enumerate(names):
                    a = 1 + 10 * l + 2 * i
                    start = ts[a - 1]
                    per[n].append((ts[a] - start) / 1e3); bar[n].append((ts[a + 1] - ts[a]) / 1e3)
            key = f'token{tok}_cta{cta}'
            res[key] = {'L': 2050 + tok, 'total_us': (ts[-1] - ts[0]) / 1e3,
                        'phase_us_mean': {n: float(np.mean(per[n][1:])) for n in names},
                        'barrier_wait_us_mean': {n: float(np.mean(bar[n][1:])) for n in names},
                        'lm_head_us': (ts[-2] - ts[-3]) / 1e3, 'lm_barrier_us': (ts[-1] - ts[-2]) / 1e3,
                        'layer0_first_phase_us(incl sample+embed)': per['qkv'][0]}
            print(key, json.dumps(res[key]), flush=True)
    eng.lib.er_debug_phase_timeline(eng.h, -1, 0)
    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    json.dump(res, open(out_path, 'w'), indent=1)

if __name__ == '__main__':
    main()
