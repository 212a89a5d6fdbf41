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
    dbg = dict((kv.split('=')[0], int(kv.split('=')[1])) for kv in sys.argv[3:])       # er_debug_set switches: key=value ...
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    T = max(tokens) + 8
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=T, debug=dbg)
    eng.load_state_dict(sd); del sd
    cond = synth.synth_point_cloud(0, opt.point_num)
    res = {}
    NL = opt.num_layers
    nslots = 1 + NL * 16 + 2
    segs = ['res+LN2', 'qkv_gemv', 'qkv_epi', 'B1', 'attn', 'B2', 'combine', 'out_proj', 'B3', 'res+LN1', 'fc1', 'B4', 'h1_load', 'fc2', 'B5']
    for tok in tokens:
        for cta in (0, 147):
            eng.encode_cond(cond[0].cuda(), 4000); eng.prefill([1])
            eng.lib.er_debug_phase_timeline(eng.h, tok, cta)
            eng.decode(tok + 4, mode='greedy')
            buf = (C.c_uint64 * 4096)()
            eng.lib.er_debug_read_timeline(eng.h, buf, 4096)
            ts = np.array(list(buf)[:nslots], dtype=np.float64)
            att = np.array(list(buf)[2048:2048 + 148], dtype=np.float64) / 1e3
            last = np.array(list(buf)[2048 + 256:2048 + 256 + 148])
            print('attention phase per CTA (us), layer 5: min %.2f median %.2f max %.2f; by split:' % (att[:144].min(), np.median(att[:144]), att[:144].max()), [round(float(att[s::9][:16].mean()), 2) for s in range(9)], 'last-arrivers mean %.2f others %.2f' % (att[:144][last[:144] == 1].mean(), att[:144][last[:144] == 0].mean()), flush=True)
            ahead = np.array(list(buf)[2048:2048 + nslots], dtype=np.float64)
            acc = {n: [] for n in segs}
            for l in range(1, NL):          # skip layer 0 (its first segment includes sample + embed)
                pb = 1 + 16 * l
                prev = ts[pb - 2]           # B5 of the previous layer (slot pb-16+14)
                for i, n in enumerate(segs):
                    acc[n].append((ts[pb + i] - prev) / 1e3); prev = ts[pb + i]
            waits = [ts[1 + 16 * l + 15] / 1e3 for l in range(1, NL)]
            ah = {n: round(float(np.mean([ahead[1 + 16 * l + i] - ahead[1 + 16 * l + i - 1] for l in range(1, NL)])), 1) for i, n in enumerate(segs)}
            key = f'token{tok}_cta{cta}'
            res[key] = {'L': 2050 + tok, 'ring_wait_us_per_layer(thread0)': float(np.mean(waits)), 'failed_try_waits_thread0_per_segment': ah, 'token_us': (ts[2 + 16 * NL] - ts[0]) / 1e3, 'layer_us': float(sum(np.mean(acc[n]) for n in segs)),
                        'segments_us_mean': {n: round(float(np.mean(acc[n])), 2) for n in segs},
                        'lm_head_us': (ts[1 + 16 * NL] - ts[16 * NL - 1]) / 1e3, 'lm_barrier_us': (ts[2 + 16 * NL] - ts[1 + 16 * NL]) / 1e3}
            print(key, json.dumps(res[key]), flush=True)
    eng.lib.er_debug_phase_timeline(eng.h, -1, 0)
    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    json.dump(res, open(out_path, 'w'), indent=1)

if __name__ == '__main__':
    main()
