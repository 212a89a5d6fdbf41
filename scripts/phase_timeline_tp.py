"""GPU: per-phase timeline of the TENSOR-PARALLEL decode layer (er_debug_set decode_fuse=1), one CTA, one token, %globaltimer stamps of the
PROF kernel instantiation -> table + json.  Usage: phase_timeline_tp.py [tokens=200,8000] [out.json] [key=value ...]
Segments (us, mean over layers 1..23):  LN2 = poll of the y2 all-reduce + residual + LayerNorm | qkv = P1 GEMV | pub = publish q/k/v words |
kvapp = cache append | qpoll = wait for the head's q words | kv = K and V passes | merge = publish partial + poll S partials + fold |
oproj = row-parallel out_proj (tensor cores) | red1 = y1 atomics issue | LN1 = poll of the y1 all-reduce + residual + LayerNorm | fc1 | fc2 |
red2 = y2 atomics issue."""
import ctypes as C
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


def main():
    tokens = [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['200', '8000'])]
    out_path = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/phase_timeline_tp.json'
    dbg = {'decode_fuse': 1}
    dbg.update(dict((kv.split('=')[0], int(kv.split('=')[1])) for kv in sys.argv[3:]))
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=max(tokens) + 8, debug=dbg)
    eng.load_state_dict(sd); del sd
    cond = synth.synth_point_cloud(0, opt.point_num)
    NL = opt.num_layers
    nslots = 1 + NL * 16 + 2
    # (name, from slot, to slot); slot 16 = slot 0 of the next layer
    segs = [('qkv', 0, 1), ('pub', 1, 2), ('kvapp', 2, 3), ('qpoll', 3, 5), ('kv', 5, 6), ('merge', 6, 15), ('oproj', 15, 12), ('red1', 12, 4),
            ('LN1', 7, 9), ('fc1', 9, 11), ('fc2', 11, 13), ('red2', 13, 10), ('LN2', 14, 16)]
    res = {}
    for tok in tokens:
        for cta in (0, 8, 75, 147):
            eng.encode_cond(cond[0].cuda(), 4000); eng.prefill([1])
            eng.lib.er_debug_phase_timeline(eng.h, tok, cta)
            eng.decode(tok + 4, mode='greedy')
            buf = (C.c_uint64 * 4096)()
            eng.lib.er_debug_read_timeline(eng.h, buf, 4096)
            ts = np.array(list(buf)[:nslots + 16], dtype=np.float64)
            acc = {n: [] for n, _, _ in segs}
            for l in range(1, NL - 1):
                pb = 1 + 16 * l
                for n, a, b in segs:
                    if ts[pb + a] > 0 and ts[pb + b] > 0:
                        acc[n].append((ts[pb + b] - ts[pb + a]) / 1e3)
            layer_us = float(np.mean([(ts[1 + 16 * (l + 1)] - ts[1 + 16 * l]) / 1e3 for l in range(1, NL - 1)]))
            key = f'token{tok}_cta{cta}'
            res[key] = {'L': 2050 + tok, 'layer_us': layer_us, 'token_us': (ts[2 + 16 * NL] - ts[0]) / 1e3,
                        'segments_us_mean': {n: round(float(np.mean(v)), 2) if v else None for n, v in acc.items()}}
            print(key, json.dumps(res[key]), flush=True)
    eng.lib.er_debug_phase_timeline(eng.h, -1, 0)
    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    json.dump(res, open(out_path, 'w'), indent=1)


if __name__ == '__main__':
    main()
