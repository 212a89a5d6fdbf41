"""GPU: one engine, tokens/s of a T-token greedy ArAE decode for a sweep of the runtime decode-kernel knobs (er_debug_set keys that do not
change the weight packing).  Usage: sweep_decode.py [T=6000]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine

T = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
opt = replace(config_defaults['ArAE'], generate_mode='greedy')
eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=T)
eng.load_state_dict(synth.synth_state_dict(opt, seed=0, eos_logit=-30.0, dtype=torch.float16))
cond = synth.synth_point_cloud(0, opt.point_num)[0].cuda()
DEFAULTS = {'pf_dist': 128 * 1024, 'poll_rounds': 4, 'split_handicap_fuse': 0}


def run():
    best = 1e9
    for rep in range(2):
        eng.encode_cond(cond, 4000); eng.prefill([1])
        torch.cuda.synchronize(); t0 = time.time()
        eng.decode(T, sync=False)
        torch.cuda.synchronize(); best = min(best, time.time() - t0)
    return T / best


eng.encode_cond(cond, 4000); eng.prefill([1]); eng.decode(64)
res = {'T': T, 'default': run()}
print('default', res['default'], flush=True)
for key, vals in (('pf_dist', (65536, 98304, 196608, 262144)), ('poll_rounds', (1, 2, 8, 16)), ('split_handicap_fuse', (1, 2, 3))):
    for v in vals:
        eng.debug_set(key, v)
        res[f'{key}={v}'] = run()
        print(key, v, res[f'{key}={v}'], flush=True)
    eng.debug_set(key, DEFAULTS[key])
res['default_again'] = run()
print(json.dumps(res))
