"""GPU: tokens/s of one greedy ArAE decode of T tokens.  Usage: decode_speed.py T[,T2...] [label] [key=value ...]   (er_debug_set switches,
e.g. decode_fuse=1 pf_dist=131072 split_handicap=4)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine

lens = [int(x) for x in sys.argv[1].split(',')]
label = sys.argv[2] if len(sys.argv) > 2 else ''
dbg = dict((kv.split('=')[0], int(kv.split('=')[1])) for kv in sys.argv[3:])
opt = replace(config_defaults['ArAE'], generate_mode='greedy')
eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=max(lens), debug=dbg)
eng.load_state_dict(synth.synth_state_dict(opt, seed=0, eos_logit=-30.0))
cond = synth.synth_point_cloud(0, opt.point_num)[0].cuda()
for T in lens:
    best = 1e9
    for rep in range(2 if T <= 4000 else 1):
        eng.encode_cond(cond, 4000); eng.prefill([1])
        if rep == 0:
            eng.decode(64); eng.encode_cond(cond, 4000); eng.prefill([1])
        torch.cuda.synchronize(); t0 = time.time()
        eng.decode(T, sync=False)
        torch.cuda.synchronize(); best = min(best, time.time() - t0)
    print(f'{label} T={T}: {T / best:.1f} tok/s', flush=True)
