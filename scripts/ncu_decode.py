"""Small driver for ncu: ArAE engine, prefill, then ONE decode launch of N tokens (the kernel ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L_extra = int(sys.argv[2]) if len(sys.argv) > 2 else 0     # decode this many tokens first (un-profiled launch) to grow the cache
opt = replace(config_defaults['ArAE'], generate_mode='greedy')
sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=L_extra + T + 8)
eng.load_state_dict(sd); del sd
cond = synth.synth_point_cloud(0, opt.point_num)
eng.encode_cond(cond[0].cuda(), 4000); eng.prefill([1])
out = eng.decode(L_extra + T, mode='greedy', tokens_per_launch=(L_extra if L_extra else 0) or (L_extra + T))
print('tokens', out['tokens'][-4:])
