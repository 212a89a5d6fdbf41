import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine
from oracle.er_oracle import Oracle
opt = synth.tiny_options()
sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=64, max_points=opt.point_num)
eng.load_state_dict(sd)
orc = Oracle(opt, sd, mode='ledger')
cond = synth.synth_point_cloud(0, opt.point_num)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
eng.encode_cond(cond[0].cuda(), 1000); eng.prefill([1])
out = eng.decode(T, mode='greedy', want_logits=True)
ref = orc.generate(cond, 1000, max_new_tokens=T, generate_mode='greedy', forced_tokens=list(out['tokens']))
d = (out['logits_pre'].cpu() - ref['logits_pre']).abs()
print('tokens', out['tokens'], ref['tokens'])
for t in range(T):
    print(t, 'max err', d[t].max().item(), 'nan', int(torch.isnan(out['logits_pre'][t]).sum()))
