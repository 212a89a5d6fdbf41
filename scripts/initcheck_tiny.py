"""GPU: tiny encode + prefill + decode + teacher-forced forward, meant to run under `compute-sanitizer --tool initcheck`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine
opt = synth.tiny_options()
sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=64, max_points=opt.point_num, max_tf_rows=2 * (opt.num_cond_tokens + 64))
eng.load_state_dict(sd)
cond = synth.synth_point_cloud(0, opt.point_num)
eng.encode_cond(cond[0].cuda(), 1000); eng.prefill([1])
print('prefill done', flush=True)
r = eng.decode(int(sys.argv[1]) if len(sys.argv) > 1 else 40, mode='greedy', want_logits=True)
print('decode done', r['tokens'][:8], flush=True)
