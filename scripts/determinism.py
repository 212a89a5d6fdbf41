"""GPU: run-to-run determinism of the tiny configuration (dense path and decode path separately)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine

opt = synth.tiny_options()
sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=400, max_tf_rows=4 * (opt.num_cond_tokens + 130))
eng.load_state_dict(sd)
cond = synth.synth_point_cloud(0, opt.point_num)
rng = np.random.RandomState(0)
V = synth.vocab_size_of(opt)
toks = torch.from_numpy(rng.randint(3, V, size=(2, 96)).astype(np.int64)); toks[:, 0] = 1
labels = torch.full((2, opt.num_cond_tokens + 96), -100, dtype=torch.int64); labels[:, opt.num_cond_tokens + 1:] = toks[:, 1:]
outs = []
for rep in range(3):
    r = eng.forward_tf(torch.cat([cond, cond]), toks, labels, [4000, 4000], 0.0, want_logits=True)
    outs.append(r['logits'].clone() if isinstance(r, dict) else r[-1].clone())
print('forward_tf logits identical across runs:', [bool(torch.equal(outs[0], o)) for o in outs[1:]], 'max diff', float(max((outs[0] - o).abs().max() for o in outs[1:])))
for chunk in (0, 53):
    res = []
    for rep in range(3):
        eng.encode_cond(cond[0].cuda(), 4000); eng.prefill([1])
        res.append(eng.decode(300, want_logits=True, tokens_per_launch=chunk)['logits_pre'].clone())
    print(f'decode chunk={chunk}: identical across runs:', [bool(torch.equal(res[0], o)) for o in res[1:]], 'max diff', float(max((res[0] - o).abs().max() for o in res[1:])),
          'first differing step', [int((res[0] != o).any(dim=1).float().argmax()) for o in res[1:]])
    if chunk == 0:
        base = res[0]
print('chunk 53 vs chunk 0 identical:', bool(torch.equal(base, res[0])), 'first differing step', int((base != res[0]).any(dim=1).float().argmax()))
