"""GPU: where does poisoned (never written) device memory leak into results?  NaN counts after each stage (tiny or arae)."""
import os, sys
os.environ.setdefault('ER_POISON_ALLOC', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine
over = {}
for a in sys.argv[1:]:
    k, v = a.split('='); over[k] = int(v)
opt = synth.tiny_options(**over)
print('options', over, flush=True)
sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=320, max_points=opt.point_num, max_tf_rows=2 * (opt.num_cond_tokens + 64))
eng.load_state_dict(sd)
cond = synth.synth_point_cloud(0, opt.point_num)
emb, lat = eng.encode_cond(cond[0].cuda(), 1000, want_embeds=True, want_latents=True)
torch.cuda.synchronize()
print('latents NaN', int(torch.isnan(lat.float()).sum()), 'of', lat.numel(), '| cond embeds NaN', int(torch.isnan(emb).sum()), 'of', emb.numel(),
      'rows with NaN', torch.isnan(emb).any(dim=1).nonzero().flatten().tolist()[:10], flush=True)
eng.prefill([1])
T = 300
forced = [5] + [6 + (7 * i) % 100 for i in range(T - 1)]
r = eng.decode(T, mode='greedy', want_logits=True, forced=forced)
lg = r['logits_pre']
bad = [i for i in range(lg.shape[0]) if bool(torch.isnan(lg[i]).any())]
print('decode steps', lg.shape[0], 'steps with NaN logits:', bad[:20], '... count', len(bad), flush=True)
eng.encode_cond(cond[0].cuda(), 1000); eng.prefill([1])
r2 = eng.decode(T, mode='greedy', want_logits=True, forced=forced)
lg2 = r2['logits_pre']
bad2 = [i for i in range(lg2.shape[0]) if bool(torch.isnan(lg2[i]).any())]
diff = (lg != lg2).any(dim=1).nonzero().flatten().tolist()
print('second run NaN steps', bad2[:20], 'steps differing from first run', diff[:20], flush=True)
V = synth.vocab_size_of(opt)
toks = torch.randint(3, V, (2, 40)); toks[:, 0] = 1
labels = torch.full((2, opt.num_cond_tokens + 40), -100, dtype=torch.int64); labels[:, opt.num_cond_tokens + 1:] = toks[:, 1:]
losses, logits = eng.forward_tf(torch.cat([cond, cond]), toks, labels, [1000, 1000], 0.0, want_logits=True)
print('forward_tf logits NaN', int(torch.isnan(logits).sum()), 'of', logits.numel(), 'losses', losses, flush=True)
