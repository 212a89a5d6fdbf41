"""Driver for ncu captures of the DiT denoiser: 4 images (denoiser batch 8), preset size, `steps` guided DDIM steps launched directly
(graph replay off, so that every kernel is a plain launch).  Usage: ncu ... python scripts/ncu_dit.py [steps=2] [layers=24]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch

from core.models_dit import DDIMScheduler
from edgerunner_b200 import synth
from edgerunner_b200.dit_engine import DiTEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NL = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device('cuda:0')
cfg = dict(hidden_dim=1024, num_heads=16, latent_size=2048, latent_dim=64, num_layers=NL)
eng = DiTEngine(dev, 1024, 16, NL, 2048, 64, 257, 1280)
eng.load_state_dict(synth.synth_dit_state_dict(**cfg, cond_dim=1280, seed=0))
eng.debug_set('graph', 0)
sched = DDIMScheduler()
sched.set_timesteps(100)
ts = sched.timesteps[:S]
g = torch.Generator().manual_seed(0)
cond = torch.randn(4, 257, 1024, generator=g).to(dev)
lat = torch.randn(4, 2048, 64, generator=g).to(dev)
eng.run(cond, lat, ts.numpy().astype(np.float32), sched.step_coefficients(ts).numpy(), 7.5, True, 'v_prediction')
torch.cuda.synchronize()
print('done', float(lat.abs().mean()))
