"""Helper for test_gpu_parity.py::test_poisoned_memory_and_repeatability — runs in its own process because the allocation poisoning
(`er_debug_set(NULL, "poison_alloc", 1)`) is process-wide.  Every device allocation starts as 0xFF bytes (NaN in fp16 and fp32): a read of
anything the engine did not write first turns the logits into NaN; three decodes from the same state must agree bit for bit."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine
from edgerunner_b200 import _lib


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'tiny'
    if which == 'tiny':
        opt, T = synth.tiny_options(), 300
    else:
        from dataclasses import replace
        from core.options import config_defaults
        opt, T = replace(config_defaults['ArAE'], generate_mode='greedy'), 80
    _lib.check(_lib.load().er_debug_set(None, b'poison_alloc', 1))
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=T + 8, max_points=opt.point_num)
    eng.load_state_dict(sd)
    cond = synth.synth_point_cloud(0, opt.point_num)[0].cuda()
    runs = []
    for rep in range(3):
        eng.encode_cond(cond, 1000); eng.prefill([1])
        r = eng.decode(T, mode='greedy', want_logits=True)
        runs.append((r['tokens'], r['logits_pre'].clone()))
    nan = [int(torch.isnan(l).sum()) for _, l in runs]
    same = [bool(np.array_equal(runs[0][0], t) and torch.equal(runs[0][1], l)) for t, l in runs[1:]]
    first_diff = [int((runs[0][1] != l).any(dim=1).float().argmax()) if not s else -1 for (t, l), s in zip(runs[1:], same)]
    print(json.dumps(dict(config=which, steps=T, nan=nan, identical_to_first=same, first_differing_step=first_diff)))


if __name__ == '__main__':
    main()
