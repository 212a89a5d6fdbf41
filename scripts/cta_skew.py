"""GPU: per-CTA phase boundaries of one layer (5) of one token -> who is late, by how much (profiles/r01_cta_skew_*.json)."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine

SEG = ['LN2_done', 'qkv_gemv', 'qkv_pub', 'B1', 'attn_done', 'B2', 'attn_loaded', 'out_proj', 'B3', 'LN1_done', 'fc1', 'B4', 'h1_loaded', 'fc2', 'B5']


def main():
    tok = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    out_path = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/cta_skew.json'
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=tok + 16)
    eng.load_state_dict(synth.synth_state_dict(opt, seed=0, eos_logit=-30.0))
    cond = synth.synth_point_cloud(0, opt.point_num)[0].cuda()
    eng.encode_cond(cond, 4000); eng.prefill([1])
    eng.lib.er_debug_phase_timeline(eng.h, tok, 0)
    eng.decode(tok + 4, mode='greedy')
    buf = (C.c_uint64 * 8192)()
    eng.lib.er_debug_read_timeline(eng.h, buf, 8192)
    eng.lib.er_debug_phase_timeline(eng.h, -1, 0)
    a = np.array(list(buf)[4096:4096 + 16 * 148], dtype=np.float64).reshape(148, 16)[:, :15]
    t0 = a[:, 0].min()
    a = (a - t0) / 1e3                                   # us since the first CTA finished LN2 of layer 5
    res = {'token': tok, 'L': 2050 + tok, 'stamps': SEG, 'per_stamp': {}}
    for k, n in enumerate(SEG):
        col = a[:, k]
        order = np.argsort(col)
        res['per_stamp'][n] = dict(min=float(col.min()), median=float(np.median(col)), max=float(col.max()),
                                   p90=float(np.percentile(col, 90)), last5=[int(i) for i in order[-5:]], first5=[int(i) for i in order[:5]])
        print(f'{n:12s} min {col.min():7.2f} med {np.median(col):7.2f} p90 {np.percentile(col, 90):7.2f} max {col.max():7.2f}  last {list(order[-5:])} first {list(order[:3])}', flush=True)
    dur = np.diff(a, axis=1)
    for k in range(14):
        col = dur[:, k]
        print(f'  {SEG[k]:>12s}->{SEG[k + 1]:12s} dur min {col.min():6.2f} med {np.median(col):6.2f} max {col.max():6.2f} (cta {int(col.argmax())})', flush=True)
    res['durations_us'] = {f'{SEG[k]}->{SEG[k + 1]}': dict(min=float(dur[:, k].min()), median=float(np.median(dur[:, k])), max=float(dur[:, k].max()), argmax=int(dur[:, k].argmax())) for k in range(14)}
    res['raw_us'] = a.round(3).tolist()
    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    json.dump(res, open(out_path, 'w'))


if __name__ == '__main__':
    main()
