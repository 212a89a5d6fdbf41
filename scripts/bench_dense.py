"""GPU: the dense (tensor-core) path at BASELINE configs[3] shape — ArAE teacher-forced forward, seq_len 8192 (+ 2049-token condition prefix
+ BOS/EOS = 10 243 rows), batch 4 per GPU — through er_forward_tf; plus the 2 050-row generate prefill.  Reports ms, algorithmic TFLOP/s
(2 x 680 752 128 per row through the decoder + causal attention 2 N^2 C per layer per sample + encoder) against
MEASURED_PEAKS.json bf16_tflops_sustained.  Usage: bench_dense.py [B=4] [T=8194] [reps=3]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from core.options import config_defaults
from edgerunner_b200 import synth
from edgerunner_b200.engine import Engine


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8194
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    P, C, NL = opt.num_cond_tokens, opt.hidden_dim, opt.num_layers
    N = P + T
    eng = Engine(opt, torch.device('cuda:0'), max_new_tokens=64, max_tf_rows=B * N)
    eng.load_state_dict(synth.synth_state_dict(opt, seed=0, eos_logit=-30.0))
    conds = torch.cat([synth.synth_point_cloud(b, opt.point_num) for b in range(B)]).cuda()
    g = torch.Generator().manual_seed(0)
    tokens = torch.randint(6, 518, (B, T), generator=g)
    tokens[:, 0] = opt.bos_token_id
    labels = torch.cat([torch.full((B, P), -100, dtype=torch.long), tokens.long()], dim=1)
    nf = [4000] * B
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    eng.forward_tf(conds, tokens, labels, nf, opt.kl_weight)
    torch.cuda.synchronize()
    l0 = eng.kernel_launches()
    ev[0].record()
    for _ in range(reps):
        losses, _ = eng.forward_tf(conds, tokens, labels, nf, opt.kl_weight)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    gemm = 2 * 680_752_128 * B * N
    attn = 2 * N * N * C * NL * B             # causal half of 4 N^2 C
    enc = 0.16e12 * B
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) if os.path.exists('MEASURED_PEAKS.json') else {}
    peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
    tf = (gemm + attn + enc) / (ms * 1e-3) / 1e12
    out = {'config': f'ArAE teacher-forced forward B={B} N={N} (P={P} + T={T})', 'ms': ms, 'loss': float(losses[0]), 'tflops': tf, 'frac_of_sustained_peak': tf / peak,
           'peak_tflops': peak, 'gemm_tflop': gemm / 1e12, 'attn_tflop': attn / 1e12, 'launches_per_forward': (eng.kernel_launches() - l0) / reps}
    print(json.dumps(out), flush=True)
    # prefill of a generate request
    eng2 = eng
    eng2.encode_cond(conds[0], 4000)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(5):
        eng2.prefill([1])
    ev[1].record()
    torch.cuda.synchronize()
    pms = ev[0].elapsed_time(ev[1]) / 5
    n = P + 1
    pf = (2 * 680_752_128 * n + 2 * n * n * C * NL) / (pms * 1e-3) / 1e12
    print(json.dumps({'config': f'generate prefill {n} rows', 'ms': pms, 'tflops': pf}), flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/bench_dense.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
