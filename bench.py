#!/usr/bin/env python
"""bench.py — mesh-tokens/sec of the auto-regressive decode hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--max-new T]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, replicas)

A "step" is one complete pass of the hot path over one synthetic request of BASELINE.json configs[1]:
ArAE preset, seeded synthetic fp16 weights, one 8192-point cloud, test_num_face=4000, greedy,
max_new_tokens=16000 (seq_len ~ 18k): point encoder -> 2050-row prefill -> 16000-token decode (one persistent
kernel) [-> detokenize in the e2e leg].  Decode shards as independent replicas (generate asserts B == 1,
core/models.py:215): N GPUs = N requests, no collective on the data path; "scaling": "weak".

Output: ONE JSON line on rank 0 (see the task contract): value = whole-job tokens/s with inputs resident in HBM,
e2e = the same through LMM.generate with host buffers (H2D of the cloud and D2H of the ids inside the timed region),
roofline = algorithmic HBM bytes of the decode kernel / its CUDA-event duration vs MEASURED_PEAKS.json,
cpu_baseline = the CPU oracle (a port of the reference algorithm) timed on this box's host cores on a bounded sample.
`--impl reference` times that CPU port alone (the Python reference cannot travel to the GPU box).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'mesh-tokens/sec at 4k-face greedy decode'
UNIT = 'tokens/s'


def workload(args):
    from dataclasses import replace
    from core.options import config_defaults
    from edgerunner_b200 import synth
    if args.tiny:
        opt = synth.tiny_options()
        return opt, 'tiny-debug', min(args.max_new, 400), 1000
    opt = replace(config_defaults['ArAE'], generate_mode='greedy')
    return opt, 'ArAE greedy decode test_num_face=4000 max_new_tokens=%d batch=1 (BASELINE configs[1])' % args.max_new, args.max_new, 4000


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def algorithmic_decode_bytes(eng, L0, T):
    """per generated token bytes(L) = W + kv_row * L (KV read) + kv_row (KV write); the kernel runs T-1
    forward passes for T sampled tokens (the last token needs none), at cache lengths L0 .. L0+T-2."""
    W, kv = eng.weight_bytes_per_token(), eng.kv_bytes_per_row()
    n = T - 1
    return n * (W + kv) + kv * (n * L0 + n * (n - 1) // 2)


def pick_threads(step_fn, candidates):
    """The CPU port is a chain of small GEMVs: more threads is not faster.  Time a few steps per candidate, keep the best."""
    best, best_t = candidates[0], float('inf')
    for n in candidates:
        torch.set_num_threads(n)
        step_fn()
        t0 = time.perf_counter()
        step_fn(); step_fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def thread_candidates():
    n = os.cpu_count() or 1
    return sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n})


def cpu_port_tokens_per_s(opt, sd, seconds_budget, threads):
    """The CPU oracle (fp32 port of the reference's CPU path: same ops, naive attention, fp32) on a bounded sample:
    decode steps from a prefilled 2050-row cache."""
    from edgerunner_b200 import synth
    from oracle.er_oracle import Oracle     # bench.py's cpu_baseline leg is allowed to execute the oracle
    torch.set_num_threads(min(threads, 32))
    orc = Oracle(opt, sd, mode='fp32')
    cond = synth.synth_point_cloud(0, opt.point_num)
    ce = orc.encode_cond(cond, 4000)[0]
    n_max = 256 + 32
    orc.reset_cache(ce.shape[0] + 1 + n_max + 1)
    orc.prefill(ce, [opt.bos_token_id])
    L0 = orc.L
    state = {'tok': 5}

    def one():
        pre = orc.step(state['tok'])
        state['tok'] = 6 + int(torch.argmax(pre[0, 6:]))
    threads = pick_threads(one, thread_candidates())
    L0 = orc.L
    tok, n, t0 = state['tok'], 0, time.perf_counter()
    while n < n_max and (time.perf_counter() - t0 < seconds_budget or n < 8):
        pre = orc.step(tok)
        tok = 6 + int(torch.argmax(pre[0, 6:]))
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, threads, f'{n} greedy decode steps from a prefilled cache at L={L0}..{L0 + n} (fp32, torch CPU ops, {threads} threads); ' \
                   f'prefill/encoder excluded; short-L sample flatters the CPU'


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path, as the oracle port, on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from edgerunner_b200 import synth
    opt, wl, T, nf = workload(args)
    threads = os.cpu_count() or 1
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    from oracle.er_oracle import Oracle
    torch.set_num_threads(min(threads, 32))
    orc = Oracle(opt, sd, mode='fp32')
    del sd
    cond = synth.synth_point_cloud(0, opt.point_num)
    ce = orc.encode_cond(cond, nf)[0]
    per_step = 24 if not args.tiny else 16
    total = (args.steps + args.warmup) * per_step
    orc.reset_cache(ce.shape[0] + 1 + total + 2 + 3 * len(thread_candidates()) + 4)
    orc.prefill(ce, [opt.bos_token_id])
    L0 = orc.L
    tok = 5

    def one_step():
        nonlocal tok
        for _ in range(per_step):
            pre = orc.step(tok)
            tok = 6 + int(torch.argmax(pre[0, 6:]))

    def one_token():
        nonlocal tok
        pre = orc.step(tok)
        tok = 6 + int(torch.argmax(pre[0, 6:]))

    host_threads = threads
    threads = pick_threads(one_token, thread_candidates())
    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = (f'{per_step} greedy decode tokens per step from a prefilled {L0}-row cache (cache grows to {orc.L}); fp32 torch CPU ops, '
              f'{threads} threads (best of {thread_candidates()} on {host_threads} host threads); CPU port of the reference path (oracle/er_oracle.py) — the Python reference cannot travel to this box')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': wl, 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--max-new', type=int, default=16000)
    ap.add_argument('--tokens-per-launch', type=int, default=0)
    ap.add_argument('--tiny', action='store_true', help='debug: tiny model')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()

    if args.impl == 'reference':
        run_reference_arm(args)
        return

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)

    from core.models import LMM
    from core.utils import get_tokenizer
    from edgerunner_b200 import synth

    opt, wl, T, nf = workload(args)
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    model = LMM(opt)
    model.load_state_dict(sd, strict=True)
    model = model.half().eval().to(dev)
    tokenizer, _ = get_tokenizer(opt)
    eng = model.get_engine(max_new_tokens=T)
    cond_host = synth.synth_point_cloud(seed=rank, n=opt.point_num).pin_memory()      # independent request per replica
    cond_dev = cond_host.to(dev)
    L0 = opt.num_cond_tokens + 1

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    dec_ms = []

    def step_device():
        """inputs resident in HBM: encoder + prefill + decode on the device, ids stay on the device"""
        eng.encode_cond(cond_dev[0], nf)
        eng.prefill([opt.bos_token_id])
        ev[2].record()
        out = eng.decode(T, mode='greedy', tokens_per_launch=args.tokens_per_launch, sync=False)
        ev[3].record()
        return out

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    for _ in range(args.warmup):
        out = step_device()
    barrier()
    n_tok = int(out['n'].item())
    launches0 = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        out = step_device()
        torch.cuda.current_stream().synchronize()
        dec_ms.append(ev[2].elapsed_time(ev[3]))
    ev[1].record()
    barrier()
    ms_total = max_over_ranks(ev[0].elapsed_time(ev[1]))
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.kernel_launches() - launches0
    n_tok = int(out['n'].item())
    value = world * args.steps * n_tok / (ms_total / 1e3)
    dec_ms_avg = float(np.mean(dec_ms))

    # ---- e2e: the public call (LMM.generate) with host buffers ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            c = cond_host.to(dev, non_blocking=True)            # H2D of this step's input from pinned memory
            with torch.no_grad():
                meshes, toks = model.generate(c, num_faces=nf, max_new_tokens=T, tokenizer=tokenizer, clean=True)   # D2H of ids inside
            return toks[0]
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            step_e2e()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                toks = step_e2e()
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {'value': world * args.steps * len(toks) / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(cond_host.numel() * 4),
               'd2h_bytes_per_step': int(len(toks) * 4 + 4), 'ms_per_step': dt / args.steps * 1e3,
               'api': 'core.models.LMM.generate(cond, num_faces, max_new_tokens, tokenizer, clean=True) incl. meto detokenize + mesh clean-up'}

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------------------------
    peaks_path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    else:
        peak, peak_src = 6650.0, 'B200_PROFILING.md fallback 6.65 TB/s (of fallback)'
    alg = algorithmic_decode_bytes(eng, L0, n_tok)
    achieved = alg / (dec_ms_avg / 1e3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(REPO, 'profiles', 'decode_traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        # ncu cannot replay a 15 s launch: the captured launch is a short one of the same kernel; its measured
        # DRAM-bytes / algorithmic-bytes ratio is applied to this launch's algorithmic bytes
        traffic = tj['dram_over_algorithmic'] * alg if 'dram_over_algorithmic' in tj else None
        traffic_note = tj.get('note')
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'kernel': 'er::decode_persistent_kernel', 'algorithmic_bytes_per_launch': alg, 'launch_ms': dec_ms_avg,
                'peak_source': peak_src, 'frac_of_nominal_8TBs': achieved / 8000.0,
                'decode_only_tokens_per_s': (n_tok - 1) / (dec_ms_avg / 1e3), 'traffic_note': traffic_note}
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, used, sample = cpu_port_tokens_per_s(opt, sd, 15.0, threads)
        cpu = {'value': v, 'unit': UNIT, 'cores': used, 'kind': 'port', 'sample': sample + f' (thread count picked from {thread_candidates()} of {threads} host threads)'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16',
        'data': 'synthetic',
        'config': {'workload': wl, 'tokens_per_step_per_gpu': n_tok, 'prefix_rows': L0, 'parallelism': f'replicas x{world}',
                   'l2': 'inputs larger than L2: 1.36 GB weights + up to 2.66 GB KV cache streamed per token vs 126 MB L2',
                   'weights': 'seeded synthetic, fp16 (edgerunner_b200.synth)'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
    }
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
