"""Data-parallel TRAINING step on 2 GPUs over NCCL (flattened gradient all-reduce; edgerunner_b200.train.FlatTrainer); skipped on a 1-GPU
box.  Named to sort after every other GPU test file.  The same path at the BASELINE configs[3] shape: `bench.py --workload train` under torchrun
(profiles/r02_bench_train_c4_n2.json: 801 ms per step at N = 2 against 799 at N = 1)."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from test_gpu_dist import _batch, _free_port

pytestmark = pytest.mark.gpu

def _make_trainer(opt, sd, dev):
    from core.models import LMM
    from edgerunner_b200.train import FlatTrainer
    model = LMM(opt); model.load_state_dict(sd, strict=True); model = model.to(dev).train()
    model.config.dropout = 0.0                      # no dropout: the mask's index space depends on how the batch is split over ranks
    return FlatTrainer(model, total_steps=100, max_batch=4, max_tokens=48, lr=1e-3, warmup_ratio=0.0)


def _train_worker(rank, ws, port, q, path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=ws, device_id=torch.device('cuda', rank))
    from edgerunner_b200 import synth
    opt = synth.tiny_options(nof_dropout_ratio=0.0)
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    tr = _make_trainer(opt, sd, f'cuda:{rank}')
    data = _batch(opt, synth.vocab_size_of(opt), 4, 48)
    shard = {k: (v[rank::ws].to(f'cuda:{rank}') if k == 'conds' else v[rank::ws]) for k, v in data.items()}
    out = tr.step(shard)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save(tr.grad.cpu(), path)
    q.put((rank, float(out['loss']), float(tr.grad.double().sum()), float(tr.param.double().sum()), float(tr.param16.double().sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_dp_training_step_two_ranks(tmp_path):
    """2 ranks x 2 samples: after the NCCL all-reduce every rank holds the SAME averaged gradient / updated weights (bit for bit), and that gradient is
    the single-process gradient of the 4-sample batch (equal token counts per rank, so mean of per-rank means = the batch mean) up to fp16 noise"""
    from edgerunner_b200 import synth
    opt = synth.tiny_options(nof_dropout_ratio=0.0)
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    tr = _make_trainer(opt, sd, 'cuda:0')
    data = _batch(opt, synth.vocab_size_of(opt), 4, 48)
    data['conds'] = data['conds'].cuda()
    tr.step(data)
    single = tr.grad.cpu()
    ws, port, path = 2, _free_port(), str(tmp_path / 'grad_rank0.pt')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, ws, port, q, path)) for r in range(ws)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(ws))
    [p.join(timeout=60) for p in procs]
    assert res[0][2:] == res[1][2:], res                                  # identical gradient and weight checksums on both ranks
    dp = torch.load(path)
    err = float((dp - single).norm() / single.norm())
    assert err < 5e-3, err
