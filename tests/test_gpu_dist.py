"""Data-parallel teacher-forced forward on 2 GPUs over NCCL (BASELINE configs[3]; run with `gpurun --gpus 2`; skipped on a 1-GPU box):
each rank runs er_forward_tf2 on its half of the batch, ONE NCCL all-reduce of {ce_sum, n_tokens, kl} (edgerunner_b200.dist.dp_reduce_losses,
through LMM.forward), and the result equals the single-process loss of the whole batch BIT FOR BIT (the sums travel as fp64; per-token losses
are fp32, so the fp64 partial sums are exact and their order does not matter)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _batch(opt, V, B, T):
    from edgerunner_b200 import synth
    P = opt.num_cond_tokens
    g = torch.Generator().manual_seed(11)
    tokens = torch.randint(6, V, (B, T), generator=g)
    tokens[:, 0] = opt.bos_token_id
    labels = torch.cat([torch.full((B, P), -100, dtype=torch.long), tokens.long()], dim=1)
    conds = torch.cat([synth.synth_point_cloud(20 + b, opt.point_num) for b in range(B)])
    return {'conds': conds, 'tokens': tokens, 'labels': labels, 'masks': torch.ones((B, P + T), dtype=torch.bool),
            'num_faces': torch.tensor([1000, 3000, 500, 7000][:B]), 'num_tokens': torch.full((B,), T)}


def _worker(rank, ws, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=ws, device_id=torch.device('cuda', rank))
    from core.models import LMM
    from edgerunner_b200 import synth
    opt = synth.tiny_options()
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    model = LMM(opt); model.load_state_dict(sd, strict=True); model = model.half().eval().to(f'cuda:{rank}')
    data = _batch(opt, model.vocab_size, 4, 48)
    shard = {k: v[rank::ws].to(f'cuda:{rank}') if k in ('conds',) else v[rank::ws] for k, v in data.items()}
    out = model(shard)
    q.put((rank, float(out['loss']), float(out['loss_ce']), float(out['loss_kl'])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_dp_forward_nccl_two_ranks_bit_exact():
    from core.models import LMM
    from edgerunner_b200 import synth
    opt = synth.tiny_options()
    sd = synth.synth_state_dict(opt, seed=0, eos_logit=-30.0)
    model = LMM(opt); model.load_state_dict(sd, strict=True); model = model.half().eval().to('cuda:0')
    data = _batch(opt, model.vocab_size, 4, 48)
    data['conds'] = data['conds'].cuda()
    single = model(data)
    ws, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(ws))
    [p.join(timeout=60) for p in procs]
    for r, loss, ce, kl in res:
        assert ce == float(single['loss_ce']), (r, ce, float(single['loss_ce']))
        assert kl == float(single['loss_kl']), (r, kl, float(single['loss_kl']))
        assert loss == float(single['loss']), (r, loss, float(single['loss']))

