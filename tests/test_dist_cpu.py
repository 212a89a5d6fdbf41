"""N > 1 host logic on CPU: world_size-2 gloo processes (request sharding, token gather, the data-parallel loss all-reduce)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    from edgerunner_b200 import dist as erd
    mine = erd.shard_requests(5)
    toks = [np.arange(10 * i, 10 * i + 3 + i) for i in mine]           # a fake "generation" per request
    allt = erd.gather_tokens(toks)
    slow = erd.max_over_ranks(1.0 + rank)
    # data-parallel loss: each rank holds a shard of per-token CE values
    rng = np.random.RandomState(0)
    ce_all = rng.rand(37).astype(np.float32) * 3
    kl_all = rng.rand(ws).astype(np.float32)
    shard = ce_all[rank::ws]
    loss, lce, lkl = erd.dp_reduce_losses(torch.tensor(shard.sum()), torch.tensor(float(len(shard))), torch.tensor(kl_all[rank]), 1e-3)
    # flattened gradient buffer: sliced, reverse-order, averaged
    fg = erd.FlatGradAllReduce(1003, 'cpu', n_slices=4)
    va, vb = fg.views([(10, 50), (503,)])
    va.fill_(float(rank + 1)); vb.copy_(torch.arange(503, dtype=torch.float32) * (rank + 1))
    flat = fg.launch().wait()
    ok_flat = bool(torch.all(flat[:500] == 1.5)) and bool(torch.allclose(flat[500:], torch.arange(503, dtype=torch.float32) * 1.5)) \
        and fg.bounds[0][1] == 1003 and fg.bounds[-1][0] == 0 and all(a[0] == b[1] for a, b in zip(fg.bounds[:-1], fg.bounds[1:]))
    assert ok_flat
    q.put((rank, mine, None if allt is None else [a.tolist() for a in allt], slow, float(loss), float(lce), float(lkl),
           float(ce_all.mean()), float(kl_all.sum())))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ws, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(ws))
    [p.join(timeout=60) for p in procs]
    (r0, mine0, all0, slow0, loss0, lce0, lkl0, ce_mean, kl_sum), (r1, mine1, all1, slow1, loss1, lce1, lkl1, _, _) = res
    assert mine0 == [0, 2, 4] and mine1 == [1, 3]
    assert all1 is None and all0 == [list(range(10 * i, 10 * i + 3 + i)) for i in range(5)]
    assert slow0 == slow1 == 2.0
    np.testing.assert_allclose([lce0, lce1], ce_mean, rtol=1e-6)
    np.testing.assert_allclose([lkl0, lkl1], kl_sum, rtol=1e-6)
    np.testing.assert_allclose(loss0, ce_mean + 1e-3 * kl_sum, rtol=1e-6)
    assert loss0 == loss1


def test_single_process_fallbacks():
    from edgerunner_b200 import dist as erd
    assert erd.shard_requests(3) == [0, 1, 2]
    assert erd.max_over_ranks(3.5) == 3.5
    l, ce, kl = erd.dp_reduce_losses(torch.tensor(6.0), torch.tensor(3.0), torch.tensor(0.5), 2.0)
    assert float(ce) == 2.0 and float(l) == 3.0
