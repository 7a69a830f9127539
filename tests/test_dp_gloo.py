"""Data-parallel host logic on CPU: world_size 2 over gloo, with the oracle standing in for the CUDA encoders
(the trainer is model-agnostic; the CUDA path's own loss-scaling hook is covered by tests/test_gpu_dp.py).

Checks SURVEY.md 8(e): G ranks == the mean of G single-process reference runs on the shards ('shard_mean'),
and the whole-batch loss with the global link normaliser ('global_norm'), each with ONE all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(seed=0, B=6, N=24, D=5, H=8, C=3):
    from helpers import synth_batch
    from oracle import diffpool_oracle as orc
    x, adj, nb, label = synth_batch(seed, B, N, D, 3, N, C, density=0.25)
    torch.manual_seed(seed)
    model = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25).double()
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    return orc, model, t(x), t(adj), nb, torch.tensor(label)


def _expected(mode, world, clip):
    """Single-process restatement: run the oracle on every shard, combine gradients as the mode says."""
    from graph_pooling_b200 import dp
    orc, model, x, adj, nb, label = _make()
    params = list(model.parameters())
    acc = [torch.zeros_like(p) for p in params]
    g64 = nb.astype(np.int64)
    entries_global = float(np.sum(g64 * g64))
    for r in range(world):
        sh = dp.shard_batch(r, world, x, adj, nb, label, assign_x=x)
        model.zero_grad()
        yp = model(sh['x'], sh['adj'], sh['nb'], assign_x=sh['assign_x'])
        total = model.loss(yp, sh['label'], sh['adj'], sh['nb'])
        if mode == 'global_norm':
            link = model.link_loss
            l64 = sh['nb'].astype(np.int64)
            total = (total - link) * (len(sh['nb']) / len(nb)) + link * (float(np.sum(l64 * l64)) / entries_global)
        total.backward()
        for a, p in zip(acc, params):
            a += p.grad
    if mode == 'shard_mean':
        acc = [a / world for a in acc]
    flat = torch.cat([a.reshape(-1) for a in acc])
    norm = flat.norm(2)
    coef = torch.clamp(clip / (norm + 1e-6), max=1.0)
    return flat * coef, norm


def _worker(rank, world, port, mode, q):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        torch.set_num_threads(1)
        from graph_pooling_b200 import dp
        orc, model, x, adj, nb, label = _make()
        if rank == 1:                                   # parameters must come from rank 0's broadcast
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        tr = dp.DataParallelTrainer(model, optimizer=None, clip=0.5, mode=mode)
        tr.broadcast_parameters(0)
        sh = dp.shard_batch(rank, world, x, adj, nb, label, assign_x=x)
        tr.step(sh['x'], sh['adj'], sh['nb'], sh['label'], assign_x=sh['assign_x'], global_num_nodes=nb,
                global_batch=len(nb))
        want, _ = _expected(mode, world, 0.5)
        got = tr.grads.flat
        err = float((got - want).norm() / want.norm())
        # every parameter's .grad is a view of the flat buffer
        off, views_ok = 0, True
        for p in tr.grads.params:
            views_ok &= p.grad.data_ptr() == got.data_ptr() + off * got.element_size()
            off += p.numel()
        q.put((rank, err, views_ok, None))
        dist.destroy_process_group()
    except Exception as e:                              # pragma: no cover
        q.put((rank, None, False, repr(e)))


@pytest.mark.parametrize('mode', ['shard_mean', 'global_norm'])
def test_two_ranks_match_single_process_restatement(mode):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, views_ok, exc in res:
        assert exc is None, 'rank %d raised %s' % (rank, exc)
        assert views_ok
        assert err < 1e-10, 'rank %d gradient mismatch %.3e' % (rank, err)


def test_shard_bounds_cover_and_balance():
    from graph_pooling_b200 import dp
    for n in (0, 1, 7, 20, 256):
        for w in (1, 2, 3, 8):
            b = [dp.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
