"""CUDA side of the data-parallel path (dp.py): the loss-scaling hook of the drop-in encoders (CE weight
shard_size / global_batch, link loss normalised by the GLOBAL sum of n_b^2) and the flat
gradient buffer, on one GPU, against the oracle evaluating the same formula in fp64."""
import numpy as np
import pytest
import torch

from helpers import rel_l2, synth_batch
from oracle import diffpool_oracle as orc

pytestmark = pytest.mark.gpu


def test_global_norm_shard_loss_and_flat_gradients():
    from graph_pooling_b200 import dp, encoders
    B, N, D, H, C = 6, 40, 5, 16, 3
    x, adj, nb, label = synth_batch(4, B, N, D, 3, N, C, density=0.2)
    torch.manual_seed(1)
    mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
    mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
    mc.load_state_dict(mo.state_dict())
    mc, mo = mc.cuda(), mo.double()
    # pretend this process is rank 1 of 2: its shard is graphs 3..5 of a global batch of 6
    sh = dp.shard_batch(1, 2, torch.tensor(x), torch.tensor(adj), nb, torch.tensor(label))
    ce_scale = 3.0 / 6.0
    g64, l64 = nb.astype(np.int64), sh['nb'].astype(np.int64)
    e_glob, e_loc = float(np.sum(g64 * g64)), float(np.sum(l64 * l64))

    xo, ao = sh['x'].double(), sh['adj'].double()
    yo = mo(xo, ao, sh['nb'], assign_x=xo)
    tot = mo.loss(yo, sh['label'], ao, sh['nb'])
    want = (tot - mo.link_loss) * ce_scale + mo.link_loss * (e_loc / e_glob)
    want.backward()

    tr = dp.DataParallelTrainer(mc, optimizer=None, clip=None, mode='global_norm')
    tr.world = 2                                         # exercise the scaling branch without a process group
    xc, ac, lc = sh['x'].cuda(), sh['adj'].cuda(), sh['label'].cuda()
    tr.grads.zero()
    yp = mc(xc, ac, sh['nb'], assign_x=xc)
    loss = tr._loss(yp, lc, ac, sh['nb'], nb, B)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert mc._ce_scale == 1.0 and mc._entries_override is None        # hook restored
    flat_ref = torch.cat([p.grad.reshape(-1) for p in mo.parameters()]).numpy()
    assert rel_l2(tr.grads.flat.cpu().numpy(), flat_ref) < 2e-5
    off = 0
    for p in tr.grads.params:
        assert p.grad.data_ptr() == tr.grads.flat.data_ptr() + off * 4
        off += p.numel()


def test_flat_adam_matches_torch_adam_with_clipping():
    """dp.FlatAdam (gp_sumsq_f32 + gp_adam_step_f32 over flat buffers) == clip_grad_norm_(2.0) + torch.optim.Adam."""
    import copy
    import torch
    from graph_pooling_b200 import dp
    torch.manual_seed(0)
    shapes = [(30, 17), (17,), (5, 5, 3), (1,)]
    pa = [torch.nn.Parameter(torch.randn(s, device='cuda')) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ref = torch.optim.Adam(pb, lr=1e-2)
    opt = dp.FlatAdam(pa, lr=1e-2, clip=2.0)
    for it in range(6):
        gs = [torch.randn(s, device='cuda') * (3.0 if it % 2 else 0.05) for s in shapes]   # clipped / not clipped
        opt.grads.zero()
        for p, q, g in zip(pa, pb, gs):
            p.grad.copy_(g)
            q.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(pb, 2.0)
        ref.step()
        opt.step()
        torch.cuda.synchronize()
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), it
    assert float(opt.step_dev.item()) == 6.0


def test_clip_on_flat_buffer_matches_torch():
    """FlatGradients.clip_ on CUDA (gp_sumsq_f32 + gp_clip_scale_f32, with the pending 1/world factor of a SUM
    all-reduce) == scale by 1/world, then torch clip_grad_norm_."""
    from graph_pooling_b200 import dp
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, device='cuda')) for s in [(40, 9), (9,), (3, 3)]]
    fg = dp.FlatGradients(ps)
    for big, scale in ((1, 0.5), (0, 1.0), (1, 1.0)):
        g = torch.randn_like(fg.flat) * (5.0 if big else 0.01)
        fg.flat.copy_(g)
        fg.pending_scale = scale
        fg.clip_(2.0)
        ref = g * scale
        ref = ref * torch.clamp(2.0 / (ref.norm() + 1e-6), max=1.0)
        assert torch.allclose(fg.flat, ref, rtol=1e-5, atol=1e-7) and fg.pending_scale == 1.0


def test_backward_delivers_gradients_into_attached_buffer():
    """An attached FlatGradients receives the backward's parameter gradients through ONE gp_multi_axpy_f32 launch
    (accumulating: two backward passes add up), identical to what autograd's AccumulateGrad produces."""
    from graph_pooling_b200 import dp, encoders
    import copy
    B, N, D, H, C = 4, 40, 5, 16, 3
    x, adj, nb, label = synth_batch(8, B, N, D, 3, N, C, density=0.2)
    torch.manual_seed(2)
    ma = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25).cuda()
    mb = copy.deepcopy(ma)
    xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
    fg = dp.FlatGradients(ma.parameters()).attach(ma)
    fg.zero()
    for m in (ma, mb):
        for _ in range(2):
            yp = m(xc, ac, nb, assign_x=xc)
            m.loss(yp, lc, ac, nb).backward()
    torch.cuda.synchronize()
    ref = torch.cat([p.grad.reshape(-1) for p in mb.parameters()])
    assert rel_l2(fg.flat.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    off = 0
    for p in ma.parameters():
        assert p.grad.data_ptr() == fg.flat.data_ptr() + off * 4
        off += p.numel()


@pytest.mark.parametrize('precision', [0, 1])
def test_nccl_two_ranks(precision):
    """torchrun, 2 GPUs, NCCL: reduced gradient == mean of shard gradients == mean of the oracle's shard gradients;
    replicas bit-identical after FlatAdam steps (tests/dp_nccl_worker.py).  Skipped on a single-GPU box."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, GP_DP_PRECISION=str(precision))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', str(29731 + precision),
                        os.path.join(here, 'dp_nccl_worker.py')], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
    res = json.loads(line)
    assert res['ok'] and res['replica_param_spread'] == 0.0, res
