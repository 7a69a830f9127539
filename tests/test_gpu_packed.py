"""GPU: the PACKED small-graph schedule (csrc/packed.cu, engine_pk.py) against the fp64 oracle -- outputs, S, losses
and every parameter gradient at the fp32 bar of tests/test_gpu_model.py -- on the shapes it is selected for
(soft-assign, one pooling level, N <= 128), including a large batch (BatchNorm sums over many CTAs), device-resident
node counts, no node counts at all, tiny graphs, asymmetric / weighted adjacencies and separate assignment features.
The dense fp32 schedule (GP_NO_PACKED=1) must agree with it as well."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_l2, synth_batch
from test_gpu_model import enc, oracle_vs_candidate, run_candidate, soft_factory

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,N,D,H,K_ratio,nb_mode,kw', [
    (300, 100, 3, 30, 0.1, 'rand', dict(density=0.06)),
    (64, 128, 7, 20, 0.25, 'rand', dict(density=0.05)),
    (9, 24, 3, 5, 0.25, 'tiny', {}),
    (6, 40, 4, 12, 0.2, 'full', {}),
    (6, 40, 4, 12, 0.2, 'none', {}),
    (5, 48, 5, 16, 0.25, 'rand', dict(symmetric=False, assign_D=9)),
    (4, 40, 4, 8, 0.25, 'rand', dict(weighted=True, linkpred=False, bias=False)),
])
def test_packed_matches_oracle(B, N, D, H, K_ratio, nb_mode, kw):
    kw = dict(kw)
    bias, linkpred, assign_D = kw.pop('bias', True), kw.get('linkpred', True), kw.get('assign_D')
    mk = soft_factory(N, D, H, H + 1, 3, ratio=K_ratio, bias=bias, linkpred=linkpred,
                      assign_D=-1 if assign_D is None else assign_D)
    mc = oracle_vs_candidate(mk, 31 + B, B, N, D, 3, nb_mode=nb_mode, **kw)
    assert mc._plan.packed, 'the packed schedule was not selected'


def test_packed_agrees_with_dense_schedule_and_device_node_counts():
    B, N, D = 40, 100, 3
    mk = soft_factory(N, D, 30, 30, 6, ratio=0.1)
    torch.manual_seed(5)
    m = mk(enc()).cuda()
    x, adj, nb, label = synth_batch(77, B, N, D, 1, N, 6, 0.08)
    res = {}
    for tag in ('packed', 'dense', 'packed_dev_nb'):
        if tag == 'dense':
            os.environ['GP_NO_PACKED'] = '1'
        try:
            nbx = torch.from_numpy(nb.astype(np.int32)).cuda() if tag == 'packed_dev_nb' else nb
            yp, loss = run_candidate(m, x, adj, nbx, label, True)
        finally:
            os.environ.pop('GP_NO_PACKED', None)
        assert m._plan.packed == (tag != 'dense')
        res[tag] = (yp.detach().cpu().numpy(), loss.item(), m.assign_tensor.detach().cpu().numpy(),
                    np.concatenate([p.grad.cpu().numpy().ravel() for p in m.parameters()]))
    for tag in ('dense', 'packed_dev_nb'):
        tol = 1e-6 if tag == 'packed_dev_nb' else 2e-5
        assert rel_l2(res[tag][0], res['packed'][0]) < tol
        assert abs(res[tag][1] - res['packed'][1]) < tol * max(1.0, abs(res['packed'][1]))
        assert rel_l2(res[tag][2], res['packed'][2]) < tol
        assert rel_l2(res[tag][3], res['packed'][3]) < (1e-5 if tag == 'packed_dev_nb' else 2e-4)
    # pad rows of S are exact zeros
    S = res['packed'][2]
    for b in range(B):
        assert not S[b, nb[b]:].any()


def test_packed_step_under_graph_replay_tracks_eager_training():
    """graphed.GraphedTrainStep captures the packed schedule (node counts on the device only: rows, runs and lists are
    all derived there) and its replays track eager clip_grad_norm + Adam training on the same batches."""
    import copy
    from graph_pooling_b200 import graphed
    N, D, H, C, B = 100, 3, 30, 6, 24
    torch.manual_seed(8)
    m0 = soft_factory(N, D, H, H, C, ratio=0.1)(enc()).cuda()
    me, mg = copy.deepcopy(m0), copy.deepcopy(m0)
    opt = torch.optim.Adam(me.parameters(), lr=1e-3)
    gs = graphed.GraphedTrainStep(mg, lr=1e-3, clip=2.0)
    for i in range(4):
        x, adj, nb, label = synth_batch(90 + i, B, N, D, 2, N, C, 0.05)
        xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
        me.zero_grad()
        yp = me(xc, ac, nb, assign_x=xc)
        loss = me.loss(yp, lc, ac, nb)
        assert me._plan.packed
        loss.backward()
        torch.nn.utils.clip_grad_norm_(me.parameters(), 2.0)
        opt.step()
        del yp
        _, lg = gs.step(xc, ac, nb, lc)
        assert abs(lg.item() - loss.item()) <= 2e-4 * abs(loss.item()), (i, lg.item(), loss.item())
        del loss
    assert mg._plan.packed
