"""GPU: the edge-list feed (gp_adj_from_edges, engine_tc.PreparedAdjacency.from_edges, feed.EdgeListFeed) builds the
same bf16 adjacency operand as the dense fp32 feed, and a train step fed through it is identical."""
import numpy as np
import pytest
import torch

from helpers import synth_batch

pytestmark = pytest.mark.gpu


def _edges(adj, nb):
    B = adj.shape[0]
    eptr, chunks = [0], []
    for b in range(B):
        u, v = np.nonzero(np.triu(adj[b], 1))
        chunks.append(np.stack([u, v], 1).astype(np.int32))
        eptr.append(eptr[-1] + len(u))
    return np.concatenate(chunks).reshape(-1, 2), np.asarray(eptr, np.int32)


def test_operand_from_edges_equals_operand_from_dense():
    from graph_pooling_b200 import engine as E, engine_tc as T
    B, N = 5, 200
    x, adj, nb, label = synth_batch(3, B, N, 4, 10, N, 2, 0.05)
    e, eptr = _edges(adj, nb)
    ws = E.Workspace(torch.device('cuda'))
    ad = torch.tensor(adj).cuda()
    nbd = torch.tensor(nb).cuda()
    op, flags = T.adj_prepare(ws, ad, nbd, B, N)
    pa = T.PreparedAdjacency(B, N, 'cuda')
    pa.from_edges(torch.tensor(e).cuda(), torch.tensor(eptr).cuda(), int(np.diff(eptr).max()))
    torch.cuda.synchronize()
    torch.cuda.synchronize()
    a, b = op.t.view(torch.int16), pa.op.t.view(torch.int16)
    assert torch.equal(a, b)
    assert flags.tolist() == [0, 0] and pa.flags.tolist() == [0, 0]


def test_train_step_through_the_edge_list_feed():
    from graph_pooling_b200 import encoders, feed
    B, N, D = 4, 256, 8
    x, adj, nb, label = synth_batch(9, B, N, D, 20, N, 2, 0.04)
    e, eptr = _edges(adj, nb)
    torch.manual_seed(1)
    m = encoders.SoftPoolingGcnEncoder(N, D, 16, 16, 2, 3, 16, assign_ratio=0.25).cuda()
    m.precision = 1
    xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
    res = []
    for mode in ('dense', 'edges'):
        m.zero_grad()
        if mode == 'dense':
            a_in = ac
        else:
            fd = feed.EdgeListFeed(B, N, 'cuda', max_edges=len(e))
            fd.copy(0, torch.tensor(e).pin_memory(), torch.tensor(eptr).pin_memory(), int(np.diff(eptr).max()))
            a_in = fd.prepared(0)
        yp = m(xc, a_in, nb, assign_x=xc)
        loss = m.loss(yp, lc, a_in, nb)
        loss.backward()
        torch.cuda.synchronize()
        res.append((yp.detach().clone(), loss.item(), float(m.link_loss),
                    torch.cat([p.grad.flatten() for p in m.parameters()]).clone()))
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]
    # split-K reductions accumulate with atomics: the gradients agree to the last few bits, not bitwise
    assert float((res[0][3] - res[1][3]).norm() / res[0][3].norm()) < 1e-5


def test_bf16_features_and_16bit_edges_change_nothing():
    """The tensor-core schedule rounds the features to bf16 itself: feeding them as bf16 (and the edge ids as int16)
    gives the bit-identical forward."""
    from graph_pooling_b200 import encoders, feed
    B, N, D = 3, 128, 16
    x, adj, nb, label = synth_batch(4, B, N, D, 20, N, 2, 0.05)
    e, eptr = _edges(adj, nb)
    torch.manual_seed(2)
    m = encoders.SoftPoolingGcnEncoder(N, D, 16, 16, 2, 3, 16, assign_ratio=0.25).cuda()
    m.precision = 1
    xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
    with torch.no_grad():
        y0 = m(xc, ac, nb, assign_x=xc)
        fd = feed.EdgeListFeed(B, N, 'cuda', max_edges=len(e), edge_dtype=torch.int16)
        fd.copy(0, torch.tensor(e.astype(np.int16)).pin_memory(), torch.tensor(eptr).pin_memory(), int(np.diff(eptr).max()))
        xb = xc.to(torch.bfloat16)
        y1 = m(xb, fd.prepared(0), nb, assign_x=xb)
    assert torch.equal(y0, y1)
    m.precision = 0
    with pytest.raises(ValueError):
        m(xb, ac, nb, assign_x=xb)
