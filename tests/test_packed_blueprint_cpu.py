"""CPU: the packed small-graph schedule (tests/packed_blueprint.py: only real rows exist, pad rows are per-layer
constants with a count per node index, hand-written backward) reproduces the oracle's autograd results in float64 --
outputs, S, losses and every parameter gradient -- including non-trivial biases (pad rows then carry bias gradient),
graphs with n_b = N, tiny graphs, and the 'no pad row wins the readout' corner."""
import numpy as np
import pytest
import torch

import packed_blueprint as pb
from helpers import synth_batch
from oracle import diffpool_oracle as orc


def _stack(first, block, last):
    return pb.Stack([(m.weight.detach(), None if m.bias is None else m.bias.detach()) for m in [first] + list(block) + [last]])


@pytest.mark.parametrize('seed,B,N,D,H,K_ratio,nmin,nmax,bias', [(0, 6, 20, 3, 8, 0.25, 2, 20, True),
                                                                   (1, 5, 16, 4, 6, 0.2, 16, 16, True),
                                                                   (2, 7, 24, 3, 5, 0.25, 1, 3, True),
                                                                   (3, 4, 12, 5, 7, 0.34, 3, 12, False)])
def test_blueprint_matches_oracle(seed, B, N, D, H, K_ratio, nmin, nmax, bias):
    class A_:
        pass
    a = A_()
    a.bias = bias
    torch.manual_seed(seed)
    C = 3
    m = orc.SoftPoolingGcnEncoder(N, D, H, H + 1, C, 3, H, assign_ratio=K_ratio, args=a).double()
    g = torch.Generator().manual_seed(seed + 10)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith('bias'):
                p.copy_(0.3 * torch.randn(p.shape, generator=g, dtype=torch.float64))
    x, adj, nb, label = synth_batch(seed, B, N, D, nmin, nmax, C, 0.3)
    xt, at, lt = torch.tensor(x).double(), torch.tensor(adj).double(), torch.tensor(label)
    yp, loss = orc.train_step(m, xt, at, lt, nb)
    params = dict(emb=_stack(m.conv_first, m.conv_block, m.conv_last),
                  assign=_stack(m.assign_conv_first, m.assign_conv_block, m.assign_conv_last),
                  post=_stack(m.conv_first2, m.conv_block2, m.conv_last2),
                  assign_pred=(m.assign_pred.weight.detach(), None if m.assign_pred.bias is None else m.assign_pred.bias.detach()),
                  pred=[(l.weight.detach(), l.bias.detach()) for l in m.pred_model if isinstance(l, torch.nn.Linear)])
    r = pb.diffpool_step(params, xt, at, nb, lt)
    tol = 1e-9
    assert torch.allclose(r['ypred'], yp.detach(), atol=tol)
    assert abs(float(r['loss']) - float(loss)) < tol and abs(float(r['link']) - float(m.link_loss)) < tol
    for gi in range(B):
        assert torch.allclose(r['S'][gi], m.assign_tensors[0][gi, :nb[gi]].detach(), atol=tol)

    def chk(mods, grads, name):
        dW, db = grads
        for l, mod in enumerate(mods):
            assert torch.allclose(dW[l], mod.weight.grad, atol=tol, rtol=1e-7), (name, l, 'W')
            if mod.bias is not None:
                assert torch.allclose(db[l], mod.bias.grad, atol=tol, rtol=1e-7), (name, l, 'b')
    chk([m.conv_first] + list(m.conv_block) + [m.conv_last], r['emb'], 'emb')
    chk([m.assign_conv_first] + list(m.assign_conv_block) + [m.assign_conv_last], r['assign'], 'assign')
    chk([m.conv_first2] + list(m.conv_block2) + [m.conv_last2], r['post'], 'post')
    assert torch.allclose(r['assign_pred'][0], m.assign_pred.weight.grad, atol=tol, rtol=1e-7)
    if m.assign_pred.bias is not None:
        assert torch.allclose(r['assign_pred'][1], m.assign_pred.bias.grad, atol=tol, rtol=1e-7)
    lins = [l for l in m.pred_model if isinstance(l, torch.nn.Linear)]
    for (dw, dbb), l in zip(r['pred'], lins):
        assert torch.allclose(dw, l.weight.grad, atol=tol, rtol=1e-7) and torch.allclose(dbb, l.bias.grad, atol=tol, rtol=1e-7)
