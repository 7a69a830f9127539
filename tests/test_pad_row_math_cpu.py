"""CPU: the facts a packed (ragged) layout for ENZYMES-sized graphs relies on ("packed ragged
layout"), checked against the oracle's own layer code (encoders.py:315-328, 1048-1064):

  1. after a GraphConv every pad row (n >= n_b) holds the SAME vector normalize(bias) -- A's pad rows are zero;
  2. the per-node-index BatchNorm statistics are therefore (real rows of the graphs that reach node n) + a count
     times one constant: nothing about pad rows needs to be stored or read;
  3. with a masked consumer (no gradient enters pad rows) the pad rows still receive a BatchNorm-backward gradient,
     but it is the same vector for every pad row of a node index, so their contribution to the bias gradient is
     count(n) * that vector.
Test infrastructure for the round-2 kernels: nothing here touches the CUDA path."""
import numpy as np
import torch

from helpers import synth_batch
from oracle import diffpool_oracle as orc


def _layer(x, adj, w, b):
    y = orc.graph_conv(x, adj, w, b, add_self=False, normalize=True)
    return y, orc.bn_per_node(torch.relu(y))


def test_pad_rows_are_one_constant_and_bn_stats_decompose():
    B, N, D, H = 7, 20, 5, 6
    x, adj, nb, _ = synth_batch(3, B, N, D, 2, N - 3, 2, 0.3)
    xt, at = torch.tensor(x, dtype=torch.float64), torch.tensor(adj, dtype=torch.float64)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(D, H, generator=g, dtype=torch.float64)
    b = torch.randn(H, generator=g, dtype=torch.float64)
    y, h = _layer(xt, at, w, b)
    ypad = b / b.norm().clamp_min(orc.EPS_NORM)
    for i in range(B):                                              # fact 1
        assert torch.allclose(y[i, nb[i]:], ypad.expand(N - nb[i], H), atol=1e-14)
    r = torch.relu(y)
    rp = torch.relu(ypad)
    for n in range(N):                                              # fact 2
        reach = [i for i in range(B) if nb[i] > n]
        cnt_pad = B - len(reach)
        s1 = sum(r[i, n].sum() for i in reach) + cnt_pad * rp.sum()
        s2 = sum((r[i, n] ** 2).sum() for i in reach) + cnt_pad * (rp ** 2).sum()
        mean = s1 / (B * H)
        var = s2 / (B * H) - mean ** 2
        ref_mean, ref_var = r[:, n].mean(), r[:, n].var(unbiased=False)
        assert abs(mean - ref_mean) < 1e-13 and abs(var - ref_var) < 1e-13
        hp = (rp - mean) / torch.sqrt(var + orc.EPS_BN)             # the BN output of every pad row at node n
        for i in range(B):
            if nb[i] <= n:
                assert torch.allclose(h[i, n], hp, atol=1e-12)


def test_pad_row_gradient_is_per_node_constant():
    B, N, D, H = 6, 16, 4, 5
    x, adj, nb, _ = synth_batch(5, B, N, D, 2, N - 2, 2, 0.3)
    xt, at = torch.tensor(x, dtype=torch.float64), torch.tensor(adj, dtype=torch.float64)
    g = torch.Generator().manual_seed(1)
    w = torch.randn(D, H, generator=g, dtype=torch.float64)
    b = torch.randn(H, generator=g, dtype=torch.float64, requires_grad=True)
    # V = (A.X).W + b kept as a leaf-like tensor so that dL/dV (what the layer-backward kernel emits) is visible
    v = (torch.matmul(torch.matmul(at, xt), w) + b)
    v.retain_grad()
    y = v / v.norm(dim=2, keepdim=True).clamp_min(orc.EPS_NORM)
    h = orc.bn_per_node(torch.relu(y))
    mask = orc.construct_mask(N, nb, 'cpu', torch.float64)
    up = torch.randn(B, N, H, generator=g, dtype=torch.float64)
    ((h * mask) * up).sum().backward()                              # masked consumer: no gradient enters pad rows
    dv = v.grad
    db_real = torch.zeros(H, dtype=torch.float64)
    db_pad = torch.zeros(H, dtype=torch.float64)
    for n in range(N):                                              # fact 3
        pads = [i for i in range(B) if nb[i] <= n]
        for i in range(B):
            if nb[i] > n:
                db_real += dv[i, n]
        if pads:
            for i in pads[1:]:
                assert torch.allclose(dv[i, n], dv[pads[0], n], atol=1e-13)
            db_pad += len(pads) * dv[pads[0], n]
    assert float(db_pad.abs().sum()) > 0                            # the pad rows DO carry bias gradient
    assert torch.allclose(db_real + db_pad, b.grad, atol=1e-12)
