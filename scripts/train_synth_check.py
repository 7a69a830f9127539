"""Stability check of the tensor-core mode at the large configuration: N steps of train.py:196-210 on a fixed synthetic
batch (256-node-count-2048 shapes, smaller batch), fp32 mode vs bf16 mode from the same initial weights; prints the
loss trajectories (both must fall and stay finite, and track each other)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_pooling_b200 import dp, encoders, synth
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
batch = synth.make_batch('cfg4_diffpool_256x2048', seed=0, device='cuda', B=B)
cfg = batch['cfg']
out = {}
for prec in (0, 1):
    torch.manual_seed(0)
    m = synth.build_model(encoders, cfg).cuda()
    m.precision = prec
    opt = dp.FlatAdam(list(m.parameters()), lr=1e-3, clip=2.0)
    x, adj, nb, label = batch['x'], batch['adj'], batch['nb'], batch['label']
    ls = []
    for i in range(steps):
        opt.grads.zero()
        yp = m(x, adj, nb, assign_x=x)
        loss = m.loss(yp, label, adj, nb)
        loss.backward()
        opt.step()
        ls.append((loss.item(), m.link_loss.item()))
    out[prec] = ls
    print('precision %s: loss %.4f -> %.4f   link %.4f -> %.4f   finite=%s' % ('bf16' if prec else 'fp32', ls[0][0], ls[-1][0], ls[0][1], ls[-1][1], bool(np.isfinite(np.array(ls)).all())))
d = max(abs(a[0] - b[0]) / abs(a[0]) for a, b in zip(out[0], out[1]))
print('max relative loss gap fp32 vs bf16 over %d steps: %.3e' % (steps, d))
