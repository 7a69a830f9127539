import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import synth_batch, rel_l2
from graph_pooling_b200 import encoders
from test_gpu_loss_options import build
N, D, H, C, B = 70, 5, 24, 4, 6
for seed in (11, 12, 13):
  for dens in (0.15, 0.08):
    for sym in (1, 0):
        mo, mc = build(N, D, H, C, 0.2, 'bce', 0.0, 3, linkpred=True)
        x, adj, nb, label = synth_batch(seed, B, N, D, 5, N, C, dens, symmetric=bool(sym))
        xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
        res = {}
        for tag in ('packed', 'dense'):
            if tag == 'dense': os.environ['GP_NO_PACKED'] = '1'
            mc.zero_grad()
            yp = mc(xc, ac, nb, assign_x=xc)
            loss = mc.loss(yp, lc, ac, nb)
            loss.backward(); torch.cuda.synchronize()
            os.environ.pop('GP_NO_PACKED', None)
            res[tag] = torch.cat([p.grad.flatten() for p in mc.parameters()]).double().cpu()
        e = float((res['packed'] - res['dense']).norm() / res['dense'].norm())
        print('seed', seed, 'dens', dens, 'sym', sym, 'maxdeg', int(adj.sum(2).max()), 'maxindeg', int(adj.sum(1).max()), 'nb', list(nb), 'grad rel diff %.2e' % e)
