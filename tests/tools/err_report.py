"""Per-tensor error report: candidate (GPU) and fp32 oracle vs the fp64 golden vectors."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import GOLDEN, golden_inputs, golden_model, load_golden, rel_l2
from graph_pooling_b200 import encoders
for name in GOLDEN:
    g = load_golden(name)
    soft = str(g['kind']) == 'soft'
    m = golden_model(g, encoders, device='cuda')
    x, adj, nb, label = golden_inputs(g)
    xc, ac, lc = x.cuda(), adj.cuda(), label.cuda()
    yp = m(xc, ac, nb, assign_x=xc) if soft else m(xc, ac, nb)
    loss = m.loss(yp, lc, ac, nb) if soft else m.loss(yp, lc)
    loss.backward(); torch.cuda.synchronize()
    print('==', name, 'ypred cand %.2e oracle32 %.2e' % (rel_l2(yp.detach().cpu().numpy(), g['f64.ypred']), rel_l2(g['f32.ypred'], g['f64.ypred'])),
          'loss cand %.2e' % abs(loss.item() - float(g['f64.loss'])))
    if soft:
        print('   S cand %.2e oracle32 %.2e' % (rel_l2(m.assign_tensor.detach().cpu().numpy(), g['f64.S']), rel_l2(g['f32.S'], g['f64.S'])))
    for k, p in m.named_parameters():
        g64 = g['f64.grad.' + k]
        print('   %-28s |g|=%.2e cand %.2e oracle32 %.2e' % (k, np.linalg.norm(g64), rel_l2(p.grad.cpu().numpy(), g64), rel_l2(g['f32.grad.' + k], g64)))
