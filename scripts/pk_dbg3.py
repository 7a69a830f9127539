import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import synth_batch
from graph_pooling_b200 import encoders, engine_pk
from test_gpu_loss_options import build
N, D, H, C, B = 70, 5, 24, 4, 6
mo, mc = build(N, D, H, C, 0.2, 'bce', 0.0, 3, linkpred=True)
x, adj, nb, label = synth_batch(11, B, N, D, 5, N, C, 0.15, symmetric=True)
xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
args = {}
for tag in ('packed', 'dense'):
    if tag == 'dense': os.environ['GP_NO_PACKED'] = '1'
    yp = mc(xc, ac, nb, assign_x=xc)
    os.environ.pop('GP_NO_PACKED', None)
    tape = yp.grad_fn.tape if hasattr(yp.grad_fn, 'tape') else None
    fn = yp.grad_fn
    t = getattr(fn, 'tape', None)
    args[tag] = (t['arg'].cpu().numpy().copy(), t['out'].cpu().numpy().copy())
a, b = args['packed'][0], args['dense'][0]
diff = np.argwhere(a != b)
print('argmax differences (graph, column):', diff.tolist())
for g, c in diff[:10]:
    print('  g', g, 'col', c, 'packed arg', a[g, c], 'dense arg', b[g, c], 'out packed %.9g dense %.9g' % (args['packed'][1][g, c], args['dense'][1][g, c]))
