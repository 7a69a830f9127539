"""GPU debug aid: the packed small-graph schedule against the dense fp32 schedule (GP_NO_PACKED=1) and the fp64
oracle, per output and per parameter gradient."""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import rel_l2, synth_batch  # noqa: E402
from oracle import diffpool_oracle as orc  # noqa: E402
from graph_pooling_b200 import encoders  # noqa: E402


def run(m, x, adj, nb, label, ax=None, linkpred=True):
    m.zero_grad()
    xc, ac, lc = torch.as_tensor(x).cuda(), torch.as_tensor(adj).cuda(), torch.as_tensor(label).cuda()
    yp = m(xc, ac, nb, assign_x=xc if ax is None else torch.as_tensor(ax).cuda())
    loss = m.loss(yp, lc, ac, nb) if linkpred else m.loss(yp, lc)
    loss.backward()
    torch.cuda.synchronize()
    return (yp.detach().cpu().numpy(), loss.item(), m.assign_tensor.detach().cpu().numpy(),
            float(m.link_loss) if linkpred else 0.0, {k: p.grad.cpu().numpy().copy() for k, p in m.named_parameters()})


def case(seed, B, N, D, H, Eo, C, L, ratio, nmin, nmax, density=0.12, bias=True, symmetric=True, weighted=False,
         linkpred=True, assign_D=None):
    class A:
        pass
    a = A()
    a.bias = bias
    mk = lambda mod: mod.SoftPoolingGcnEncoder(N, D, H, Eo, C, L, H, assign_ratio=ratio, num_pooling=1,
                                               assign_input_dim=-1 if assign_D is None else assign_D, args=a,
                                               linkpred=linkpred)
    torch.manual_seed(seed)
    mo = mk(orc)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in mo.named_parameters():
            if k.endswith('bias'):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
    x, adj, nb, label = synth_batch(seed, B, N, D, nmin, nmax, C, density, symmetric, weighted)
    ax = None if assign_D is None else np.random.RandomState(seed + 7).randn(B, N, assign_D).astype(np.float32)
    m64 = copy.deepcopy(mo).double()
    yp, loss = orc.train_step(m64, torch.tensor(x).double(), torch.tensor(adj).double(), torch.tensor(label), nb,
                              assign_x=None if ax is None else torch.tensor(ax).double(), linkpred=linkpred)
    o = (yp.detach().numpy(), loss.item(), m64.assign_tensors[0].detach().numpy(),
         float(m64.link_loss) if linkpred else 0.0, {k: p.grad.numpy() for k, p in m64.named_parameters()})
    mc = mk(encoders)
    mc.load_state_dict(mo.state_dict(), strict=True)
    mc = mc.cuda()
    os.environ['GP_NO_PACKED'] = '1'
    d = run(mc, x, adj, nb, label, ax, linkpred)
    del os.environ['GP_NO_PACKED']
    p = run(mc, x, adj, nb, label, ax, linkpred)
    assert mc._plan.packed, 'packed schedule not selected'
    print('case seed=%d B=%d N=%d D=%d H=%d L=%d K=%d nb=%s' % (seed, B, N, D, H, L, int(N * ratio), list(nb[:8])))
    worst = 0.0
    for name, i in (('ypred', 0), ('S', 2)):
        e1, e2 = rel_l2(p[i], o[i]), rel_l2(d[i], o[i])
        worst = max(worst, e1)
        print('  %-28s packed %.2e   dense %.2e' % (name, e1, e2))
    print('  %-28s packed %.2e   dense %.2e' % ('loss', abs(p[1] - o[1]) / max(1, abs(o[1])), abs(d[1] - o[1]) / max(1, abs(o[1]))))
    print('  %-28s packed %.2e   dense %.2e' % ('link', abs(p[3] - o[3]), abs(d[3] - o[3])))
    G = max(np.linalg.norm(v) for v in o[4].values())
    for k in o[4]:
        e1 = np.linalg.norm(p[4][k] - o[4][k]) / max(np.linalg.norm(o[4][k]), 1e-7 * G)
        e2 = np.linalg.norm(d[4][k] - o[4][k]) / max(np.linalg.norm(o[4][k]), 1e-7 * G)
        worst = max(worst, e1)
        print('  %-28s packed %.2e   dense %.2e %s' % (k, e1, e2, '  <<<' if e1 > 1e-4 else ''))
    return worst


if __name__ == '__main__':
    w = []
    w.append(case(0, 6, 20, 3, 8, 9, 3, 3, 0.25, 2, 20))
    w.append(case(1, 20, 100, 3, 30, 30, 6, 3, 0.1, 1, 100))
    w.append(case(2, 5, 16, 4, 6, 7, 3, 3, 0.2, 16, 16))
    w.append(case(3, 7, 24, 3, 5, 6, 2, 3, 0.25, 1, 3))
    w.append(case(4, 4, 40, 4, 8, 8, 2, 3, 0.25, 3, 40, bias=False, weighted=True, linkpred=False))
    w.append(case(5, 5, 48, 5, 16, 12, 3, 3, 0.25, 1, 48, symmetric=False, assign_D=9))
    w.append(case(6, 4, 32, 6, 10, 14, 2, 2, 0.25, 1, 32))
    w.append(case(7, 4, 32, 6, 10, 14, 2, 4, 0.25, 1, 32))
    w.append(case(8, 300, 100, 3, 30, 30, 6, 3, 0.1, 2, 100, density=0.06))
    print('worst', max(w))
