"""Error of the GP_BF16 (tensor-core) mode vs the fp64 oracle on seeded batches (GPU)."""
import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import rel_l2, synth_batch
from graph_pooling_b200 import encoders
from oracle import diffpool_oracle as orc

def run(N, D, H, C, B, ratio, P=1, n_min=None, density=0.05, seed=0, prec=1):
    torch.manual_seed(seed)
    mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio, num_pooling=P)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in mo.named_parameters():
            if k.endswith('bias'): p.copy_(0.2 * torch.randn(p.shape, generator=g))
    mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio, num_pooling=P)
    mc.load_state_dict(mo.state_dict()); mc = mc.cuda(); mc.precision = prec
    x, adj, nb, label = synth_batch(seed, B, N, D, n_min or N // 4, N, C, density)
    m64 = copy.deepcopy(mo).double()
    yo, lo = orc.train_step(m64, torch.tensor(x).double(), torch.tensor(adj).double(), torch.tensor(label), nb)
    xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
    yp = mc(xc, ac, nb, assign_x=xc); loss = mc.loss(yp, torch.tensor(label).cuda(), ac, nb); loss.backward()
    torch.cuda.synchronize()
    errs = {k: rel_l2(p.grad.cpu().numpy(), q.grad.numpy()) for (k, p), (_, q) in zip(mc.named_parameters(), m64.named_parameters())}
    worst = max(errs, key=errs.get)
    print('N=%d D=%d H=%d B=%d K=%d P=%d prec=%d: ypred %.2e S %.2e loss %.2e link %.2e | grads median %.2e worst %.2e (%s)' % (
        N, D, H, B, int(N * ratio), P, prec, rel_l2(yp.detach().cpu().numpy(), yo.detach().numpy()),
        rel_l2(mc.assign_tensors[0].detach().cpu().numpy(), m64.assign_tensors[0].detach().numpy()),
        abs(loss.item() - lo.item()) / abs(lo.item()), abs(mc.link_loss.item() - m64.link_loss.item()) / abs(m64.link_loss.item()),
        float(np.median(list(errs.values()))), errs[worst], worst), flush=True)

if __name__ == '__main__':
    for prec in (0, 1):
        run(64, 8, 16, 3, 4, 0.25, prec=prec)
        run(256, 16, 32, 2, 4, 0.25, prec=prec)
        run(100, 3, 30, 6, 20, 0.1, n_min=2, density=0.08, prec=prec)
        run(512, 64, 64, 2, 3, 0.25, density=0.02, prec=prec)
        run(256, 16, 32, 2, 3, 0.25, P=2, prec=prec)
