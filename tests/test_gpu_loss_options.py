"""GPU: the north-star loss options the reference does NOT contain -- Frobenius link loss and row entropy
(SURVEY appendix A.6).  Oracle = oracle/diffpool_oracle.py frobenius_link_loss / row_entropy_loss (the DiffPool
paper's definitions; parity unpinned by the reference).  Defaults (link_loss='bce', entropy_weight=0) must leave
the reference behaviour untouched -- that is what every other test file checks.

Tolerances: fp32 mode 1e-5 rel-L2 on outputs / loss, gradients graded like tests/test_gpu_model.py;
bf16 mode: loss 5e-3 relative, flattened gradient rel-L2 <= 0.15 and cosine >= 0.99."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2, synth_batch
from oracle import diffpool_oracle as orc

pytestmark = pytest.mark.gpu


def oracle_step(m64, x, adj, nb, label, link, ent_w, assign_x=None):
    m64.zero_grad()
    xt, at = torch.tensor(x).double(), torch.tensor(adj).double()
    yp = m64(xt, at, nb, assign_x=xt if assign_x is None else torch.tensor(assign_x).double())
    s0 = m64.assign_tensors[0]
    loss = F.cross_entropy(yp, torch.tensor(label))
    parts = {}
    if link == 'frobenius':
        parts['link'] = orc.frobenius_link_loss(s0, at, nb)
        loss = loss + parts['link']
    elif link == 'bce':
        parts['link'] = orc.link_pred_loss(s0, at, nb)
        loss = loss + parts['link']
    if ent_w:
        parts['ent'] = orc.row_entropy_loss(s0, nb)
        loss = loss + ent_w * parts['ent']
    loss.backward()
    return yp, loss, parts


def build(N, D, H, C, ratio, link, ent_w, seed, linkpred=True):
    from graph_pooling_b200 import encoders
    torch.manual_seed(seed)
    mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio, linkpred=linkpred)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in mo.named_parameters():
            if k.endswith('bias'):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio, linkpred=linkpred,
                                        link_loss=link or 'bce', entropy_weight=ent_w)
    mc.load_state_dict(mo.state_dict())
    return mo, mc.cuda()


@pytest.mark.parametrize('link,ent_w,use_nb,sym', [('frobenius', 0.0, 1, 1), ('frobenius', 0.5, 1, 0),
                                                   ('bce', 0.3, 1, 1), (None, 1.0, 1, 1), ('frobenius', 0.2, 0, 1)])
def test_fp32_loss_options(link, ent_w, use_nb, sym):
    N, D, H, C, B = 70, 5, 24, 4, 6
    mo, mc = build(N, D, H, C, 0.2, link, ent_w, 3, linkpred=link is not None)
    # seed 12: with seed 11 the pooled level has two clusters whose maxima differ by 2 ulp (1.97006226 vs 1.97006261);
    # the fp32 schedules then disagree with the fp64 oracle on WHICH row wins the max readout -- a legitimate fp32
    # artefact of a non-differentiable point, not an error -- and every gradient downstream moves by ~1e-4
    x, adj, nb, label = synth_batch(12, B, N, D, 5, N, C, 0.15, symmetric=bool(sym))
    nbo = nb if use_nb else None
    m64 = copy.deepcopy(mo).double()
    yo, lo, parts = oracle_step(m64, x, adj, nbo, label, link, ent_w)
    xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
    yp = mc(xc, ac, nbo, assign_x=xc)
    loss = mc.loss(yp, lc, ac, nbo) if link is not None else mc.loss(yp, lc)
    loss.backward()
    torch.cuda.synchronize()
    assert rel_l2(yp.detach().cpu().numpy(), yo.detach().numpy()) < 1e-5
    assert abs(loss.item() - lo.item()) < 1e-5 * max(1.0, abs(lo.item()))
    if 'link' in parts:
        assert abs(mc.link_loss.item() - parts['link'].item()) < 1e-5 * max(1.0, abs(parts['link'].item()))
    if 'ent' in parts:
        assert abs(mc.entropy_loss.item() - parts['ent'].item()) < 1e-5 * max(1.0, abs(parts['ent'].item()))
    G = max(float(p.grad.norm()) for p in m64.parameters())
    for (k, pc), (_, po) in zip(mc.named_parameters(), m64.named_parameters()):
        err = float((pc.grad.cpu().double() - po.grad).norm())
        assert err <= 2e-5 * float(po.grad.norm()) + 2e-6 * G, (k, err, float(po.grad.norm()))


@pytest.mark.parametrize('link,ent_w,N,H', [('frobenius', 0.0, 256, 32), ('frobenius', 0.3, 200, 24), ('bce', 0.5, 256, 32)])
def test_bf16_loss_options(link, ent_w, N, H):
    D, C, B = 16, 2, 4
    mo, mc = build(N, D, H, C, 0.25, link, ent_w, 5)
    mc.precision = 1
    x, adj, nb, label = synth_batch(21, B, N, D, N // 4, N, C, 0.05)
    m64 = copy.deepcopy(mo).double()
    yo, lo, parts = oracle_step(m64, x, adj, nb, label, link, ent_w)
    xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
    yp = mc(xc, ac, nb, assign_x=xc)
    loss = mc.loss(yp, lc, ac, nb)
    loss.backward()
    torch.cuda.synchronize()
    assert rel_l2(yp.detach().cpu().numpy(), yo.detach().numpy()) < 2e-2
    assert abs(loss.item() - lo.item()) < 5e-3 * abs(lo.item())
    gc = np.concatenate([p.grad.cpu().numpy().ravel() for p in mc.parameters()]).astype(np.float64)
    go = np.concatenate([p.grad.numpy().ravel() for p in m64.parameters()])
    cos = float(gc @ go / (np.linalg.norm(gc) * np.linalg.norm(go)))
    assert rel_l2(gc, go) < 0.15 and cos > 0.99, (rel_l2(gc, go), cos)


def test_defaults_keep_reference_loss():
    """link_loss / entropy_weight default to the reference's behaviour (masked BCE, no entropy term)."""
    from graph_pooling_b200 import encoders
    m = encoders.SoftPoolingGcnEncoder(40, 3, 16, 16, 3, 3, 16)
    assert m.link_loss_kind == 'bce' and m.entropy_weight == 0.0
    with pytest.raises(ValueError):
        encoders.SoftPoolingGcnEncoder(40, 3, 16, 16, 3, 3, 16, link_loss='l1')
