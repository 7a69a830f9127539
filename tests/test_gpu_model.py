"""GPU: module-level parity of the drop-in encoders (forward, loss, every parameter gradient)
against (a) the golden vectors produced by the actual reference and (b) the oracle on fresh
seeded batches covering the reference's edge cases.

Tolerance (fp32 path, BASELINE.json north_star "<=1e-5 relative"): outputs are graded by rel-L2
against the fp64 result; gradients by
    ||cand - g64|| <= max(1e-5, 4 x relerr(oracle_fp32, fp64)) * ||g64|| + 1e-7 * G
because the fp32 oracle itself is up to 1e-4 from fp64 on some bias gradients; G is
the largest parameter-gradient norm of the model, so the last term is an fp32-epsilon absolute floor
for gradients that cancel to ~0 (e.g. the 2-class output bias, whose entries sum to zero).
"""
import copy

import numpy as np
import pytest
import torch

from helpers import GOLDEN, golden_inputs, golden_model, load_golden, rel_l2, synth_batch
from oracle import diffpool_oracle as orc

pytestmark = pytest.mark.gpu
OUT_TOL = 1e-5


def enc():
    from graph_pooling_b200 import encoders
    return encoders


def run_candidate(m, x, adj, nb, label, soft, assign_x=None, linkpred=True):
    m.zero_grad()
    xc, ac, lc = torch.as_tensor(x).cuda(), torch.as_tensor(adj).cuda(), torch.as_tensor(label).cuda()
    if soft:
        yp = m(xc, ac, nb, assign_x=xc if assign_x is None else torch.as_tensor(assign_x).cuda())
        loss = m.loss(yp, lc, ac, nb) if linkpred else m.loss(yp, lc)
    else:
        yp = m(xc, ac, nb)
        loss = m.loss(yp, lc)
    loss.backward()
    torch.cuda.synchronize()
    return yp, loss


def grade_grads(cand, g32, g64, floor=1e-7):
    """cand/g32/g64: dict name -> numpy grad.  Per parameter: error <= max(1e-5, 4 x the fp32 oracle's own error)
    relative to that parameter's gradient, plus `floor` x the largest parameter gradient (cancellation-dominated
    gradients -- a bias behind an L2 normalise -- are tiny sums of large terms)."""
    G = max(np.linalg.norm(v) for v in g64.values())
    for k in g64:
        scale = max(np.linalg.norm(g64[k]), 1e-30)
        a_c = np.linalg.norm(cand[k].astype(np.float64) - g64[k])
        e_o = np.linalg.norm(g32[k].astype(np.float64) - g64[k]) / scale
        assert a_c <= max(1e-5, 4 * e_o) * scale + floor * G, \
            '%s: cand err %.3g, fp32-oracle err %.3g' % (k, a_c / scale, e_o)


@pytest.mark.parametrize('name', GOLDEN)
def test_golden_vectors(name):
    g = load_golden(name)
    soft = str(g['kind']) == 'soft'
    m = golden_model(g, enc(), device='cuda')
    x, adj, nb, label = golden_inputs(g)
    yp, loss = run_candidate(m, x, adj, nb, label, soft)
    assert rel_l2(yp.detach().cpu().numpy(), g['f64.ypred']) < OUT_TOL
    assert abs(loss.item() - float(g['f64.loss'])) < OUT_TOL * max(1.0, abs(float(g['f64.loss'])))
    if soft:
        assert rel_l2(m.assign_tensor.detach().cpu().numpy(), g['f64.S']) < OUT_TOL
        assert abs(m.link_loss.item() - float(g['f64.link_loss'])) < OUT_TOL
    cand = {k: p.grad.cpu().numpy() for k, p in m.named_parameters()}
    grade_grads(cand, {k: g['f32.grad.' + k] for k in cand}, {k: g['f64.grad.' + k] for k in cand})


def oracle_vs_candidate(make, seed, B, N, D, C, nb_mode='rand', soft=True, symmetric=True, weighted=False,
                        linkpred=True, density=0.12, bias_scale=0.3, assign_D=None):
    torch.manual_seed(seed)
    mo = make(orc)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in mo.named_parameters():
            if k.endswith('bias'):
                p.copy_(bias_scale * torch.randn(p.shape, generator=g))
    mc = make(enc())
    mc.load_state_dict(mo.state_dict(), strict=True)
    mc = mc.cuda()
    n_min, n_max = {'rand': (1, N), 'full': (N, N), 'tiny': (1, 2)}[nb_mode if nb_mode != 'none' else 'full']
    x, adj, nb, label = synth_batch(seed, B, N, D, n_min, n_max, C, density, symmetric, weighted)
    ax = None
    if assign_D is not None:
        ax = np.random.RandomState(seed + 7).randn(B, N, assign_D).astype(np.float32)
    nbo = None if nb_mode == 'none' else nb
    res = {}
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        m = copy.deepcopy(mo).to(dt)
        xt, at = torch.tensor(x, dtype=dt), torch.tensor(adj, dtype=dt)
        yp, loss = orc.train_step(m, xt, at, torch.tensor(label), nbo,
                                  assign_x=None if ax is None else torch.tensor(ax, dtype=dt), linkpred=linkpred)
        res[tag] = (yp.detach().numpy(), loss.item(), {k: p.grad.numpy() for k, p in m.named_parameters()},
                    m.assign_tensors[0].detach().numpy() if soft else None)
    yp, loss = run_candidate(mc, x, adj, nbo, label, soft, assign_x=ax, linkpred=linkpred)
    y64, l64, g64, s64 = res['f64']
    assert rel_l2(yp.detach().cpu().numpy(), y64) < OUT_TOL
    assert abs(loss.item() - l64) < OUT_TOL * max(1.0, abs(l64))
    if soft:
        assert rel_l2(mc.assign_tensors[0].detach().cpu().numpy(), s64) < OUT_TOL
    grade_grads({k: p.grad.cpu().numpy() for k, p in mc.named_parameters()}, res['f32'][2], g64)
    return mc


def soft_factory(N, D, H, E_, C, L=3, ratio=0.25, P=1, assign_D=-1, bias=True, linkpred=True):
    class A:
        pass
    a = A()
    a.bias = bias
    return lambda mod: mod.SoftPoolingGcnEncoder(N, D, H, E_, C, L, H, assign_ratio=ratio, num_pooling=P,
                                                 assign_input_dim=assign_D, args=a, linkpred=linkpred)


@pytest.mark.parametrize('nb_mode', ['rand', 'full', 'tiny', 'none'])
def test_soft_enzymes_shape_nb_modes(nb_mode):
    oracle_vs_candidate(soft_factory(100, 3, 30, 30, 6, ratio=0.1), 10, 20, 100, 3, 6, nb_mode=nb_mode)


def test_soft_nonsymmetric_adj_and_separate_assign_features():
    oracle_vs_candidate(soft_factory(48, 5, 16, 12, 3, assign_D=9), 11, 5, 48, 5, 3, symmetric=False, assign_D=9)


def test_soft_weighted_adj_no_linkpred_no_bias():
    oracle_vs_candidate(soft_factory(40, 4, 8, 8, 2, bias=False, linkpred=False), 12, 4, 40, 4, 2, weighted=True, linkpred=False)


def test_soft_wide_tiles():
    # crosses the 64- and 128-wide tile boundaries of every contraction
    oracle_vs_candidate(soft_factory(200, 20, 70, 40, 2, ratio=0.4), 13, 3, 200, 20, 2, density=0.05)


def test_soft_two_layers_and_four_layers():
    oracle_vs_candidate(soft_factory(32, 6, 10, 14, 2, L=2), 14, 4, 32, 6, 2)
    oracle_vs_candidate(soft_factory(32, 6, 10, 14, 2, L=4), 15, 4, 32, 6, 2)


def test_soft_num_pooling_2():
    # repaired-intent P=2 (R5-R7): no runnable ground truth in the reference, oracle only
    oracle_vs_candidate(soft_factory(64, 5, 12, 12, 2, ratio=0.25, P=2), 16, 4, 64, 5, 2)


def test_soft_batch_of_one_graph_smaller_than_K():
    oracle_vs_candidate(soft_factory(40, 3, 8, 8, 2, ratio=0.5), 17, 2, 40, 3, 2, nb_mode='tiny')


@pytest.mark.parametrize('bn,concat,L', [(True, True, 3), (False, True, 3), (True, False, 3), (True, True, 2)])
def test_base_variants(bn, concat, L):
    make = lambda mod: mod.GcnEncoderGraph(7, 20, 24, 3, L, bn=bn, concat=concat)
    oracle_vs_candidate(make, 20 + L, 5, 60, 7, 3, soft=False)


def test_base_with_pred_hidden_and_dd_like_dims():
    make = lambda mod: mod.GcnEncoderGraph(89, 20, 20, 2, 3, pred_hidden_dims=[50, 10])
    oracle_vs_candidate(make, 30, 3, 150, 89, 2, soft=False, density=0.03)


@pytest.mark.parametrize('nb_mode,hidden,bn', [('rand', [], True), ('none', [10], True), ('tiny', [], False)])
def test_set2set_encoder(nb_mode, hidden, bn):
    """method=base-set2set (encoders.py:1137-1157 + set2set.py): GCN concat, mask, n LSTM/attention steps, pred_model --
    forward, loss and every gradient (LSTM weights included) against the oracle; SURVEY 8(f) N4."""
    make = lambda mod: mod.GcnSet2SetEncoder(6, 10, 12, 3, 3, pred_hidden_dims=hidden, bn=bn)
    oracle_vs_candidate(make, 40 + len(hidden), 5, 30, 6, 3, nb_mode=nb_mode, soft=False)


def test_set2set_module_standalone_and_wide_features():
    """The Set2Set module on its own (d = 96: several feature chunks per lane, n = 70 steps) against the oracle's."""
    torch.manual_seed(5)
    B, n, d = 3, 70, 96
    so = orc.Set2Set(d, 2 * d).double()
    sc = enc().Set2Set(d, 2 * d)
    sc.load_state_dict({k: v.float() for k, v in so.state_dict().items()})
    sc = sc.cuda()
    emb = 0.5 * torch.randn(B, n, d, dtype=torch.float64)
    eo = emb.clone().requires_grad_()
    oo = so(eo)
    g = torch.randn_like(oo)
    (oo * g).sum().backward()
    ec = emb.float().cuda().requires_grad_()
    oc = sc(ec)
    (oc * g.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel_l2(oc.detach().cpu().numpy(), oo.detach().numpy()) < 1e-5
    assert rel_l2(ec.grad.cpu().numpy(), eo.grad.numpy()) < 1e-4
    for (k, pc), (_, po) in zip(sc.named_parameters(), so.named_parameters()):
        assert rel_l2(pc.grad.cpu().numpy(), po.grad.numpy()) < 1e-4, k


@pytest.mark.parametrize('kind,prec', [('base', 0), ('soft', 0), ('base', 1), ('soft', 1), ('s2s', 0)])
def test_dropout_matches_oracle_with_the_same_masks(kind, prec):
    """dropout > 0 (conv_block layers only, training mode only: encoders.py:316-317,1015,1187).  The library draws its
    own masks (counter-based hash of a seed taken from torch's generator), so parity is checked with the SAME masks:
    they are rebuilt from the forward's seed and injected into the oracle's nn.Dropout slots."""
    from graph_pooling_b200 import engine as E_
    e, p = enc(), 0.3
    N, D, H, C, B, L, K = 40, 6, 16, 3, 5, 4, 10          # L = 4: two conv_block layers per stack
    soft = kind == 'soft'
    if kind == 'base':
        make = lambda mod: mod.GcnEncoderGraph(D, H, H, C, L, dropout=p)
    elif kind == 's2s':
        make = lambda mod: mod.GcnSet2SetEncoder(D, H, H, C, L, dropout=p)
    else:
        make = lambda mod: mod.SoftPoolingGcnEncoder(N, D, H, H, C, L, H, assign_ratio=0.25, dropout=p)
    torch.manual_seed(77)
    mo = make(orc)
    g = torch.Generator().manual_seed(78)
    with torch.no_grad():
        for k, q in mo.named_parameters():
            if k.endswith('bias'):
                q.copy_(0.2 * torch.randn(q.shape, generator=g))
    mc = make(e)
    mc.load_state_dict(mo.state_dict(), strict=True)
    mc = mc.cuda()
    mc.precision = prec
    x, adj, nb, label = synth_batch(79, B, N, D, 5, N, C, 0.12)
    # eval mode: nn.Dropout is the identity -> identical to a model built without dropout
    mc.eval()
    m0 = (e.GcnEncoderGraph(D, H, H, C, L) if kind == 'base' else e.GcnSet2SetEncoder(D, H, H, C, L) if kind == 's2s'
          else e.SoftPoolingGcnEncoder(N, D, H, H, C, L, H, assign_ratio=0.25)).cuda()
    m0.load_state_dict(mc.state_dict())
    m0.precision = prec
    xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
    with torch.no_grad():
        assert torch.equal(mc(xc, ac, nb, assign_x=xc), m0(xc, ac, nb, assign_x=xc))
    mc.train()
    yp, loss = run_candidate(mc, x, adj, nb, label, soft)
    plan = mc._plan
    assert plan.seed != 0
    y_again, _ = run_candidate(mc, x, adj, nb, label, soft)
    assert not torch.equal(yp, y_again)                   # a fresh mask per forward call
    yp, loss = run_candidate(mc, x, adj, nb, label, soft)
    plan = mc._plan

    def mask(seed, l, Bn, rows):
        ones = torch.ones(Bn * rows, H, device='cuda')
        out = torch.empty_like(ones)
        E_.dropout(ones.data_ptr(), H, Bn * rows, H, p, E_.layer_seed(seed, l), out.data_ptr(), H)
        torch.cuda.synchronize()
        return out.view(Bn, rows, H).cpu()

    def inject(m):
        blocks, seed, rows = (m.conv_block_after_pool[0], plan.seed + 4096, K) if soft else (m.conv_block, plan.seed, N)
        for i, blk in enumerate(blocks):
            # the tensor-core schedule runs the pooled level at r8(K) rows per graph (dead clusters): same mask stream
            alloc = (rows + 7) // 8 * 8 if (soft and prec == 1) else rows
            mk = mask(seed, i + 1, B, alloc)[:, :rows].contiguous()
            keep = float((mk > 0).float().mean())
            assert abs(keep - (1 - p)) < 0.06
            assert bool(((mk == 0) | ((mk - 1 / (1 - p)).abs() < 1e-5)).all())
            del blk.dropout_layer
            object.__setattr__(blk, 'dropout_layer', (lambda mk: lambda t: t * mk.to(t.dtype))(mk))
        return m

    res = {}
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        m = inject(copy.deepcopy(mo).to(dt).train())
        yo, lo = orc.train_step(m, torch.tensor(x, dtype=dt), torch.tensor(adj, dtype=dt), torch.tensor(label), nb)
        res[tag] = (yo.detach().numpy(), lo.item(), {k: q.grad.numpy() for k, q in m.named_parameters()})
    y64, l64, g64 = res['f64']
    cand = {k: q.grad.cpu().numpy() for k, q in mc.named_parameters()}
    if prec == 0:
        assert rel_l2(yp.detach().cpu().numpy(), y64) < OUT_TOL
        assert abs(loss.item() - l64) < OUT_TOL * max(1.0, abs(l64))
        grade_grads(cand, res['f32'][2], g64)
    else:
        assert rel_l2(yp.detach().cpu().numpy(), y64) < 2e-2
        assert abs(loss.item() - l64) < 5e-3 * abs(l64)
        gc = np.concatenate([cand[k].ravel() for k in sorted(cand)]).astype(np.float64)
        go = np.concatenate([g64[k].ravel() for k in sorted(g64)])
        assert rel_l2(gc, go) < 0.15 and float(gc @ go / (np.linalg.norm(gc) * np.linalg.norm(go))) > 0.99


def test_padding_invariance_and_pad_row_features_ignored():
    """SURVEY 8(a) probes: re-padding to a larger N (K fixed) and garbage in pad rows of x leave ypred unchanged."""
    e = enc()
    torch.manual_seed(0)
    m = e.SoftPoolingGcnEncoder(40, 3, 16, 16, 4, 3, 16, assign_ratio=0.25).cuda()
    x, adj, nb, label = synth_batch(5, 6, 40, 3, 2, 30, 4)
    with torch.no_grad():
        y0 = m(torch.tensor(x).cuda(), torch.tensor(adj).cuda(), nb).cpu()
        x2 = x.copy()
        for b in range(6):
            x2[b, nb[b]:] = 123.0
        y1 = m(torch.tensor(x2).cuda(), torch.tensor(adj).cuda(), nb).cpu()
        xp = np.zeros((6, 56, 3), np.float32); xp[:, :40] = x
        ap = np.zeros((6, 56, 56), np.float32); ap[:, :40, :40] = adj
        y2 = m(torch.tensor(xp).cuda(), torch.tensor(ap).cuda(), nb).cpu()
    assert torch.equal(y0, y1)
    assert rel_l2(y2.numpy(), y0.numpy()) < 1e-6


def test_training_loop_matches_oracle_for_20_steps():
    """Same initial weights, batches and Adam/clip as train.py:173,209-210: losses track the oracle."""
    e = enc()
    torch.manual_seed(3)
    mo = orc.SoftPoolingGcnEncoder(50, 3, 20, 20, 6, 3, 20, assign_ratio=0.2)
    mc = e.SoftPoolingGcnEncoder(50, 3, 20, 20, 6, 3, 20, assign_ratio=0.2)
    mc.load_state_dict(mo.state_dict())
    mc = mc.cuda()
    oo = torch.optim.Adam(mo.parameters(), lr=1e-3)
    oc = torch.optim.Adam(mc.parameters(), lr=1e-3)
    for step in range(20):
        x, adj, nb, label = synth_batch(100 + step, 8, 50, 3, 3, 50, 6)
        _, lo = orc.train_step(mo, torch.tensor(x), torch.tensor(adj), torch.tensor(label), nb, optimizer=oo)
        mc.zero_grad()
        xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
        yp = mc(xc, ac, nb, assign_x=xc)
        lc = mc.loss(yp, torch.tensor(label).cuda(), ac, nb)
        lc.backward()
        torch.nn.utils.clip_grad_norm_(mc.parameters(), 2.0)
        oc.step()
        assert abs(lc.item() - lo.item()) < 2e-4 * max(1.0, abs(lo.item())), step


def test_soft_large_batch_of_small_graphs_uses_batch_split_bn():
    """ENZYMES-sized graphs in a large batch (B*d >= 16384 with N < 296): the BatchNorm forward / backward kernels
    split each node's batch over several CTAs (Chan-combined statistics) -- same parity bar as everywhere else."""
    oracle_vs_candidate(soft_factory(40, 3, 30, 30, 6, ratio=0.1), 18, 640, 40, 3, 6, nb_mode='rand')
    make = lambda mod: mod.GcnEncoderGraph(5, 36, 20, 3, 3, bn=True)
    oracle_vs_candidate(make, 19, 512, 24, 5, 3, soft=False)


def test_forward_without_backward_does_not_leak():
    """evaluate() in train.py runs forwards that are never back-propagated: the autograd tape must not keep the
    step's buffers alive through a reference cycle (outputs are stored in the tape as detached aliases)."""
    import gc
    m = enc().SoftPoolingGcnEncoder(64, 4, 16, 16, 3, 3, 16, assign_ratio=0.25).cuda()
    x, adj, nb, label = synth_batch(3, 8, 64, 4, 8, 64, 3)
    xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
    for prec in (0, 1):
        m.precision = prec
        m(xc, ac, nb, assign_x=xc)
        torch.cuda.synchronize()
        gc.collect()
        base = torch.cuda.memory_allocated()
        for _ in range(20):
            yp = m(xc, ac, nb, assign_x=xc)
            m.loss(yp, torch.tensor(label).cuda(), ac, nb)
        del yp
        gc.collect()
        torch.cuda.synchronize()
        assert torch.cuda.memory_allocated() <= base + (1 << 20), (prec, torch.cuda.memory_allocated() - base)


@pytest.mark.parametrize('hop,precision,nb_mode', [(2, 0, 'rand'), (3, 0, 'rand'), (2, 0, 'none'), (2, 1, 'rand')])
def test_link_loss_adj_hop(hop, precision, nb_mode):
    """loss(..., adj_hop > 1) (encoders.py:1312-1317, never passed by the reference's callers): the predicted adjacency
    is sum_h (S S^T)^h clamped at 1 (R3) -- the clamp is ACTIVE here, unlike adj_hop = 1.  fp32 rule in fp32 mode; in
    the tensor-core mode the encoder is bf16 and this loss branch runs on the fp32 schedule (bf16 bound)."""
    N, D, H, C, B = 40, 5, 16, 3, 5
    torch.manual_seed(20 + hop)
    mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
    mc = enc().SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
    mc.load_state_dict(mo.state_dict(), strict=True)
    mc = mc.cuda()
    mc.precision = precision
    x, adj, nb, label = synth_batch(21, B, N, D, N if nb_mode == 'none' else 3, N, C, 0.15)
    nbo = None if nb_mode == 'none' else nb
    res = {}
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        m = copy.deepcopy(mo).to(dt)
        xt, at = torch.tensor(x, dtype=dt), torch.tensor(adj, dtype=dt)
        yp = m(xt, at, nbo, assign_x=xt)
        loss = m.loss(yp, torch.tensor(label), at, nbo, adj_hop=hop)
        loss.backward()
        res[tag] = (loss.item(), float(m.link_loss), {k: p.grad.numpy() for k, p in m.named_parameters()})
    xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
    yp = mc(xc, ac, nbo, assign_x=xc)
    loss = mc.loss(yp, lc, ac, nbo, adj_hop=hop)
    loss.backward()
    torch.cuda.synchronize()
    l64, k64, g64 = res['f64']
    cand = {k: p.grad.cpu().numpy() for k, p in mc.named_parameters()}
    if precision == 0:
        assert abs(loss.item() - l64) < OUT_TOL * max(1.0, abs(l64))
        assert abs(mc.link_loss.item() - k64) < OUT_TOL * max(1.0, abs(k64))
        grade_grads(cand, res['f32'][2], g64)
    else:
        assert abs(loss.item() - l64) < 5e-3 * abs(l64)
        fc = np.concatenate([cand[k].ravel() for k in sorted(cand)]).astype(np.float64)
        fo = np.concatenate([g64[k].ravel() for k in sorted(g64)])
        assert rel_l2(fc, fo) < 0.1 and float(fc @ fo / (np.linalg.norm(fc) * np.linalg.norm(fo))) > 0.995
