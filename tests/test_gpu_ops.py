"""GPU: op-level parity of the C-ABI kernels against the oracle's functional pieces (fp64 on CPU).

Tolerance (fp32 path): rel-L2 <= 1e-5 against the fp64 oracle, as BASELINE.json's north_star states.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_l2, synth_batch
from oracle import diffpool_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    return torch.as_tensor(a).cuda().contiguous()


def E():
    from graph_pooling_b200 import engine
    return engine


def st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize('M,N,K,batch,ta,tb,lim', [
    (70, 45, 33, 3, 0, 0, 0), (128, 128, 64, 2, 0, 0, 0), (300, 200, 150, 2, 1, 0, 0), (65, 130, 17, 4, 0, 1, 0),
    (257, 96, 260, 3, 1, 1, 0), (10, 6, 50, 5, 0, 1, 0), (200, 200, 200, 4, 0, 0, 1), (3, 3, 3, 1, 0, 0, 0),
    (512, 384, 256, 2, 0, 0, 0),
])
def test_bgemm_matches_fp64(M, N, K, batch, ta, tb, lim):
    e = E()
    rs = np.random.RandomState(M + N + K)
    A = rs.randn(batch, K, M) if ta else rs.randn(batch, M, K)
    Bm = rs.randn(batch, N, K) if tb else rs.randn(batch, K, N)
    A32, B32 = dev(A.astype(np.float32)), dev(Bm.astype(np.float32))
    out = torch.full((batch, M, N), 7.0, device='cuda')
    sA = (M * K, 1, M) if ta else (M * K, K, 1)
    sB = (N * K, 1, K) if tb else (N * K, N, 1)
    limv = None
    Aop = np.swapaxes(A, 1, 2) if ta else A
    Bop = np.swapaxes(Bm, 1, 2) if tb else Bm
    A64 = Aop.astype(np.float32).astype(np.float64)
    B64 = Bop.astype(np.float32).astype(np.float64)
    if lim:
        l = rs.randint(1, min(M, K) + 1, size=batch).astype(np.int32)
        limv = dev(l)
        ref = np.zeros((batch, M, N))
        for b in range(batch):
            ref[b, :l[b]] = A64[b, :l[b], :l[b]] @ B64[b, :l[b]]
    else:
        ref = A64 @ B64
    e.bgemm(A32.data_ptr(), B32.data_ptr(), out.data_ptr(), M, N, K, batch, sA, sB, (M * N, N, 1),
            lim=None if limv is None else limv.data_ptr(), lim_m=lim, lim_k=lim)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu().numpy(), ref) < TOL


def test_bgemm_splitk_bias_relu_beta():
    e = E()
    rs = np.random.RandomState(0)
    A, Bm = rs.randn(1, 40, 5000).astype(np.float32), rs.randn(1, 5000, 24).astype(np.float32)
    out = torch.zeros(1, 40, 24, device='cuda')
    Ad, Bd = dev(A), dev(Bm)
    e.bgemm(Ad.data_ptr(), Bd.data_ptr(), out.data_ptr(), 40, 24, 5000, 1, (0, 5000, 1), (0, 24, 1),
            (0, 24, 1), split_k=8)
    ref = A.astype(np.float64) @ Bm.astype(np.float64)
    assert rel_l2(out.cpu().numpy(), ref) < TOL
    # bias + relu + beta accumulate
    bias = rs.randn(24).astype(np.float32)
    A2, B2 = rs.randn(2, 40, 30).astype(np.float32), rs.randn(2, 30, 24).astype(np.float32)
    base = rs.randn(2, 40, 24).astype(np.float32)
    out = dev(base.copy())
    a_, b_, bi_ = dev(A2), dev(B2), dev(bias)
    e.bgemm(a_.data_ptr(), b_.data_ptr(), out.data_ptr(), 40, 24, 30, 2, (1200, 30, 1), (720, 24, 1), (960, 24, 1),
            bias=bi_.data_ptr(), relu=1, beta=1.0, alpha=0.5)
    ref = np.maximum(0.5 * (A2.astype(np.float64) @ B2) + bias, 0) + base
    assert rel_l2(out.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize('B,N,din,dout,add_self,use_nb', [(4, 50, 3, 30, 0, 1), (3, 64, 30, 10, 1, 1),
                                                          (2, 130, 89, 20, 0, 0), (5, 17, 8, 8, 0, 1)])
def test_graphconv_fwd_bwd(B, N, din, dout, add_self, use_nb):
    from graph_pooling_b200 import encoders
    x, adj, nb, _ = synth_batch(11, B, N, din, 1, N, 2, density=0.2, symmetric=False)
    rs = np.random.RandomState(5)
    w = (rs.randn(din, dout) * 0.3).astype(np.float32)
    bias = (rs.randn(dout) * 0.2).astype(np.float32)
    gy = rs.randn(B, N, dout).astype(np.float32)
    # oracle fp64
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    at = torch.tensor(adj, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
    yo = orc.graph_conv(xt, at, wt, bt, add_self=bool(add_self), normalize=True)
    yo.backward(torch.tensor(gy, dtype=torch.float64))
    # candidate (module-level GraphConv Function through the C ABI)
    xc, ac = dev(x).requires_grad_(), dev(adj).requires_grad_()
    wc, bc = dev(w).requires_grad_(), dev(bias).requires_grad_()
    yc = encoders._GraphConvFn.apply(xc, ac, wc, bc, bool(add_self), True)
    yc.backward(dev(gy))
    torch.cuda.synchronize()
    assert rel_l2(yc.detach().cpu().numpy(), yo.detach().numpy()) < TOL
    for a, b, name in ((xc, xt, 'dx'), (ac, at, 'dadj'), (wc, wt, 'dw'), (bc, bt, 'db')):
        assert rel_l2(a.grad.cpu().numpy(), b.grad.numpy()) < 2e-5, name


@pytest.mark.parametrize('B,N,d,ldh', [(6, 40, 30, 90), (3, 10, 7, 7), (20, 100, 30, 70), (2, 3, 1, 4)])
def test_relu_bn_fwd(B, N, d, ldh):
    from graph_pooling_b200._lib import call
    rs = np.random.RandomState(1)
    y = rs.randn(B, N, d).astype(np.float32)
    ref = orc.bn_per_node(torch.relu(torch.tensor(y, dtype=torch.float64))).numpy()
    yc = dev(y)
    h = torch.zeros(B, N, ldh, device='cuda')
    mean, invstd = torch.empty(N, device='cuda'), torch.empty(N, device='cuda')
    call('gp_relu_bn_fwd', yc.data_ptr(), h.data_ptr(), ldh, mean.data_ptr(), invstd.data_ptr(), B, N, d, 1, 1, st())
    torch.cuda.synchronize()
    assert rel_l2(h[:, :, :d].cpu().numpy(), ref) < TOL
    assert float(h[:, :, d:].abs().sum()) == 0.0


def test_readout_max_ties_and_mask():
    from graph_pooling_b200._lib import call
    B, N, F = 3, 9, 40
    rs = np.random.RandomState(2)
    z = rs.randn(B, N, F).astype(np.float32)
    z[0, 2, :] = z[0, 5, :] = 10.0            # tie -> lowest index (2)
    z[1, :, 3] = -1.0                          # all negative: masked pad row (0) wins when nb < N
    nb = np.array([9, 4, 1], np.int32)
    zc, nbc = dev(z), dev(nb)
    out = torch.empty(B, F, device='cuda')
    arg = torch.empty(B, F, device='cuda', dtype=torch.int32)
    call('gp_readout_max_fwd', zc.data_ptr(), F, nbc.data_ptr(), B, N, F, out.data_ptr(), arg.data_ptr(), F, st())
    m = orc.construct_mask(N, nb, 'cpu', torch.float32)
    ref, _ = torch.max(torch.tensor(z) * m, dim=1)
    assert torch.equal(out.cpu(), ref)
    arg = arg.cpu().numpy()
    assert (arg[0] == 2).all()
    assert arg[1, 3] == -1 and out[1, 3].item() == 0.0
    # unmasked
    arg2 = dev(np.zeros((B, F), np.int32))
    call('gp_readout_max_fwd', zc.data_ptr(), F, None, B, N, F, out.data_ptr(), arg2.data_ptr(), F, st())
    assert torch.equal(out.cpu(), torch.max(torch.tensor(z), dim=1)[0])


def test_readout_max_row_limit_and_pad_helpers():
    """gp_readout_max_fwd_x scans the first N rows of graphs stored `pitch` rows apart (dead clusters of a padded
    assignment width are not part of the readout); gp_pad_copy_f32 / gp_fill_i32 build the padded parameter copies."""
    from graph_pooling_b200._lib import call
    B, N, pitch, F = 3, 10, 16, 40
    rs = np.random.RandomState(5)
    z = rs.randn(B, pitch, F).astype(np.float32)
    z[:, N:, :] = 50.0                          # rows beyond N must never win
    z[1, :N, 7] = -2.0                          # all real rows negative: the max stays negative (no zero pad row)
    zc = dev(z)
    out = torch.empty(B, F, device='cuda')
    arg = torch.empty(B, F, device='cuda', dtype=torch.int32)
    call('gp_readout_max_fwd_x', zc.data_ptr(), F, pitch, None, B, N, F, out.data_ptr(), arg.data_ptr(), F, st())
    ref, idx = torch.max(torch.tensor(z[:, :N]), dim=1)
    assert torch.equal(out.cpu(), ref) and out[1, 7].item() == -2.0
    assert np.array_equal(arg.cpu().numpy(), idx.numpy().astype(np.int32))
    src = dev(rs.randn(5, 9).astype(np.float32))
    dst = torch.full((8, 12), 3.0, device='cuda')
    call('gp_pad_copy_f32', src.data_ptr(), C.c_longlong(9), C.c_longlong(5), 7, dst.data_ptr(), C.c_longlong(12),
         C.c_longlong(6), 10, C.c_float(-1.5), st())
    d = dst.cpu().numpy()
    assert np.array_equal(d[:5, :7], src.cpu().numpy()[:, :7])
    assert (d[5, :10] == -1.5).all() and (d[:5, 7:10] == -1.5).all()
    assert (d[6:] == 3.0).all() and (d[:, 10:] == 3.0).all()      # outside [rows_dst, cols_dst]: untouched
    iv = torch.zeros(11, device='cuda', dtype=torch.int32)
    call('gp_fill_i32', iv.data_ptr(), C.c_longlong(10), 250, st())
    assert iv.cpu().tolist() == [250] * 10 + [0]


@pytest.mark.parametrize('K', [1, 10, 33, 512])
def test_softmax_mask_fwd_bwd(K):
    from graph_pooling_b200._lib import call
    B, N = 4, 13
    rs = np.random.RandomState(K)
    t = (rs.randn(B, N, K) * 3).astype(np.float32)
    ds = rs.randn(B, N, K).astype(np.float32)
    nb = np.array([13, 1, 7, 12], np.int32)
    tt = torch.tensor(t, dtype=torch.float64, requires_grad=True)
    m = orc.construct_mask(N, nb, 'cpu', torch.float64)
    so = torch.softmax(tt, dim=-1) * m
    so.backward(torch.tensor(ds, dtype=torch.float64))
    s, nbc, dsc = dev(t.copy()), dev(nb), dev(ds)
    dt = torch.empty_like(s)
    call('gp_softmax_mask_fwd', s.data_ptr(), nbc.data_ptr(), B, N, K, st())
    call('gp_softmax_mask_bwd', s.data_ptr(), dsc.data_ptr(), nbc.data_ptr(), B, N, K, dt.data_ptr(), st())
    torch.cuda.synchronize()
    assert rel_l2(s.cpu().numpy(), so.detach().numpy()) < TOL
    assert rel_l2(dt.cpu().numpy(), tt.grad.numpy()) < TOL


@pytest.mark.parametrize('B,N,K,F,use_nb,sym', [(3, 50, 12, 20, 1, 0), (2, 130, 32, 90, 1, 1), (2, 24, 6, 8, 0, 0),
                                                   (20, 100, 10, 90, 1, 1), (5, 33, 7, 13, 1, 0), (3, 128, 32, 96, 0, 1)])
def test_pool_fwd_bwd(B, N, K, F, use_nb, sym):
    from graph_pooling_b200._lib import call
    _, adj, nb, _ = synth_batch(3, B, N, 2, 2, N, 2, density=0.2, symmetric=bool(sym))
    if not use_nb:
        nb = np.full(B, N, np.int32)
    rs = np.random.RandomState(4)
    m = orc.construct_mask(N, nb, 'cpu', torch.float64)
    s_np = rs.rand(B, N, K) * m.numpy()
    z_np = rs.randn(B, N, F)
    s_np, z_np = s_np.astype(np.float32), z_np.astype(np.float32)
    gx, ga = rs.randn(B, K, F).astype(np.float32), rs.randn(B, K, K).astype(np.float32)
    s_t = torch.tensor(s_np, dtype=torch.float64, requires_grad=True)
    z_t = torch.tensor(z_np, dtype=torch.float64, requires_grad=True)
    xo, ao = orc.pool(s_t, z_t * m, torch.tensor(adj, dtype=torch.float64))
    (xo * torch.tensor(gx, dtype=torch.float64)).sum().backward(retain_graph=True)
    (ao * torch.tensor(ga, dtype=torch.float64)).sum().backward()
    sc, zc, ac = dev(s_np), dev(z_np), dev(adj)
    nbc = dev(nb) if use_nb else None
    xp, t, ap = (torch.empty(B, K, F, device='cuda'), torch.empty(B, K, N, device='cuda'),
                 torch.empty(B, K, K, device='cuda'))
    nbp = None if nbc is None else nbc.data_ptr()
    call('gp_pool_fwd', sc.data_ptr(), zc.data_ptr(), F, ac.data_ptr(), nbp, B, N, K, F, xp.data_ptr(), t.data_ptr(),
         ap.data_ptr(), 0, st())
    dz, dsb, ws = (torch.empty(B, N, F, device='cuda'), torch.empty(B, N, K, device='cuda'),
                   torch.empty(B, N, K, device='cuda'))
    gxc, gac = dev(gx), dev(ga)
    call('gp_pool_bwd', gxc.data_ptr(), gac.data_ptr(), sc.data_ptr(), zc.data_ptr(), F, ac.data_ptr(),
         t.data_ptr(), nbp, B, N, K, F, dz.data_ptr(), F, 0, dsb.data_ptr(), 0, None, ws.data_ptr(), 0, st())
    torch.cuda.synchronize()
    assert rel_l2(xp.cpu().numpy(), xo.detach().numpy()) < TOL
    assert rel_l2(ap.cpu().numpy(), ao.detach().numpy()) < TOL
    assert rel_l2(dz.cpu().numpy(), z_t.grad.numpy()) < TOL
    mm = m.numpy()
    assert rel_l2(dsb.cpu().numpy() * mm, s_t.grad.numpy() * mm) < TOL     # pad rows of dS are masked downstream
    # T = S^T A is an output of the forward (the backward reads it); pad columns are zero
    t_ref = torch.tensor(s_np, dtype=torch.float64).transpose(1, 2) @ torch.tensor(adj, dtype=torch.float64)
    assert rel_l2(t.cpu().numpy(), t_ref.numpy()) < TOL
    # accumulate flags: dZ / dS are added to what the buffers hold (real rows; pad rows are left alone)
    dz2, ds2 = torch.ones_like(dz), torch.ones_like(dsb)
    call('gp_pool_bwd', gxc.data_ptr(), gac.data_ptr(), sc.data_ptr(), zc.data_ptr(), F, ac.data_ptr(),
         t.data_ptr(), nbp, B, N, K, F, dz2.data_ptr(), F, 1, ds2.data_ptr(), 1, None, ws.data_ptr(), 0, st())
    torch.cuda.synchronize()
    assert rel_l2((dz2 - 1).cpu().numpy(), dz.cpu().numpy()) < 1e-5
    assert rel_l2((ds2 - 1).cpu().numpy() * mm, dsb.cpu().numpy() * mm) < 1e-5


@pytest.mark.parametrize('B,N,K,use_nb,sym', [(3, 70, 10, 1, 1), (2, 130, 33, 1, 0), (2, 20, 4, 0, 1)])
def test_linkloss_fwd_bwd(B, N, K, use_nb, sym):
    from graph_pooling_b200._lib import call
    e = E()
    _, adj, nb, _ = synth_batch(8, B, N, 2, 1, N, 2, density=0.15, symmetric=bool(sym))
    nbo = nb if use_nb else None
    rs = np.random.RandomState(6)
    m = orc.construct_mask(N, nb if use_nb else np.full(B, N), 'cpu', torch.float64)
    s_np = (torch.softmax(torch.tensor(rs.randn(B, N, K) * 2), dim=-1) * m).numpy().astype(np.float32)
    s_t = torch.tensor(s_np, dtype=torch.float64, requires_grad=True)
    lo = orc.link_pred_loss(s_t, torch.tensor(adj, dtype=torch.float64), nbo)
    lo.backward()
    sc, ac = dev(s_np), dev(adj)
    nbc = dev(nb) if use_nb else None
    nbp = None if nbc is None else nbc.data_ptr()
    T = (N + 63) // 64
    partial = torch.empty(B * T * T + 256, device='cuda')
    gsym = torch.empty(B, N, N, device='cuda')
    call('gp_linkloss_fwd', sc.data_ptr(), ac.data_ptr(), nbp, B, N, K, partial.data_ptr(), gsym.data_ptr(), st())
    entries = float(np.sum(nb.astype(np.int64) ** 2)) if use_nb else float(B * N * N)
    total, link = torch.empty(1, device='cuda'), torch.empty(1, device='cuda')
    ce = torch.full((1,), 0.25, device='cuda')
    call('gp_loss_finalize', partial.data_ptr(), B * T * T, C.c_double(1.0 / entries), ce.data_ptr(),
         total.data_ptr(), link.data_ptr(), st())
    ds = torch.empty(B, N, K, device='cuda')
    e.bgemm(gsym.data_ptr(), sc.data_ptr(), ds.data_ptr(), N, K, N, B, (N * N, N, 1), (N * K, K, 1), (N * K, K, 1),
            lim=nbp, lim_m=int(use_nb), lim_k=int(use_nb), alpha=1.0 / entries)
    torch.cuda.synchronize()
    assert abs(link.item() - lo.item()) < TOL * abs(lo.item())
    assert abs(total.item() - (lo.item() + 0.25)) < 2e-6 * (abs(lo.item()) + 1)
    mm = m.numpy()
    assert rel_l2(ds.cpu().numpy() * mm, s_t.grad.numpy() * mm) < 2e-5


def test_ce_fwd_bwd():
    from graph_pooling_b200._lib import call
    rs = np.random.RandomState(9)
    B, Cc = 20, 6
    lg = (rs.randn(B, Cc) * 2).astype(np.float32)
    lab = rs.randint(0, Cc, size=B).astype(np.int64)
    lt = torch.tensor(lg, dtype=torch.float64, requires_grad=True)
    lo = torch.nn.functional.cross_entropy(lt, torch.tensor(lab))
    (lo * 1.5).backward()
    lc, labc = dev(lg), dev(lab)
    loss, probs, dl = torch.empty(1, device='cuda'), torch.empty(B, Cc, device='cuda'), torch.empty(B, Cc, device='cuda')
    up = torch.full((1,), 1.5, device='cuda')
    call('gp_ce_fwd', lc.data_ptr(), labc.data_ptr(), B, Cc, loss.data_ptr(), probs.data_ptr(), st())
    call('gp_ce_bwd', probs.data_ptr(), labc.data_ptr(), up.data_ptr(), B, Cc, dl.data_ptr(), st())
    torch.cuda.synchronize()
    assert abs(loss.item() - lo.item()) < TOL
    assert rel_l2(dl.cpu().numpy(), lt.grad.numpy()) < TOL
