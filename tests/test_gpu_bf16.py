"""GPU: the GP_BF16 tensor-core mode of the drop-in encoders against the fp64 oracle.

Stated bound of this mode (bf16 operands, fp32 accumulation; BASELINE.json north_star "bf16
tensor-core path with its own stated bound"):
  ypred: rel-L2 <= 1.5e-2;  S: rel-L2 <= 5e-3;  loss: relative <= 2.5e-3;
  full flattened parameter gradient: rel-L2 <= 0.09 and cosine similarity >= 0.995
-- about twice the worst value measured over the cases of this file on a B200 (ypred 7.2e-3, S 2.4e-3, loss 1.0e-3,
gradient 0.044 / cosine 0.999: profiles/r2_bf16_errors.md; round 1 tested 0.15 / 0.99 without recording them).
At the BASELINE shapes the bound is tighter still (tests/test_gpu_baseline_shapes.py).
(The fp32 mode's bound is 1e-5, tests/test_gpu_model.py.)"""

BF16_OUT, BF16_S, BF16_LOSS, BF16_GRAD, BF16_COS = 1.5e-2, 5e-3, 2.5e-3, 0.09, 0.995
import copy
import json
import os

import numpy as np
import pytest
import torch

from helpers import rel_l2, synth_batch
from oracle import diffpool_oracle as orc

pytestmark = pytest.mark.gpu


def _record(**kw):
    """Measured errors of the small-shape cases -> gpurun_out/r2_bf16_small_errors.jsonl (profiles/r2_bf16_errors.md)."""
    try:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, 'r2_bf16_small_errors.jsonl'), 'a') as f:
            f.write(json.dumps(kw) + '\n')
    except OSError:
        pass


@pytest.mark.parametrize('N,D,H,C,B,ratio,P,n_min,density', [
    (64, 8, 16, 3, 4, 0.25, 1, 16, 0.1), (256, 16, 32, 2, 4, 0.25, 1, 64, 0.05),
    (100, 3, 30, 6, 20, 0.1, 1, 2, 0.08), (512, 64, 64, 2, 3, 0.25, 1, 128, 0.02),
    (200, 89, 24, 2, 3, 0.25, 2, 50, 0.05),
    # hidden width on the vectorised layer-backward path but concat strides that miss its alignment
    # (Fa = 2*64+30 = 158, the cfg3 / cfg5 situation: K = 250 / 1250)
    (120, 10, 64, 2, 3, 0.25, 2, 30, 0.05), (600, 12, 32, 2, 2, 0.25, 1, 40, 0.02)])
def test_bf16_mode_bounds(N, D, H, C, B, ratio, P, n_min, density):
    from graph_pooling_b200 import encoders
    torch.manual_seed(N)
    mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio, num_pooling=P)
    g = torch.Generator().manual_seed(N + 1)
    with torch.no_grad():
        for k, p in mo.named_parameters():
            if k.endswith('bias'):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio, num_pooling=P)
    mc.load_state_dict(mo.state_dict())
    mc = mc.cuda()
    mc.precision = 1
    x, adj, nb, label = synth_batch(N, B, N, D, n_min, N, C, density)
    m64 = copy.deepcopy(mo).double()
    yo, lo = orc.train_step(m64, torch.tensor(x).double(), torch.tensor(adj).double(), torch.tensor(label), nb)
    xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
    yp = mc(xc, ac, nb, assign_x=xc)
    loss = mc.loss(yp, torch.tensor(label).cuda(), ac, nb)
    loss.backward()
    torch.cuda.synchronize()
    assert rel_l2(yp.detach().cpu().numpy(), yo.detach().numpy()) < BF16_OUT
    assert rel_l2(mc.assign_tensors[0].detach().cpu().numpy(), m64.assign_tensors[0].detach().numpy()) < BF16_S
    assert abs(loss.item() - lo.item()) < BF16_LOSS * abs(lo.item())
    gc = np.concatenate([p.grad.cpu().numpy().ravel() for p in mc.parameters()]).astype(np.float64)
    go = np.concatenate([p.grad.numpy().ravel() for p in m64.parameters()])
    cos = float(gc @ go / (np.linalg.norm(gc) * np.linalg.norm(go)))
    _record(case='small N=%d D=%d H=%d B=%d P=%d' % (N, D, H, B, P), ypred=rel_l2(yp.detach().cpu().numpy(), yo.detach().numpy()),
            S=rel_l2(mc.assign_tensors[0].detach().cpu().numpy(), m64.assign_tensors[0].detach().numpy()),
            loss=abs(loss.item() - lo.item()) / abs(lo.item()), grad_flat=rel_l2(gc, go), grad_cos=cos)
    assert rel_l2(gc, go) < BF16_GRAD and cos > BF16_COS, (rel_l2(gc, go), cos)


@pytest.mark.parametrize('N,H,ratio,P,extra', [(200, 24, 0.25, 2, 0), (120, 64, 0.25, 2, 1), (600, 32, 0.25, 1, 1),
                                               (100, 30, 0.1, 1, 0)])
def test_padded_cluster_count_is_exact(monkeypatch, N, H, ratio, P, extra):
    """Cluster counts that are not multiples of 8 (cfg3: K = 250 / 62, cfg5: K = 1250) run at r8(K) with DEAD
    clusters (zero weights, logit bias -1e30).  The dead clusters must change nothing: the padded schedule is compared
    with the unpadded one (GP_NO_KPAD=1, same bf16 operands) -- outputs, S (shape and values), losses and every
    parameter gradient.  extra=1 adds a user loss on the level-0 assignment tensor (its gradient reaches the encoder through
    the padded-copy path instead of the link loss's own padded buffer)."""
    from graph_pooling_b200 import encoders
    D, C, B = 12, 3, 3
    torch.manual_seed(7 * N)
    m = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio, num_pooling=P).cuda()
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith('bias'):
                p.copy_(0.2 * torch.randn(p.shape, device='cuda'))
    m.precision = 1
    assert any(k % 8 for k in m.assign_dims)
    x, adj, nb, label = synth_batch(N + 3, B, N, D, max(N // 4, 2), N, C, 0.05)
    xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
    res = []
    for nopad in (0, 1):
        if nopad:
            monkeypatch.setenv('GP_NO_KPAD', '1')
        m.zero_grad(set_to_none=True)
        yp = m(xc, ac, nb, assign_x=xc)
        loss = m.loss(yp, lc, ac, nb)
        tot = loss + (0.01 * (m.assign_tensors[0] ** 2).sum() if extra else 0.0)
        tot.backward()
        torch.cuda.synchronize()
        assert [tuple(t.shape[1:]) for t in m.assign_tensors] == \
            [(N if i == 0 else m.assign_dims[i - 1], m.assign_dims[i]) for i in range(P)]
        res.append((yp.detach().cpu().numpy(), [t.detach().cpu().numpy() for t in m.assign_tensors], loss.item(),
                    m.link_loss.item(), {k: p.grad.cpu().numpy().copy() for k, p in m.named_parameters()}))
    (y0, s0, l0, ll0, g0), (y1, s1, l1, ll1, g1) = res
    # same arithmetic on the real clusters; only the order of a few fp32 reductions (vector vs scalar row kernels,
    # split-K shapes) differs, plus the bf16 roundings that those last-bit differences can flip
    assert rel_l2(y0, y1) < 2e-3
    for a, b in zip(s0, s1):
        assert rel_l2(a, b) < 2e-3
    assert abs(l0 - l1) < 1e-4 * abs(l1) and abs(ll0 - ll1) < 1e-4 * abs(ll1)
    f0 = np.concatenate([g0[k].ravel() for k in sorted(g0)]).astype(np.float64)
    f1 = np.concatenate([g1[k].ravel() for k in sorted(g1)]).astype(np.float64)
    assert rel_l2(f0, f1) < 2e-2, rel_l2(f0, f1)
    for k in g0:
        assert g0[k].shape == g1[k].shape and np.isfinite(g0[k]).all(), k


@pytest.mark.parametrize('B,N,sym,weighted,use_nb,u8', [(3, 200, 1, 0, 1, 0), (2, 130, 0, 0, 1, 0), (2, 64, 1, 1, 0, 0),
                                                       (3, 257, 1, 0, 1, 1), (2, 96, 0, 0, 0, 1)])
def test_adj_prepare_flags_and_values(B, N, sym, weighted, use_nb, u8):
    """gp_adj_prepare: exact bf16 copy (zero padded), 'not symmetric' / 'not {0,1}' flags, fp32 and uint8 feeds."""
    from graph_pooling_b200 import engine as E, engine_tc as T
    _, adj, nb, _ = synth_batch(N + sym, B, N, 2, N // 2, N, 2, 0.2, symmetric=bool(sym), weighted=bool(weighted))
    a = torch.tensor(adj).cuda()
    if u8:
        a = a.to(torch.uint8)
    nbd = torch.tensor(nb).cuda() if use_nb else None
    op, flags = T.adj_prepare(E.Workspace(a.device), a, nbd, B, N)
    torch.cuda.synchronize()
    out = op.t.float().cpu().numpy()
    assert out.shape == (B, N, (N + 31) // 32 * 32)     # adjacency operands: rows padded to 32 elements (row epilogue)
    assert np.array_equal(out[:, :, :N], torch.tensor(adj).bfloat16().float().numpy()) and not out[:, :, N:].any()
    assert flags.cpu().tolist() == [int(not sym), int(bool(weighted))]


def test_bf16_mode_nonsymmetric_and_u8_feed():
    """Non-symmetric adjacency takes the general (two-product) backward; a uint8 adjacency feed gives the same
    result as the fp32 feed bit for bit (both become the same bf16 operand)."""
    from graph_pooling_b200 import encoders
    N, D, H, C, B = 192, 8, 32, 3, 3
    outs = {}
    for sym in (0, 1):
        torch.manual_seed(7)
        mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
        mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
        mc.load_state_dict(mo.state_dict())
        mc = mc.cuda()
        mc.precision = 1
        x, adj, nb, label = synth_batch(31, B, N, D, 60, N, C, 0.06, symmetric=bool(sym))
        m64 = copy.deepcopy(mo).double()
        yo, lo = orc.train_step(m64, torch.tensor(x).double(), torch.tensor(adj).double(), torch.tensor(label), nb)
        go = np.concatenate([p.grad.numpy().ravel() for p in m64.parameters()])
        for feed in ('f32', 'u8'):
            mc.zero_grad()
            xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
            if feed == 'u8':
                ac = ac.to(torch.uint8)
            yp = mc(xc, ac, nb, assign_x=xc)
            loss = mc.loss(yp, torch.tensor(label).cuda(), ac, nb)
            loss.backward()
            torch.cuda.synchronize()
            gc = np.concatenate([p.grad.cpu().numpy().ravel() for p in mc.parameters()]).astype(np.float64)
            cos = float(gc @ go / (np.linalg.norm(gc) * np.linalg.norm(go)))
            assert abs(loss.item() - lo.item()) < 5e-3 * abs(lo.item())
            assert rel_l2(gc, go) < 0.15 and cos > 0.99, (sym, feed, rel_l2(gc, go), cos)
            outs[(sym, feed)] = (yp.detach().cpu().numpy(), gc)
        assert np.array_equal(outs[(sym, 'f32')][0], outs[(sym, 'u8')][0])
        # gradients: the split-K weight-gradient GEMMs accumulate with atomics (order varies run to run)
        assert rel_l2(outs[(sym, 'u8')][1], outs[(sym, 'f32')][1]) < 1e-5


def test_dual_stack_matches_separate_stacks(monkeypatch):
    """Lock-step embedding + assignment GCN (one A.X / A^T.dU pass per layer for both) must give the same result
    as running the two stacks separately (GP_NO_DUAL=1): same kernels, same operands, only the launch grouping
    differs -- identical forward, gradients equal up to the split-K atomics' summation order.  The lock-step backward
    also keeps three gradient intermediates (dZ, dza, the per-layer dX) in bf16: switched off for the like-for-like
    comparison (GP_F32_GRADS=1) and measured separately -- 1.4e-4 of the gradient norm, two orders below the mode's
    own error."""
    from graph_pooling_b200 import encoders
    N, D, H, C, B = 160, 12, 32, 3, 4
    torch.manual_seed(3)
    mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25).cuda()
    mc.precision = 1
    with torch.no_grad():
        for k, p in mc.named_parameters():
            if k.endswith('bias'):
                p.normal_(0, 0.2)
    x, adj, nb, label = synth_batch(5, B, N, D, 40, N, C, 0.06)
    res = {}
    for mode, same_x in (('dual', True), ('sep', True), ('dual', False), ('sep', False), ('dual16', True)):
        if mode == 'sep':
            monkeypatch.setenv('GP_NO_DUAL', '1')
        else:
            monkeypatch.delenv('GP_NO_DUAL', raising=False)
        if mode == 'dual16':
            monkeypatch.delenv('GP_F32_GRADS', raising=False)
        else:
            monkeypatch.setenv('GP_F32_GRADS', '1')
        mc.zero_grad()
        xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
        xa = xc if same_x else (xc * 0.5 + 0.1).contiguous()
        yp = mc(xc, ac, nb, assign_x=xa)
        loss = mc.loss(yp, torch.tensor(label).cuda(), ac, nb)
        loss.backward()
        torch.cuda.synchronize()
        res[(mode, same_x)] = (yp.detach().cpu().numpy(), loss.item(),
                               np.concatenate([p.grad.cpu().numpy().ravel() for p in mc.parameters()]))
    for same_x in (True, False):
        a, b = res[('dual', same_x)], res[('sep', same_x)]
        assert np.array_equal(a[0], b[0]) and a[1] == b[1]
        assert rel_l2(a[2], b[2]) < 1e-5
    a, b = res[('dual16', True)], res[('dual', True)]      # bf16 gradient intermediates: same forward, gradient within 1e-3
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    assert rel_l2(a[2], b[2]) < 1e-3


@pytest.mark.parametrize('B,N,H,ratio,n_min', [(2, 2048, 128, 0.25, 2048), (2, 2048, 128, 0.25, 700), (1, 5000, 64, 0.25, 3000)])
def test_bf16_mode_against_fp32_mode_at_full_size(B, N, H, ratio, n_min):
    """BASELINE.json's full sizes (2048-node dense graphs, K = 512; a 5000-node ragged graph, K = 1250) are beyond what
    the CPU oracle finishes in seconds: there the tensor-core mode is checked against this library's own fp32 mode
    (itself pinned to the oracle at small sizes) on the same inputs and weights -- every tile boundary, the
    upper-band link-loss schedule, the symmetric shortcuts and the padding-aware skips at their real extents."""
    from graph_pooling_b200 import encoders
    D, C = H, 2
    torch.manual_seed(N + B)
    m = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=ratio).cuda()
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith('bias'):
                p.normal_(0, 0.2)
    g = torch.Generator(device='cuda').manual_seed(1)
    nb = np.random.RandomState(N).randint(n_min, N + 1, size=B).astype(np.int32)
    real = torch.arange(N, device='cuda')[None, :] < torch.as_tensor(nb.astype(np.int64), device='cuda')[:, None]
    u = torch.triu(torch.rand(B, N, N, device='cuda', generator=g) < 8.0 / N, diagonal=1)
    u = u & real[:, None, :] & real[:, :, None]
    adj = (u | u.transpose(1, 2)).float()
    x = torch.randn(B, N, D, device='cuda', generator=g) * real[:, :, None].float()
    label = torch.randint(0, C, (B,), device='cuda', generator=g)
    out = {}
    for prec in (0, 1):
        m.precision = prec
        m.zero_grad()
        yp = m(x, adj, nb, assign_x=x)
        loss = m.loss(yp, label, adj, nb)
        loss.backward()
        torch.cuda.synchronize()
        out[prec] = (yp.detach().float().cpu().numpy(), loss.item(), m.link_loss.item(),
                     m.assign_tensor.detach().cpu().numpy(),
                     np.concatenate([p.grad.cpu().numpy().ravel() for p in m.parameters()]).astype(np.float64))
        del yp, loss
    y0, l0, k0, s0, g0 = out[0]
    y1, l1, k1, s1, g1 = out[1]
    assert rel_l2(y1, y0) < 2e-2 and rel_l2(s1, s0) < 2e-2
    assert abs(l1 - l0) < 5e-3 * abs(l0) and abs(k1 - k0) < 5e-3 * abs(k0)
    cos = float(g1 @ g0 / (np.linalg.norm(g1) * np.linalg.norm(g0)))
    assert rel_l2(g1, g0) < 0.15 and cos > 0.99, (rel_l2(g1, g0), cos)


@pytest.mark.parametrize('B,N,sym', [(4, 200, 1), (3, 130, 0), (5, 64, 1)])
def test_prepared_adjacency_from_bits_and_fp32_parts(B, N, sym):
    """Feed-side adjacency: part of the batch bit-packed on the host (gp_host_pack_adj_bits), part as fp32, both
    expanded by gp_adj_prepare_x into one bf16 operand -- identical operand, flags and model outputs as the plain
    fp32 adjacency."""
    import ctypes as C
    from graph_pooling_b200 import _lib, encoders, engine as E, engine_tc as T
    x, adj, nb, label = synth_batch(N + B, B, N, 6, N // 2, N, 3, 0.1, symmetric=bool(sym))
    ldb = (N + 7) // 8
    bits = np.zeros((B, N, ldb), np.uint8)
    bad = C.c_int(0)
    _lib.load().gp_host_pack_adj_bits(adj.ctypes.data, B * N, N, bits.ctypes.data, ldb, 4, C.addressof(bad))
    assert bad.value == 0
    nbd = torch.tensor(nb).cuda()
    ref_op, ref_flags = T.adj_prepare(E.Workspace(torch.device('cuda')), torch.tensor(adj).cuda(), nbd, B, N)
    h = B // 2
    pa = T.PreparedAdjacency(B, N, 'cuda')
    pa.add(torch.tensor(bits[:h]).cuda(), 0, nbd, 'bits')
    pa.add(torch.tensor(adj[h:]).cuda(), h, nbd, 'f32')
    torch.cuda.synchronize()
    assert torch.equal(pa.op.t, ref_op.t) and pa.flags.tolist() == ref_flags.tolist() == [int(not sym), 0]
    torch.manual_seed(0)
    m = encoders.SoftPoolingGcnEncoder(N, 6, 32, 32, 3, 3, 32, assign_ratio=0.25).cuda()
    m.precision = 1
    xc, lc = torch.tensor(x).cuda(), torch.tensor(label).cuda()
    res = []
    for a in (torch.tensor(adj).cuda(), pa):
        m.zero_grad()
        yp = m(xc, a, nb, assign_x=xc)
        loss = m.loss(yp, lc, a, nb)
        loss.backward()
        torch.cuda.synchronize()
        res.append((yp.detach().cpu().numpy(), loss.item()))
        del yp, loss
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]


def test_host_adjacency_feed_pipeline():
    """feed.HostAdjacencyFeed: fp32 adjacencies in pinned host memory, a different one per step, go through the
    pack / copy / prepare pipeline and arrive as exactly the operand gp_adj_prepare builds from the fp32 tensor."""
    from graph_pooling_b200 import engine as E, engine_tc as T, feed
    B, N = 6, 150
    batches = [synth_batch(70 + i, B, N, 4, 40, N, 2, 0.1) for i in range(4)]
    host = [torch.tensor(b[1]).pin_memory() for b in batches]
    fd = feed.HostAdjacencyFeed(B, N, 'cuda', packed_graphs=4, threads=3)
    fd.submit(host[0], 0)
    fd.copy(0)
    fd.submit(host[1], 1)
    for i in range(4):
        cur, nxt = i & 1, (i + 1) & 1
        if i + 1 < 4:
            fd.copy(nxt)
        if i + 2 < 4:
            fd.submit(host[i + 2], cur)
        nbd = torch.tensor(batches[i][2]).cuda()
        pa = fd.prepared(cur, nbd)
        ref, rflags = T.adj_prepare(E.Workspace(torch.device('cuda')), host[i].cuda(), nbd, B, N)
        torch.cuda.synchronize()
        assert torch.equal(pa.op.t, ref.t) and pa.flags.tolist() == rflags.tolist(), i
    bad = host[0].clone().pin_memory()
    bad[1, 3, 4] = 0.25
    fd.submit(bad, 0)
    with pytest.raises(ValueError):
        fd.copy(0)
    fd.close()


def test_bf16_gradient_intermediates_at_a_shape_that_takes_them(monkeypatch):
    """The lock-step backward keeps dZ / dza / dX in bf16 only where every layer is served by a layer-backward kernel
    instantiated for bf16 sources (gp_gcn_layer_bwd_bf16_sources_fast): B = 128 graphs, H = 128, K = 128 is such a
    shape.  Checks that the path is really taken (hook), that it stays within the mode's bounds against the fp64 oracle
    and within 1e-3 of the fp32-intermediate schedule."""
    from graph_pooling_b200 import encoders, _lib
    N, D, H, C, B = 512, 16, 128, 3, 128
    torch.manual_seed(11)
    mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25, num_pooling=1)
    mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25, num_pooling=1)
    mc.load_state_dict(mo.state_dict())
    mc = mc.cuda()
    mc.precision = 1
    x, adj, nb, label = synth_batch(17, B, N, D, 100, N, C, 0.02)
    m64 = copy.deepcopy(mo).double()
    yo, lo = orc.train_step(m64, torch.tensor(x).double(), torch.tensor(adj).double(), torch.tensor(label), nb)
    go = np.concatenate([p.grad.numpy().ravel() for p in m64.parameters()])

    class Rec:
        def __init__(self):
            self.bf16_calls = 0

        def begin(self, name, args):
            if name == 'gp_gcn_layer_bwd_x' and args[0]._obj.dz_bf16:
                self.bf16_calls += 1

        def end(self, tok):
            pass

    grads = {}
    for f32 in (False, True):
        if f32:
            monkeypatch.setenv('GP_F32_GRADS', '1')
        else:
            monkeypatch.delenv('GP_F32_GRADS', raising=False)
        mc.zero_grad()
        rec = Rec()
        monkeypatch.setattr(_lib, '_hook', rec)          # restored by monkeypatch even if an assertion fails
        xc, ac = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
        yp = mc(xc, ac, nb, assign_x=xc)
        loss = mc.loss(yp, torch.tensor(label).cuda(), ac, nb)
        loss.backward()
        torch.cuda.synchronize()
        monkeypatch.setattr(_lib, '_hook', None)
        assert (rec.bf16_calls > 0) == (not f32)
        grads[f32] = np.concatenate([p.grad.cpu().numpy().ravel() for p in mc.parameters()]).astype(np.float64)
        assert rel_l2(yp.detach().cpu().numpy(), yo.detach().numpy()) < BF16_OUT
        assert abs(loss.item() - lo.item()) < BF16_LOSS * abs(lo.item())
    for f32 in (False, True):
        g = grads[f32]
        cos = float(g @ go / (np.linalg.norm(g) * np.linalg.norm(go)))
        assert rel_l2(g, go) < BF16_GRAD and cos > BF16_COS, (f32, rel_l2(g, go), cos)
    assert rel_l2(grads[False], grads[True]) < 1e-3
