"""gp_gcn_layer_bwd_x (vectorised single-pass layer-tail backward, layer_bwd.cu) against fp64 autograd of
the reference ops: normalize (encoders.py:323-326) -> ReLU -> BatchNorm-per-node-index (:1062-1064,1048-1052)
-> concat slot / max readout (:1078,1097).  Covers the cluster kernel (d in {32,64,128}, cluster sizes 1..8),
the warp-per-row kernel (no BN, d up to 512) and the generic fallback (d % 4 != 0)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import diffpool_oracle as orc
from helpers import rel_l2

pytestmark = pytest.mark.gpu


def st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize('B,N,d,relu,bn,use_dz,use_dxn,use_readout,stored_h', [
    (256, 9, 128, 1, 1, 1, 1, 1, 0),      # cfg4 shape: cluster of 4 CTAs, 8 rows per thread
    (40, 17, 128, 1, 1, 1, 0, 1, 1),      # single CTA, stored BN output
    (300, 5, 64, 1, 1, 0, 1, 1, 0),       # cluster of 2..4, d4 = 16 (half-warp rows)
    (33, 20, 32, 1, 1, 1, 1, 0, 0),       # d4 = 8
    (7, 30, 128, 0, 0, 1, 1, 1, 0),       # last layer: no relu / bn -> row kernel
    (5, 21, 512, 0, 0, 1, 0, 0, 0),       # row kernel, 4 float4 per lane
    (6, 11, 256, 1, 0, 1, 1, 1, 0),       # row kernel with relu, 2 float4 per lane
    (3, 37, 1256, 0, 0, 1, 0, 0, 0),      # wide row kernel (two passes, 4 warps per block), 16 float4 slots per lane
    (4, 13, 768, 0, 0, 1, 1, 0, 0),       # wide row kernel, 8 slots
    (2, 150, 1280, 1, 0, 1, 1, 1, 0),     # wide row kernel, 16 slots, relu + readout scatter
    (2, 9, 2048, 0, 0, 1, 1, 0, 0),       # widest vectorised row
    (20, 100, 30, 1, 1, 1, 1, 1, 1),      # generic (d % 4 != 0), stored h
    (20, 40, 30, 1, 1, 1, 0, 1, 0),       # generic, recomputed Hhat
    (4, 6, 256, 1, 1, 1, 0, 0, 0),        # bn with d4 = 64 -> generic
])
def test_layer_bwd_x(B, N, d, relu, bn, use_dz, use_dxn, use_readout, stored_h):
    from graph_pooling_b200._lib import GpLayerBwd, call, load
    rs = np.random.RandomState(B * 7 + d)
    F = d + 12                                     # the layer's slot lives inside a wider concat buffer
    off = 4
    v = rs.randn(B, N, d)
    v[0, 0] = 0.0                                  # a zero row: normalize clamps (dV = dY / eps)
    gz = rs.randn(B, N, d) if use_dz else np.zeros((B, N, d))
    gx = rs.randn(B, N, d) if use_dxn else np.zeros((B, N, d))
    go = rs.randn(B, d) if use_readout else np.zeros((B, d))

    vt = torch.tensor(v, dtype=torch.float64, requires_grad=True)
    r = vt.norm(dim=2, keepdim=True).clamp_min(1e-12)
    y = vt / r
    h = torch.relu(y) if relu else y
    if bn:
        h = orc.bn_per_node(h)
    out, arg = h.max(dim=1)
    loss = (h * torch.tensor(gz + gx)).sum() + (out * torch.tensor(go)).sum()
    loss.backward()
    dv_ref = vt.grad.numpy()

    dev = lambda a, dt=torch.float32: torch.tensor(np.ascontiguousarray(a), dtype=dt, device='cuda')
    yc, rc = dev(y.detach().numpy()), dev(r.detach().numpy().reshape(B, N))
    rc[0, 0] = 1e-12
    hslot = torch.zeros(B, N, F, device='cuda')
    hslot[:, :, off:off + d] = dev(h.detach().numpy())
    x = (torch.relu(y) if relu else y).detach()
    mean = x.mean(dim=(0, 2))
    invstd = 1.0 / torch.sqrt(x.var(dim=(0, 2), unbiased=False) + 1e-5)
    dz = torch.zeros(B, N, F, device='cuda')
    dz[:, :, off:off + d] = dev(gz)
    ldo = F
    dout = torch.zeros(B, ldo, device='cuda')
    dout[:, off:off + d] = dev(go)
    argi = torch.full((B, ldo), -1, dtype=torch.int32, device='cuda')
    argi[:, off:off + d] = dev(arg.numpy(), torch.int32)
    dxn = dev(gx)

    dv = torch.empty(B, N, d, device='cuda')
    ldb = (d + 7) // 8 * 8
    dvb = torch.zeros(B, N, ldb, device='cuda', dtype=torch.bfloat16)
    db = torch.empty(d, device='cuda')
    ws = torch.empty(int(load().gp_gcn_layer_bwd_ws(B, N, d, bn)), device='cuda')
    q = GpLayerBwd()
    q.dz, q.lddz = (dz.data_ptr() + off * 4 if use_dz else None), F
    q.dxn = dxn.data_ptr() if use_dxn else None
    q.dout = dout.data_ptr() + off * 4 if use_readout else None
    q.argidx = argi.data_ptr() + off * 4 if use_readout else None
    q.ldo = ldo
    q.h, q.ldh = (hslot.data_ptr() + off * 4 if (stored_h and bn) else None), F
    q.y, q.ldy = yc.data_ptr(), d
    q.rnorm, q.mean, q.invstd = rc.data_ptr(), dev(mean.numpy()).data_ptr(), dev(invstd.numpy()).data_ptr()
    q.B, q.N, q.d, q.relu, q.bn, q.normalize = B, N, d, relu, bn, 1
    q.dv, q.dv_bf16, q.lddvb = dv.data_ptr(), dvb.data_ptr(), ldb
    q.db, q.ws = db.data_ptr(), ws.data_ptr()
    mean_d, invstd_d = dev(mean.numpy()), dev(invstd.numpy())
    q.mean, q.invstd = mean_d.data_ptr(), invstd_d.data_ptr()
    call('gp_gcn_layer_bwd_x', C.byref(q), st())
    torch.cuda.synchronize()
    got = dv.cpu().numpy()
    # the clamped (zero) row has gradient dY / 1e-12: compare it separately on a relative basis
    mask = np.ones((B, N), bool)
    mask[0, 0] = False
    assert rel_l2(got[mask], dv_ref[mask]) < 2e-5
    assert rel_l2(got[0, 0], dv_ref[0, 0]) < 1e-4
    gb = dvb[:, :, :d].float().cpu().numpy()
    assert rel_l2(gb[mask], dv_ref[mask]) < 6e-3
    assert rel_l2(db.cpu().numpy(), got.reshape(-1, d).sum(0)) < 1e-4

    # bf16-only output with db (no fp32 dV requested): the engine_tc configuration
    if d % 4 == 0:
        q.dv = None
        db2 = torch.empty(d, device='cuda')
        dvb2 = torch.zeros_like(dvb)
        q.dv_bf16, q.db = dvb2.data_ptr(), db2.data_ptr()
        call('gp_gcn_layer_bwd_x', C.byref(q), st())
        torch.cuda.synchronize()
        # (BN layers with many rows per thread take the two-CTAs-per-SM kernel, which stashes the upstream gradient as
        # bf16 between its two phases: same bound as the bf16 output itself, not bit-identical to the fp32-dV run)
        gb2 = dvb2[:, :, :d].float().cpu().numpy()
        assert rel_l2(gb2[mask], dv_ref[mask]) < 8e-3
        assert rel_l2(db2.cpu().numpy(), db.cpu().numpy()) < 2e-3


@pytest.mark.parametrize('B,N,d', [(3, 40, 128), (2, 33, 512), (2, 25, 1256)])
def test_layer_bwd_row_nb_zero_hint(B, N, d):
    """nb_zero (layers without BN): rows n >= nb_zero[b] are declared gradient-free by the caller -- their dV is written
    as zero without reading the operands.  With an upstream gradient that IS zero there, the result must equal the
    unhinted run bit for bit (row kernels of every width: 1 / 4 slots per lane and the two-pass wide kernel)."""
    from graph_pooling_b200._lib import GpLayerBwd, call, load
    rs = np.random.RandomState(d)
    nb = rs.randint(1, N + 1, size=B).astype(np.int32)
    dz = rs.randn(B, N, d).astype(np.float32)
    for b in range(B):
        dz[b, nb[b]:] = 0.0
    y = rs.randn(B, N, d).astype(np.float32)
    y /= np.linalg.norm(y, axis=2, keepdims=True)
    dev = lambda a, dt=torch.float32: torch.tensor(np.ascontiguousarray(a), dtype=dt, device='cuda')
    dzc, yc, rn, nbc = dev(dz), dev(y), dev(rs.rand(B, N) + 0.5), dev(nb, torch.int32)
    outs = []
    for hint in (0, 1):
        ldb = (d + 7) // 8 * 8
        dvb = torch.full((B, N, ldb), 7.0, device='cuda', dtype=torch.bfloat16)
        db = torch.empty(d, device='cuda')
        q = GpLayerBwd()
        q.dz, q.lddz, q.y, q.ldy, q.rnorm = dzc.data_ptr(), d, yc.data_ptr(), d, rn.data_ptr()
        q.B, q.N, q.d, q.relu, q.bn, q.normalize = B, N, d, 0, 0, 1
        q.dv, q.dv_bf16, q.lddvb, q.db = None, dvb.data_ptr(), ldb, db.data_ptr()
        q.nb_zero = nbc.data_ptr() if hint else None
        w = torch.empty(int(load().gp_gcn_layer_bwd_ws_x(C.byref(q))), device='cuda')
        q.ws = w.data_ptr()
        call('gp_gcn_layer_bwd_x', C.byref(q), st())
        torch.cuda.synchronize()
        outs.append((dvb[:, :, :d].clone(), db.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel_l2(outs[1][1].cpu().numpy(), outs[0][1].cpu().numpy()) < 1e-6
    for b in range(B):
        assert float(outs[1][0][b, nb[b]:].float().abs().sum()) == 0.0
