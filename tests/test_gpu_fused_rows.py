"""Fused forward pieces of the tensor-core schedule against fp64 restatements of the reference ops:
gp_bgemm_bf16_norm (U.W + b -> L2 normalize in the GEMM epilogue, encoders.py:322-326, plus the BatchNorm row
sums), gp_bn_finalize / gp_bn_apply (encoders.py:1062-1064,1048-1052), gp_bias_normalize_x and the masked
softmax with bf16 copy / fused bias gradient (encoders.py:1273-1275)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_l2
from oracle import diffpool_oracle as orc

pytestmark = pytest.mark.gpu


def st():
    return torch.cuda.current_stream().cuda_stream


def dev(a, dt=torch.float32):
    return torch.tensor(np.ascontiguousarray(a), dtype=dt, device='cuda')


@pytest.mark.parametrize('B,N,din,dout,off,Fw,relu', [(16, 40, 128, 128, 0, 128, 1), (3, 50, 64, 64, 64, 200, 1),
                                                       (5, 33, 40, 200, 0, 200, 0), (2, 70, 128, 20, 8, 36, 1),
                                                       # rows wider than 256: the single-buffer 512-column variant
                                                       (4, 70, 128, 512, 0, 512, 0), (3, 50, 96, 320, 64, 392, 1),
                                                       (2, 140, 128, 500, 256, 756, 0)])
def test_norm_gemm_and_bn(B, N, din, dout, off, Fw, relu):
    from graph_pooling_b200 import engine as E, engine_tc as T
    from graph_pooling_b200._lib import call
    rs = np.random.RandomState(3)
    rows = B * N
    u = rs.randn(rows, din).astype(np.float32)
    u[5] = 0.0                                           # a pad row: Y = b / ||b||
    w = (rs.randn(din, dout) * 0.2).astype(np.float32)
    bias = (rs.randn(dout) * 0.3).astype(np.float32)
    ws = E.Workspace(torch.device('cuda'))
    ub = T.cvt(ws, dev(u).data_ptr(), din, rows, din)
    wb = T.cvt(ws, dev(w).data_ptr(), dout, din, dout)
    u64 = ub.t.float().cpu().double().reshape(rows, -1)[:, :din]
    w64 = wb.t.float().cpu().double().reshape(din, -1)[:, :dout]
    v = u64 @ w64 + torch.tensor(bias, dtype=torch.float64)
    nrm = v.norm(dim=1, keepdim=True).clamp_min(1e-12)
    y_ref = v / nrm
    ycat = torch.zeros(rows, Fw, device='cuda')
    yb = torch.zeros(rows, (Fw + 7) // 8 * 8, device='cuda', dtype=torch.bfloat16)
    rnorm = torch.empty(rows, device='cuda')
    rowstat = torch.empty(rows, 2, device='cuda')
    biasd = dev(bias)
    hb = T.Op(yb.data_ptr() + off * 2, yb.shape[1], 0) if off % 8 == 0 else None
    T._norm_gemm(T.Op(ub.ptr, ub.ld, 0), T.Op(wb.ptr, wb.ld, 0), rows, dout, din, biasd.data_ptr(),
                 (ycat.data_ptr() + off * 4, Fw, 0), hb, rnorm.data_ptr(), rowstat.data_ptr(), relu)
    torch.cuda.synchronize()
    got = ycat[:, off:off + dout].cpu().numpy()
    assert rel_l2(got, y_ref.numpy()) < 2e-6
    assert float(ycat[:, :off].abs().sum()) == 0.0 and float(ycat[:, off + dout:].abs().sum()) == 0.0
    assert rel_l2(rnorm.cpu().numpy(), nrm.numpy().ravel()) < 2e-6
    if hb is not None:
        assert rel_l2(yb[:, off:off + dout].float().cpu().numpy(), y_ref.numpy()) < 4e-3
    r = torch.relu(y_ref) if relu else y_ref
    assert rel_l2(rowstat[:, 0].cpu().numpy(), r.sum(1).numpy()) < 1e-5
    assert rel_l2(rowstat[:, 1].cpu().numpy(), (r * r).sum(1).numpy()) < 1e-5

    # BN statistics from the row sums, then the apply pass (fp32 slot + bf16 copy)
    mean, invstd = torch.empty(N, device='cuda'), torch.empty(N, device='cuda')
    call('gp_bn_finalize', rowstat.data_ptr(), B, N, dout, mean.data_ptr(), invstd.data_ptr(), st())
    y32 = ycat[:, off:off + dout].contiguous()
    h = torch.zeros(rows, Fw, device='cuda')
    hbb = torch.zeros(rows, (dout + 7) // 8 * 8, device='cuda', dtype=torch.bfloat16)
    hbb2 = torch.zeros(rows, (dout + 7) // 8 * 8 + 16, device='cuda', dtype=torch.bfloat16)
    call('gp_bn_apply', y32.data_ptr(), dout, mean.data_ptr(), invstd.data_ptr(), B, N, dout, relu, 1,
         h.data_ptr() + off * 4, Fw, hbb.data_ptr(), hbb.shape[1], hbb2.data_ptr() + 8 * 2, hbb2.shape[1], st())
    torch.cuda.synchronize()
    assert torch.equal(hbb2[:, 8:8 + dout], hbb[:, :dout])          # second bf16 destination (column half)
    r3 = r.reshape(B, N, dout)
    h_ref = orc.bn_per_node(r3).reshape(rows, dout)
    assert rel_l2(mean.cpu().numpy(), r3.mean(dim=(0, 2)).numpy()) < 1e-5
    assert rel_l2(h[:, off:off + dout].cpu().numpy(), h_ref.numpy()) < 2e-5
    assert rel_l2(hbb[:, :dout].float().cpu().numpy(), h_ref.numpy()) < 4e-3


@pytest.mark.parametrize('rows,d,ld', [(100, 512, 512), (37, 132, 140), (64, 1024, 1024), (41, 1256, 1256),
                                       (19, 2048, 2052)])
def test_bias_normalize_x(rows, d, ld):
    from graph_pooling_b200._lib import call
    rs = np.random.RandomState(0)
    v = rs.randn(rows, ld).astype(np.float32)
    v[3, :d] = 0.0
    ref = torch.tensor(v[:, :d], dtype=torch.float64)
    nrm = ref.norm(dim=1, keepdim=True).clamp_min(1e-12)
    vc = dev(v)
    yb = torch.zeros(rows, (d + 7) // 8 * 8, device='cuda', dtype=torch.bfloat16)
    rn = torch.empty(rows, device='cuda')
    call('gp_bias_normalize_x', vc.data_ptr(), None, rn.data_ptr(), C.c_longlong(rows), d, ld, 1, yb.data_ptr(),
         yb.shape[1], st())
    torch.cuda.synchronize()
    assert rel_l2(vc[:, :d].cpu().numpy(), (ref / nrm).numpy()) < 1e-6
    assert np.array_equal(vc[:, d:].cpu().numpy(), v[:, d:])
    assert rel_l2(yb[:, :d].float().cpu().numpy(), (ref / nrm).numpy()) < 4e-3
    assert rel_l2(rn.cpu().numpy(), nrm.numpy().ravel()) < 1e-6


@pytest.mark.parametrize('B,N,K', [(3, 40, 512), (2, 33, 12), (4, 20, 1024), (2, 10, 256), (3, 21, 1256), (2, 9, 2048),
                                   (2, 30, 640)])
def test_softmax_x(B, N, K):
    from graph_pooling_b200._lib import call
    rs = np.random.RandomState(1)
    t = (rs.randn(B, N, K) * 3).astype(np.float32)
    nb = rs.randint(1, N + 1, size=B).astype(np.int32)
    mask = (np.arange(N)[None, :] < nb[:, None]).astype(np.float64)[:, :, None]
    tt = torch.tensor(t, dtype=torch.float64, requires_grad=True)
    s_ref = torch.softmax(tt, dim=-1) * torch.tensor(mask)
    gs = rs.randn(B, N, K)
    (s_ref * torch.tensor(gs)).sum().backward()
    tc, nbd = dev(t), dev(nb, torch.int32)
    sb = torch.zeros(B, N, (K + 7) // 8 * 8, device='cuda', dtype=torch.bfloat16)
    call('gp_softmax_mask_fwd_x', tc.data_ptr(), nbd.data_ptr(), B, N, K, sb.data_ptr(), sb.shape[2], st())
    torch.cuda.synchronize()
    assert rel_l2(tc.cpu().numpy(), s_ref.detach().numpy()) < 1e-6
    assert rel_l2(sb[:, :, :K].float().cpu().numpy(), s_ref.detach().numpy()) < 4e-3
    if K % 4 == 0:                                     # rows up to 512 floats: single pass; up to 2048: two passes
        dsd = dev(gs)
        dt = torch.empty(B, N, K, device='cuda')
        dtb = torch.zeros_like(sb)
        dcol = torch.empty(K, device='cuda')
        wsb = torch.empty((148 * 16 + 256) * K, device='cuda')
        call('gp_softmax_mask_bwd_x', tc.data_ptr(), dsd.data_ptr(), nbd.data_ptr(), B, N, K, dt.data_ptr(),
             dtb.data_ptr(), dtb.shape[2], dcol.data_ptr(), wsb.data_ptr(), st())
        torch.cuda.synchronize()
        assert rel_l2(dt.cpu().numpy(), tt.grad.numpy()) < 1e-5
        assert rel_l2(dtb[:, :, :K].float().cpu().numpy(), tt.grad.numpy()) < 4e-3
        assert rel_l2(dcol.cpu().numpy(), tt.grad.numpy().reshape(-1, K).sum(0)) < 1e-4
