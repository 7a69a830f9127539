"""GPU: the persistent multi-pair tcgen05 GEMM (gp_bgemm_bf16x) and the fused tensor-core link loss.

Inputs are bf16-rounded first, so the reference differs only by fp32 accumulation order (5e-6)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_l2, synth_batch
from oracle import diffpool_oracle as orc

pytestmark = pytest.mark.gpu


def T():
    from graph_pooling_b200 import engine_tc
    return engine_tc


def op(t):
    """[batch, rows, cols] bf16 cuda tensor -> Op"""
    return T().Op(t.data_ptr(), t.shape[2], t.shape[1] * t.shape[2], t)


def store(x64, major_is_mn):
    """logical [batch, MN, K] matrix -> bf16 storage for the requested major-ness (+ its fp64 value)."""
    t = torch.tensor(x64, dtype=torch.float32).bfloat16()
    v = t.double().numpy()
    s = t.transpose(1, 2).contiguous() if major_is_mn else t.contiguous()
    pad = (-s.shape[2]) % 8
    if pad:
        s = torch.nn.functional.pad(s, (0, pad))
    return s.cuda(), v


@pytest.mark.parametrize('M,N,K,batch,am,bm', [
    (128, 128, 64, 1, 0, 1), (2048, 128, 2048, 3, 0, 1), (512, 2048, 1024, 2, 1, 1), (300, 200, 150, 3, 1, 0),
    (2048, 512, 512, 2, 0, 0), (70, 40, 33, 2, 0, 0), (1024, 384, 256, 40, 0, 1), (128, 64, 4096, 1, 1, 1)])
def test_persistent_single_pair(M, N, K, batch, am, bm):
    rs = np.random.RandomState(M + N + K)
    A, A64 = store(rs.randn(batch, M, K), am)
    Bm, B64t = store(rs.randn(batch, N, K), bm)          # logical B^T [N, K]
    out = torch.full((batch, M, N), 3.0, device='cuda')
    T().tcgemm_multi([(op(A), am, op(Bm), bm, K, 0)], M, N, batch, Cf=(out.data_ptr(), N, M * N))
    torch.cuda.synchronize()
    ref = A64 @ np.swapaxes(B64t, 1, 2)
    assert rel_l2(out.cpu().numpy(), ref) < 5e-6


def test_multi_pair_accumulation_with_limits_and_beta():
    """dS-style: three products with different shapes / major-ness into one accumulator."""
    rs = np.random.RandomState(5)
    batch, M, N = 3, 384, 200
    K1, K2, K3 = 96, 200, 384
    lim_np = np.array([384, 130, 64], np.int32)
    A1, A1v = store(rs.randn(batch, M, K1), 0)
    B1, B1v = store(rs.randn(batch, N, K1), 0)
    A2, A2v = store(rs.randn(batch, M, K2), 1)
    B2, B2v = store(rs.randn(batch, N, K2), 1)
    A3h = rs.randn(batch, M, K3)
    for b in range(batch):                               # lim_k contract: zero beyond the limit
        A3h[b, :, lim_np[b]:] = 0
    A3, A3v = store(A3h, 0)
    B3, B3v = store(rs.randn(batch, N, K3), 1)
    C0 = torch.randn(batch, M, N, device='cuda')
    out = C0.clone()
    ob = torch.zeros(batch, M, N, device='cuda', dtype=torch.bfloat16)
    lim = torch.tensor(lim_np).cuda()
    T().tcgemm_multi([(op(A1), 0, op(B1), 0, K1, 0), (op(A2), 1, op(B2), 1, K2, 0), (op(A3), 0, op(B3), 1, K3, 1)],
                     M, N, batch, Cf=(out.data_ptr(), N, M * N), Cb=op(ob), lim=lim.data_ptr(), lim_m=1, alpha=0.5,
                     beta=1.0)
    torch.cuda.synchronize()
    ref = C0.cpu().double().numpy().copy()
    full = 0.5 * (A1v @ np.swapaxes(B1v, 1, 2) + A2v @ np.swapaxes(B2v, 1, 2) + A3v @ np.swapaxes(B3v, 1, 2))
    for b in range(batch):
        ref[b, :lim_np[b]] += full[b, :lim_np[b]]
    assert rel_l2(out.cpu().numpy(), ref) < 5e-6
    assert rel_l2(ob.float().cpu().numpy(), ref) < 4e-3


def test_split_k_bias_relu():
    rs = np.random.RandomState(6)
    A, Av = store(rs.randn(1, 128, 16384), 1)
    Bm, Bv = store(rs.randn(1, 96, 16384), 1)
    out = torch.zeros(1, 128, 96, device='cuda')
    T().tcgemm_multi([(op(A), 1, op(Bm), 1, 16384, 0)], 128, 96, 1, Cf=(out.data_ptr(), 96, 0), split_k=37)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu().numpy(), Av @ np.swapaxes(Bv, 1, 2)) < 5e-6
    A, Av = store(rs.randn(2, 200, 128), 0)
    Bm, Bv = store(rs.randn(2, 72, 128), 0)
    bias = torch.randn(72, device='cuda')
    out = torch.zeros(2, 200, 72, device='cuda')
    T().tcgemm_multi([(op(A), 0, op(Bm), 0, 128, 0)], 200, 72, 2, Cf=(out.data_ptr(), 72, 200 * 72),
                     bias=bias.data_ptr(), relu=1)
    torch.cuda.synchronize()
    ref = np.maximum(Av @ np.swapaxes(Bv, 1, 2) + bias.cpu().double().numpy(), 0)
    assert rel_l2(out.cpu().numpy(), ref) < 5e-6


@pytest.mark.parametrize('B,N,K,use_nb,sym,weighted,flag', [
    (3, 300, 40, 1, 1, 0, 0), (2, 512, 128, 1, 0, 0, 0), (2, 260, 16, 0, 1, 0, 0), (2, 200, 24, 1, 0, 1, 0),
    # symmetric adjacency with the device flag: upper-band tiles only, mirrored G chunks, single-product backward
    (3, 300, 40, 1, 1, 0, 1), (2, 768, 64, 1, 1, 0, 1), (2, 1024, 32, 0, 1, 0, 1), (2, 520, 24, 1, 1, 1, 1),
    # N a multiple of 32, symmetric {0,1}: G kept as its upper band only (mode 2), backward through gp_gemm_bf16x.tri;
    # the weighted one keeps the mirrored writes but its backward still reads the band twice
    (3, 288, 40, 1, 1, 0, 1), (2, 2048, 64, 1, 1, 0, 1), (4, 640, 24, 0, 1, 0, 1), (2, 544, 24, 1, 1, 1, 1),
    (5, 1280, 136, 1, 1, 0, 1)])
def test_fused_linkloss_tc(B, N, K, use_nb, sym, weighted, flag):
    """gp_linkloss_tc + the (G + G^T).S backward vs the fp64 oracle on bf16-rounded S / adjacency."""
    from graph_pooling_b200._lib import call
    t = T()
    _, adj, nb, _ = synth_batch(8, B, N, 2, N // 3, N, 2, density=0.1, symmetric=bool(sym), weighted=bool(weighted))
    nbo = nb if use_nb else None
    rs = np.random.RandomState(6)
    m = orc.construct_mask(N, nb if use_nb else np.full(B, N), 'cpu', torch.float64)
    s32 = (torch.softmax(torch.tensor(rs.randn(B, N, K) * 2), dim=-1) * m).float()
    sbt = s32.bfloat16()
    adjb_t = torch.tensor(adj).bfloat16()
    s_t = sbt.double().requires_grad_()
    lo = orc.link_pred_loss(s_t, adjb_t.double(), nbo)
    lo.backward()
    Kp, Np = -(-K // 8) * 8, (-(-N // 32) * 32 if flag else -(-N // 8) * 8)   # with flags: rows padded for the row epilogue
    sb = torch.zeros(B, N, Kp, dtype=torch.bfloat16); sb[:, :, :K] = sbt
    ab = torch.zeros(B, N, Np, dtype=torch.bfloat16); ab[:, :, :N] = adjb_t
    sb, ab = sb.cuda(), ab.cuda()
    nbc = torch.tensor(nb).cuda() if use_nb else None
    ws = __import__('graph_pooling_b200.engine', fromlist=['x']).Workspace(torch.device('cuda'))
    asym = torch.tensor([0, int(bool(weighted))], device='cuda', dtype=torch.int32) if flag else None   # gp_adj_prepare flags
    partial, npart, gs, upper = t.linkloss_forward(ws, op(sb), op(ab), nbc, B, N, K, True, adj_flags=asym)
    assert upper == bool(flag)
    entries = float(np.sum(nb.astype(np.int64) ** 2)) if use_nb else float(B * N * N)
    total, link = torch.empty(1, device='cuda'), torch.empty(1, device='cuda')
    call('gp_loss_finalize', partial.data_ptr(), npart, C.c_double(1.0 / entries), None, total.data_ptr(),
         link.data_ptr(), torch.cuda.current_stream().cuda_stream)
    one = torch.ones(1, device='cuda')
    dS = t.linkloss_backward(ws, gs, op(sb), nbc, B, N, K, 1.0 / entries, one.data_ptr(), asym=None if asym is None else asym[0:1],
                             upper=upper)
    torch.cuda.synchronize()
    assert abs(link.item() - lo.item()) < 2e-5 * abs(lo.item())          # __logf + fp32 sums
    mm = m.numpy()
    # G is rounded to bf16 before the backward GEMM: 2^-9 relative per entry
    assert rel_l2(dS.cpu().numpy() * mm, s_t.grad.numpy() * mm) < 4e-3


@pytest.mark.parametrize('M,N,batch', [(2048, 512, 3), (1280, 300, 5), (256, 256, 40), (5000, 264, 2)])
def test_multicast_pair_ragged_multi_pair(M, N, batch):
    """Shapes that take the CTA-pair schedule (cta_group::2: 256-row blocks, each CTA staging half of every B tile):
    a dS-style launch with three operand pairs of all major-ness combinations, per-graph limits on M and on one
    contraction, beta = 1 and a bf16 copy; more work items than resident clusters at batch 40."""
    rs = np.random.RandomState(M + N)
    K1, K2 = 192, 136
    lim_np = rs.randint(1, M + 1, size=batch).astype(np.int32)
    lim_np[0] = M
    A1, A1v = store(rs.randn(batch, M, K1), 0)
    B1, B1v = store(rs.randn(batch, N, K1), 0)
    A2, A2v = store(rs.randn(batch, M, K2), 1)
    B2, B2v = store(rs.randn(batch, N, K2), 1)
    A3h = rs.randn(batch, M, M) * 0.1
    for b in range(batch):
        A3h[b, :, lim_np[b]:] = 0
    A3, A3v = store(A3h, 0)
    B3, B3v = store(rs.randn(batch, N, M), 1)
    C0 = torch.randn(batch, M, N, device='cuda')
    out = C0.clone()
    ob = torch.zeros(batch, M, r8(N), device='cuda', dtype=torch.bfloat16)
    lim = torch.tensor(lim_np).cuda()
    T().tcgemm_multi([(op(A1), 0, op(B1), 0, K1, 0), (op(A2), 1, op(B2), 1, K2, 0), (op(A3), 0, op(B3), 1, M, 1)],
                     M, N, batch, Cf=(out.data_ptr(), N, M * N), Cb=op(ob), lim=lim.data_ptr(), lim_m=1, alpha=0.5,
                     beta=1.0)
    torch.cuda.synchronize()
    ref = C0.cpu().double().numpy().copy()
    full = 0.5 * (A1v @ np.swapaxes(B1v, 1, 2) + A2v @ np.swapaxes(B2v, 1, 2) + A3v @ np.swapaxes(B3v, 1, 2))
    for b in range(batch):
        ref[b, :lim_np[b]] += full[b, :lim_np[b]]
    assert rel_l2(out.cpu().numpy(), ref) < 2e-5      # contractions up to 5000 terms long in fp32
    assert rel_l2(ob.float().cpu().numpy()[:, :, :N], ref) < 4e-3


def r8(n):
    return (n + 7) & ~7
