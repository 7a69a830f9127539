"""GPU: the tcgen05/TMA bf16 GEMM (gp_bgemm_bf16) against an exact reference.

Inputs are rounded to bf16 first, so products are exact in fp32 and the only difference to the
fp64 reference is fp32 accumulation order: tolerance 2e-6 rel-L2 (K <= 4096)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def run_tc(A, Bm, M, N, K, batch, a_major, b_major, lim=None, lim_m=0, lim_n=0, lim_k=0, alpha=1.0, beta=0.0,
           C0=None, split_k=0, want_bf16=False, bias=None, relu=0):
    """A: bf16 tensor stored [batch, M, K] (a_major 0) or [batch, K, M] (1); same for B with N."""
    from graph_pooling_b200._lib import GpGemmBf16, call
    out = torch.zeros(batch, M, N, device='cuda') if C0 is None else C0.clone()
    ob = torch.zeros(batch, M, N, device='cuda', dtype=torch.bfloat16) if want_bf16 else None
    ldA, ldB = A.shape[2], Bm.shape[2]
    g = GpGemmBf16(A.data_ptr(), Bm.data_ptr(), out.data_ptr(), None if ob is None else ob.data_ptr(),
                   M, N, K, batch, ldA, A.shape[1] * ldA, a_major, ldB, Bm.shape[1] * ldB, b_major,
                   N, M * N, N, M * N, None if lim is None else lim.data_ptr(), lim_m, lim_n, lim_k,
                   alpha, beta, None, None if bias is None else bias.data_ptr(), relu, split_k)
    call('gp_bgemm_bf16', C.byref(g), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out, ob


def mk(batch, M, N, K, a_major, b_major, seed):
    rs = np.random.RandomState(seed)
    Kp, Mp, Np = -(-K // 8) * 8, -(-M // 8) * 8, -(-N // 8) * 8
    A = torch.tensor(rs.randn(batch, M, K), dtype=torch.float32).bfloat16()
    Bm = torch.tensor(rs.randn(batch, K, N), dtype=torch.float32).bfloat16()
    A64, B64 = A.double().numpy(), Bm.double().numpy()
    if a_major == 0:
        As = torch.zeros(batch, M, Kp, dtype=torch.bfloat16); As[:, :, :K] = A
    else:
        As = torch.zeros(batch, K, Mp, dtype=torch.bfloat16); As[:, :, :M] = A.transpose(1, 2)
    if b_major == 0:
        Bs = torch.zeros(batch, N, Kp, dtype=torch.bfloat16); Bs[:, :, :K] = Bm.transpose(1, 2)
    else:
        Bs = torch.zeros(batch, K, Np, dtype=torch.bfloat16); Bs[:, :, :N] = Bm
    return As.cuda(), Bs.cuda(), A64, B64


@pytest.mark.parametrize('a_major', [0, 1])
@pytest.mark.parametrize('b_major', [0, 1])
@pytest.mark.parametrize('M,N,K,batch', [(128, 128, 64, 1), (256, 128, 512, 2), (128, 64, 128, 1), (300, 200, 150, 3),
                                          (512, 256, 1024, 2), (70, 40, 33, 2), (2048, 128, 2048, 2)])
def test_tc_gemm_all_majors(a_major, b_major, M, N, K, batch):
    A, Bm, A64, B64 = mk(batch, M, N, K, a_major, b_major, M + N + K + a_major * 2 + b_major)
    out, _ = run_tc(A, Bm, M, N, K, batch, a_major, b_major)
    assert rel_l2(out.cpu().numpy(), A64 @ B64) < 5e-6


def test_tc_gemm_limits_beta_alpha_bf16_copy():
    batch, M, N, K = 3, 384, 128, 384
    A, Bm, A64, B64 = mk(batch, M, N, K, 0, 1, 7)
    lim_np = np.array([384, 100, 257], np.int32)
    lim = torch.tensor(lim_np).cuda()
    # contract of lim_k (include/gp_b200.h): one operand is ZERO beyond the clipped K extent, as the
    # zero-padded adjacency / masked S are on the DiffPool path (tiles are skipped, not masked)
    for b in range(batch):
        A[b, :, lim_np[b]:] = 0
        A64[b, :, lim_np[b]:] = 0
    C0 = torch.randn(batch, M, N, device='cuda')
    out, ob = run_tc(A, Bm, M, N, K, batch, 0, 1, lim=lim, lim_m=1, lim_k=1, alpha=0.5, beta=1.0, C0=C0,
                     want_bf16=True)
    ref = C0.cpu().double().numpy().copy()
    for b in range(batch):
        l = lim_np[b]
        ref[b, :l] += 0.5 * (A64[b, :l, :l] @ B64[b, :l])
    assert rel_l2(out.cpu().numpy(), ref) < 2e-6
    assert rel_l2(ob.float().cpu().numpy(), ref) < 4e-3          # bf16 rounding of the output copy


def test_tc_gemm_splitk_and_bias_relu():
    batch, M, N, K = 1, 128, 128, 8192
    A, Bm, A64, B64 = mk(batch, M, N, K, 1, 1, 9)                # dW = U^T dV pattern: both MN-major
    out, _ = run_tc(A, Bm, M, N, K, batch, 1, 1, split_k=16)
    assert rel_l2(out.cpu().numpy(), A64 @ B64) < 5e-6
    bias = torch.randn(N, device='cuda')
    A, Bm, A64, B64 = mk(2, 200, 96, 128, 0, 0, 11)
    bias = torch.randn(96, device='cuda')
    out, _ = run_tc(A, Bm, 200, 96, 128, 2, 0, 0, bias=bias, relu=1)
    ref = np.maximum(A64 @ B64 + bias.cpu().double().numpy(), 0)
    assert rel_l2(out.cpu().numpy(), ref) < 2e-6


def test_cvt_bf16_padding():
    from graph_pooling_b200._lib import call
    x = torch.randn(37, 89, device='cuda')
    y = torch.full((37, 96), 5.0, device='cuda', dtype=torch.bfloat16)
    call('gp_cvt_f32_bf16', x.data_ptr(), 89, y.data_ptr(), 96, 37, 89, 96, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(y[:, :89], x.bfloat16())
    assert float(y[:, 89:].float().abs().sum()) == 0.0
