"""GPU: the chained pooling contraction gp_pool_chain_bf16 (A' = S^T A S in one launch, T = S^T A kept on chip:
TMEM -> bf16 shared-memory tile -> second tcgen05.mma; a two-CTA cluster with a DSMEM exchange for 256 < K <= 512).

Reference: encoders.py:1279 in fp64 on the bf16-rounded operands.  The kernel rounds T to bf16 between the two
products (as the three-launch schedule does), so A' is checked twice: tightly against T_kernel . S (the second product
given the kernel's own T) and at the bf16 bound against the unrounded S^T A S."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def lib_call(*a):
    from graph_pooling_b200._lib import call
    return call(*a)


def r8(n):
    return (n + 7) & ~7


def make_case(B, N, K, nb, seed, weighted=False):
    rs = np.random.RandomState(seed)
    S = np.zeros((B, N, r8(K)), np.float32)
    A = np.zeros((B, N, r8(N)), np.float32)
    for b in range(B):
        n = int(nb[b])
        logits = rs.randn(n, K) * 2.0
        e = np.exp(logits - logits.max(1, keepdims=True))
        S[b, :n, :K] = e / e.sum(1, keepdims=True)
        a = (rs.rand(n, n) < 0.1).astype(np.float32)
        a = np.maximum(a, a.T)
        if weighted:
            a = a * rs.rand(n, n).astype(np.float32)
        A[b, :n, :n] = a
    Sb = torch.tensor(S).bfloat16().cuda()
    Ab = torch.tensor(A).bfloat16().cuda()
    return Sb, Ab


def run_chain(Sb, Ab, nb, N, K, keep_t=True, order=None, garbage=False):
    B = Sb.shape[0]
    nbd = None if nb is None else torch.tensor(np.asarray(nb, np.int32)).cuda()
    fill = float('nan') if garbage else 0.0
    t = torch.full((B, K, r8(N)), fill, device='cuda', dtype=torch.bfloat16) if keep_t else None
    ap = torch.full((B, K, K), fill, device='cuda', dtype=torch.float32)
    apb = torch.full((B, K, r8(K)), fill, device='cuda', dtype=torch.bfloat16)
    lib_call('gp_pool_chain_bf16', Sb.data_ptr(), C.c_longlong(Sb.shape[2]), Ab.data_ptr(), C.c_longlong(Ab.shape[2]),
             None if nbd is None else nbd.data_ptr(), None if order is None else order.data_ptr(), B, N, K,
             None if t is None else t.data_ptr(), C.c_longlong(0 if t is None else t.shape[2]),
             ap.data_ptr(), C.c_longlong(K), apb.data_ptr(), C.c_longlong(apb.shape[2]), None)
    torch.cuda.synchronize()
    return t, ap, apb


def check(B, N, K, nb, seed, use_lim=True, weighted=False):
    Sb, Ab = make_case(B, N, K, nb, seed, weighted)
    t, ap, apb = run_chain(Sb, Ab, nb if use_lim else None, N, K, garbage=True)
    S64 = Sb.double().cpu().numpy()[:, :, :K]
    A64 = Ab.double().cpu().numpy()[:, :, :N]
    T64 = np.einsum('bnk,bnm->bkm', S64, A64)
    AP64 = np.einsum('bkm,bmj->bkj', T64, S64)
    tk = t.double().cpu().numpy()[:, :, :N]
    apk = ap.cpu().numpy()
    for b in range(B):
        n = int(nb[b]) if use_lim else N
        nc = min(N, ((n + 127) // 128) * 128)             # T columns the kernel writes (whole 128-node tiles)
        assert np.isfinite(tk[b, :, :nc]).all()
        assert rel_l2(tk[b, :, :nc], T64[b, :, :nc]) < 4e-3
        # the second product given the kernel's own (bf16) T: only fp32 accumulation order differs
        ref2 = tk[b, :, :nc] @ S64[b, :nc]
        assert np.isfinite(apk[b]).all()
        assert rel_l2(apk[b], ref2) < 1e-5, (b, rel_l2(apk[b], ref2))
        assert rel_l2(apk[b], AP64[b]) < 4e-3
    assert rel_l2(apb.float().cpu().numpy()[:, :, :K], apk) < 4e-3
    # inference form (T never written): bit-identical A'
    _, ap2, apb2 = run_chain(Sb, Ab, nb if use_lim else None, N, K, keep_t=False)
    assert torch.equal(ap2, ap) and torch.equal(apb2[:, :, :K], apb[:, :, :K])


@pytest.mark.parametrize('B,N,K', [(2, 128, 64), (3, 300, 104), (1, 40, 8), (2, 1000, 256), (5, 512, 128)])
def test_chain_one_cta_per_row_block(B, N, K):
    check(B, N, K, [N] * B, seed=N + K)


@pytest.mark.parametrize('B,N,K', [(2, 2048, 512), (3, 700, 320), (2, 256, 264), (4, 1024, 384)])
def test_chain_cta_pair_with_dsmem_exchange(B, N, K):
    check(B, N, K, [N] * B, seed=N + K + 1)


@pytest.mark.parametrize('B,N,K,nb', [(4, 1000, 256, [1000, 333, 64, 1]), (3, 2048, 512, [2048, 777, 129]),
                                       (6, 640, 160, [640, 0, 127, 128, 129, 500]), (3, 900, 456, [900, 130, 257])])
def test_chain_ragged_node_counts(B, N, K, nb):
    check(B, N, K, nb, seed=7 * N + K)
    check(B, N, K, nb, seed=7 * N + K, use_lim=False)     # without the tile skip: same values (zero padding)


def test_chain_weighted_adjacency_and_order():
    """Real-valued A (pooled levels feed A' into the next level) and a batch permutation (longest-first walk)."""
    B, N, K = 5, 384, 96
    nb = [100, 384, 17, 256, 300]
    check(B, N, K, nb, seed=3, weighted=True)
    Sb, Ab = make_case(B, N, K, nb, 11)
    _, ap0, _ = run_chain(Sb, Ab, nb, N, K)
    order = torch.tensor(np.argsort(-np.asarray(nb)).astype(np.int32)).cuda()
    _, ap1, _ = run_chain(Sb, Ab, nb, N, K, order=order)
    assert torch.equal(ap0, ap1)


def test_chain_many_work_items_per_cluster():
    """More work items than resident clusters: barrier phases and the TMEM double buffer wrap many times."""
    B, N, K = 160, 256, 512
    rs = np.random.RandomState(0)
    nb = rs.randint(1, N + 1, size=B)
    check(B, N, K, nb, seed=5)


def test_model_chain_equals_three_launch_schedule(monkeypatch):
    """The encoder's tensor-core forward / loss / backward through the chained kernel against the three-launch schedule
    (same bf16 T, same products: only the accumulation order inside a tile may differ), with two pooling levels so the
    pooled level's real-valued A' goes through the chain too; plus a no-grad forward (T never stored)."""
    from helpers import synth_batch
    from graph_pooling_b200 import encoders as enc
    N, D, H, Cc, B = 300, 16, 32, 3, 4
    x, adj, nb, label = synth_batch(2, B, N, D, 40, N, Cc, 0.05)
    outs = []
    from graph_pooling_b200 import engine_tc, _lib

    class Names:                                  # records which C-ABI entry points the step calls
        def __init__(self):
            self.seen = set()

        def begin(self, name, args):
            self.seen.add(name)

        def end(self, tok):
            pass

    for no_chain in ('', '1'):
        monkeypatch.setattr(engine_tc, 'CHAIN_POOLING', not no_chain)
        rec = Names()
        monkeypatch.setattr(_lib, '_hook', rec)          # restored by monkeypatch even if an assertion fails
        torch.manual_seed(0)
        m = enc.SoftPoolingGcnEncoder(N, D, H, H, Cc, 3, H, assign_ratio=0.25, num_pooling=2).cuda()
        m.precision = 1
        xt, at = torch.tensor(x).cuda(), torch.tensor(adj).cuda()
        y = m(xt, at, nb, assign_x=xt)
        loss = m.loss(y, torch.tensor(label).cuda(), at, nb)
        loss.backward()
        g = torch.cat([p.grad.flatten() for p in m.parameters()])
        with torch.no_grad():
            y_ng = m(xt, at, nb, assign_x=xt)
        assert torch.equal(y_ng, y.detach())
        monkeypatch.setattr(_lib, '_hook', None)
        assert ('gp_pool_chain_bf16' in rec.seen) == (not no_chain)       # the schedule under test really ran
        outs.append((y.detach().clone(), loss.detach().clone(), g.clone()))
    (y0, l0, g0), (y1, l1, g1) = outs
    assert rel_l2(y0.cpu().numpy(), y1.cpu().numpy()) < 1e-5
    assert abs(float(l0) - float(l1)) < 1e-5 * abs(float(l1))
    assert rel_l2(g0.cpu().numpy(), g1.cpu().numpy()) < 1e-4
