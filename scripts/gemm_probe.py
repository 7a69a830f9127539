"""Stand-alone launches of the v2 tcgen05 GEMM on cfg4-like shapes (for ncu / timing)."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_pooling_b200 import engine as E, engine_tc as T

def op(t): return T.Op(t.data_ptr(), t.shape[-1], t.shape[-2] * t.shape[-1], t)
dev = torch.device('cuda')
ws = E.Workspace(dev)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
N, H, K = 2048, 128, 512

def timeit(name, fn, flops, bytes_):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('%-28s %8.3f ms  %8.1f TFLOP/s  %7.1f GB/s(alg)' % (name, ms, flops / ms / 1e9, bytes_ / ms / 1e6), flush=True)

adj = (torch.rand(B, N, N, device=dev) < 0.01).to(torch.bfloat16)
x = torch.randn(B, N, H, device=dev).bfloat16()
if which in ('all', 'ax'):
    ub = T.bfbuf(ws, B, N, H)
    timeit('A.X  bf16 out', lambda: T.tcgemm(op(adj), 0, op(x), 1, N, H, N, B, Cb=ub), 2.0 * B * N * N * H, B * (N * N + 2 * N * H) * 2)
if which in ('all', 'ax2'):
    # the lock-step embedding + assignment GCN contraction: U = A.[h | a], 256 columns (engine_tc.dual_stack_forward)
    x2 = torch.randn(B, N, 2 * H, device=dev).bfloat16()
    ub2 = T.bfbuf(ws, B, N, 2 * H)
    timeit('A.[h|a] 256 cols bf16 out', lambda: T.tcgemm(op(adj), 0, op(x2), 1, N, 2 * H, N, B, Cb=ub2),
           2.0 * B * N * N * 2 * H, B * (N * N + 2 * N * 2 * H) * 2)
if which in ('all', 'tsa'):
    # pooling contraction T = S^T A (encoders.py:1279): M=K, N=N, k=N ; S M-major, A N-major
    s_ = torch.rand(B, N, K, device=dev).bfloat16()
    tb = T.bfbuf(ws, B, K, N)
    timeit('T=S^T.A bf16 out', lambda: T.tcgemm(op(s_), 1, op(adj), 1, K, N, N, B, Cb=tb), 2.0 * B * K * N * N, B * (N * N + 2 * N * K) * 2)
if which in ('all', 'uw'):
    u = torch.randn(1, B * N, H, device=dev).bfloat16(); w = torch.randn(1, H, H, device=dev).bfloat16()
    y = torch.empty(B * N, H, device=dev)
    timeit('U.W  fp32 out', lambda: T.tcgemm(op(u), 0, op(w), 1, B * N, H, H, 1, Cf=(y.data_ptr(), H, 0)), 2.0 * B * N * H * H, B * N * H * 6)
    yb = T.bfbuf(ws, 1, B * N, H)
    timeit('U.W  bf16 out', lambda: T.tcgemm(op(u), 0, op(w), 1, B * N, H, H, 1, Cb=yb), 2.0 * B * N * H * H, B * N * H * 4)
if which in ('all', 'beta'):
    s = torch.rand(B, N, K, device=dev).bfloat16(); g = torch.randn(B, N, N, device=dev).bfloat16()
    ds = torch.zeros(B, N, K, device=dev)
    timeit('G.S  fp32 beta=0', lambda: T.tcgemm(op(g), 0, op(s), 1, N, K, N, B, Cf=(ds.data_ptr(), K, N * K)), 2.0 * B * N * N * K, 0)
    timeit('G.S  fp32 beta=1', lambda: T.tcgemm(op(g), 0, op(s), 1, N, K, N, B, Cf=(ds.data_ptr(), K, N * K), beta=1.0), 2.0 * B * N * N * K, 0)
if which in ('all', 'link'):
    s = torch.softmax(torch.randn(B, N, K, device=dev), -1).bfloat16()
    nb = torch.full((B,), N, device=dev, dtype=torch.int32)
    timeit('linkloss fused', lambda: T.linkloss_forward(ws, op(s), op(adj), nb, B, N, K, True), 2.0 * B * N * N * K, B * N * N * 4)
    timeit('linkloss fused nograd', lambda: T.linkloss_forward(ws, op(s), op(adj), nb, B, N, K, False), 2.0 * B * N * N * K, B * N * N * 2)
if which in ('all', 'chain'):
    # chained A' = S^T A S (gp_pool_chain_bf16) against the two launches it replaces (T = S^T A, A' = T S)
    import ctypes as C
    from graph_pooling_b200._lib import call
    s_ = torch.softmax(torch.randn(B, N, K, device=dev), -1).bfloat16()
    tb = T.bfbuf(ws, B, K, N)
    ap = torch.empty(B, K, K, device=dev); apb = T.bfbuf(ws, B, K, K)
    fl = 2.0 * B * (K * N * N + K * K * N)
    by_t = B * (N * N + N * K + K * N) * 2 + B * K * K * 6
    def chain(keep):
        call('gp_pool_chain_bf16', s_.data_ptr(), C.c_longlong(K), adj.data_ptr(), C.c_longlong(N), None, None, B, N, K,
             tb.ptr if keep else None, C.c_longlong(tb.ld if keep else 0), ap.data_ptr(), C.c_longlong(K), apb.ptr,
             C.c_longlong(apb.ld), None)
    def two():
        T.tcgemm(op(s_), 1, op(adj), 1, K, N, N, B, Cb=tb)
        T.tcgemm(tb, 0, op(s_), 1, K, K, N, B, Cf=(ap.data_ptr(), K, K * K), Cb=apb)
    timeit('two launches T=S^T.A, A\'=T.S', two, fl, by_t + B * K * N * 2)
    timeit('chained, T stored (training)', lambda: chain(True), fl, by_t)
    timeit('chained, T on chip only', lambda: chain(False), fl, by_t - B * K * N * 2)
