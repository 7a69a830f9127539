"""Times the fused tensor-core link loss (gp_linkloss_tc) and its backward alone.  usage: python scripts/link_probe.py [B N K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_pooling_b200 import engine as E, engine_tc as T
B, N, K = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (256, 2048, 512)
dev = torch.device('cuda')
ws = E.Workspace(dev)
sb = T.bfbuf(ws, B, N, K)
sb.t.copy_(torch.softmax(torch.randn(B, N, sb.t.shape[2], device=dev), -1))
adj = (torch.rand(B, N, N, device=dev) < 0.01)
adj = (adj | adj.transpose(1, 2)).float()
adjb, flags = T.adj_prepare(ws, adj, None, B, N)
del adj
one = torch.ones(1, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print('flags', flags.tolist())
for name, asym in (('general', None), ('adjacency flags', flags)):
    out = {}
    def fwd():
        out['r'] = T.linkloss_forward(ws, sb, adjb, None, B, N, K, True, adj_flags=asym)
    t = timeit(fwd)
    gs = out['r'][2]
    up = out['r'][3]
    tb = timeit(lambda: T.linkloss_backward(ws, gs, sb, None, B, N, K, 1e-6, one.data_ptr(), asym=None if asym is None else asym[0:1], upper=up))
    fl = 2.0 * B * N * N * K
    print('%-15s fwd %.3f ms (%.0f TFLOP/s of the full P)   bwd %.3f ms' % (name, t, fl / t / 1e9, tb))
